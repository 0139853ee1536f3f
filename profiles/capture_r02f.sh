#!/bin/bash
# Final round-2 captures at HEAD (after the padding / triangular-half skipping): launch list and per-launch DMMA-pipe
# utilisation + DRAM bytes of the default c3 bench command, after the plain command has exited 0.
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/r02f_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1200 --csv \
    --log-file gpurun_out/r02f_launches.csv $CMD > gpurun_out/r02f_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02f_busy.csv $CMD > gpurun_out/r02f_ncu_busy.log 2>&1
