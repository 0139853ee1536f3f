#!/bin/bash
# Round-2 captures (run on the GPU box through gpurun, after the plain command has exited 0):
#   1. launch list of the bench command           -> gpurun_out/r02_launches.csv   (profiles/summarize.py)
#   2. per-launch DMMA-pipe utilisation + DRAM     -> gpurun_out/r02_busy.csv       (profiles/summarize_busy.py)
#   3. ncu --set full of three consecutive middle Cholesky panel launches -> gpurun_out/prof_r02_potrf.ncu-rep
#      (profiles/traffic_from_rep.py turns it into profiles/r02_traffic.json, the source of bench.py's roofline.traffic;
#       profiles/summarize_busy.py lists the DRAM bytes of EVERY launch of the round from pass 2)
# The command is bench.py's default batch configuration (1024 experts per step: 3 slot groups of 197-198 slots).
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/r02_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1200 --csv \
    --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02_busy.csv $CMD > gpurun_out/r02_ncu_busy.log 2>&1
# (three launches: a full-set report with source is ~7 MB per launch and gpurun copies back at most 64 MiB)
ncu --set full --clock-control none --import-source on -k regex:k_potrf_panel -s 49 -c 3 \
    -o gpurun_out/prof_r02_potrf -f $CMD > gpurun_out/r02_ncu_full.log 2>&1
