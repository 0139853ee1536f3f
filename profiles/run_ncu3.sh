#!/bin/bash
# per-launch duration + DMMA pipe utilisation for a window of launches (cheap: 2 metrics)
TAG=${1:-r01d}
CMD="python bench.py --experts-per-step 592 --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/busy_$TAG.csv $CMD > gpurun_out/ncu_busy_$TAG.log 2>&1
