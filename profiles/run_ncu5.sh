#!/bin/bash
# HBM-side kernels (selection, gathers, kernel-matrix build, gradient trace): duration + DRAM bytes per launch
CMD="python bench.py --experts-per-step 1024 --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"k_select|k_cell|k_gather|k_build|k_grad_trace|k_slot_init|k_pred" -c 60 --csv \
    --log-file gpurun_out/hbm_r01f.csv $CMD > gpurun_out/ncu_hbm_r01f.log 2>&1
