#!/bin/bash
# full-set captures of the two FP64 elementwise kernels (kernel-matrix build, gradient trace)
CMD="python bench.py --experts-per-step 592 --steps 1 --warmup 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"k_build|k_grad_trace" -s 6 -c 2 -o gpurun_out/prof_r01e_elem -f $CMD > gpurun_out/ncu_full_r01e.log 2>&1
