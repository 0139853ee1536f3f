CMD="python bench.py --experts-per-step 592 --steps 1 --warmup 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_lauum2 -s 4 -c 1 -o gpurun_out/prof_r01c_lauum -f $CMD > gpurun_out/ncu_full_r01c_lauum.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trtri_pass1 -s 23 -c 1 -o gpurun_out/prof_r01c_trtri -f $CMD > gpurun_out/ncu_full_r01c_trtri.log 2>&1
