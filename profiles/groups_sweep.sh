timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "optimise or kat1" > gpurun_out/t9.log 2>&1
for G in 1 2 3 4; do GPSAT_GROUPS=$G timeout 600 python bench.py --experts-per-step 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_g$G.log 2>&1; done
GPSAT_GROUPS=3 GPSAT_MAX_SLOTS=1184 timeout 600 python bench.py --experts-per-step 2048 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_g3_s1184.log 2>&1
