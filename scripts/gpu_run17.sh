#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PT="python -m pytest -q --tb=short -p no:cacheprovider"
run k_all 900 $PT tests/test_gpu_kernels.py -m gpu
run parity_all 1200 $PT tests/test_gpu_parity.py tests/test_gpu_dropout.py -m gpu
run bench_p01 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline
run bench_p0 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --dropout 0
ECGVIT_PDL=0 run prof_p01 600 python scripts/profile_step.py 0.1 gpurun_out/profiler_step_p01.json
cat gpurun_out/summary.txt
