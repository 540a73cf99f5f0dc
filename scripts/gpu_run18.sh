#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PT="python -m pytest -q --tb=short -p no:cacheprovider"
run parity_all 1200 $PT tests/test_gpu_parity.py tests/test_gpu_dropout.py -m gpu
run bench_side_p01 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline
ECGVIT_WGRAD_STREAM=0 run bench_noside_p01 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline
run bench_side_p0 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --dropout 0
cat gpurun_out/summary.txt
