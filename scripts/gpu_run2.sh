#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PT="python -m pytest -q --tb=short -p no:cacheprovider"
run k_attn_adamw 600 $PT tests/test_gpu_kernels.py -m gpu -k "attention or adamw"
run parity_all 900 $PT tests/test_gpu_parity.py -m gpu
run gemm_bench 600 python scripts/gemm_bench.py 20
run bench 900 python bench.py --steps 10 --warmup 3 --profile-json gpurun_out/profile.json
run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline
GEMM_BENCH_NO_CUBLAS=1 run ncu_gemm 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 12 -o gpurun_out/prof_gemm python scripts/gemm_bench.py 1 ff1_fwd,ff2_fwd,ff1_wgrad
cat gpurun_out/summary.txt
