#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PT="python -m pytest -q --tb=short -p no:cacheprovider"
run k_ln 300 $PT tests/test_gpu_kernels.py -m gpu -k "layernorm"
run parity_all 900 $PT tests/test_gpu_parity.py -m gpu
run bench 900 python bench.py --steps 30 --warmup 3 --profile-json gpurun_out/profile.json
GEMM_BENCH_NO_CUBLAS=1 run ncu_gemm 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -c 6 -o gpurun_out/prof_gemm_pair python scripts/gemm_bench.py 1 ff1_fwd,ff2_fwd,ff2_dgrad
run ncu_misc 900 ncu --set full --clock-control none --import-source on -k regex:"attention_bwd|layernorm_bwd|layernorm_fwd" -s 30 -c 6 -o gpurun_out/prof_misc python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline
cat gpurun_out/summary.txt
