#!/bin/bash
# Round-2 closing visit: what the driver runs at round end (GPU test tier, smoke(), reference arm, default bench), then the
# in-graph kernel times, every GEMM shape of the step against cuBLAS, and the other configurations.
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run pytest_gpu 1800 python -m pytest tests/ -x -q -m gpu
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
run bench_ref 900 python bench.py --impl reference
run bench 900 python bench.py --profile-json gpurun_out/profile.json
ECGVIT_PDL=0 ECGVIT_WGRAD_STREAM=0 run prof 600 python scripts/profile_step.py 0.1 gpurun_out/r02_ingraph_final.json
run gemm_bench 600 python scripts/gemm_bench.py 50
run bench_api 600 python bench.py --api-loop --no-cpu-baseline
run bench_cfg4 600 python bench.py --config cfg4 --no-cpu-baseline --steps 10 --warmup 3
run bench_cfg5 600 python bench.py --config cfg5 --no-cpu-baseline --steps 10 --warmup 3
cat gpurun_out/summary.txt
