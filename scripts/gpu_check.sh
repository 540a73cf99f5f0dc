#!/bin/bash
# What the driver runs at round end, in one GPU-box visit: the GPU test tier, smoke(), the default bench and the reference arm.
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run pytest_gpu 1800 python -m pytest tests/ -x -q -m gpu
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
run bench_ref 900 python bench.py --impl reference --steps 3 --warmup 1
run bench 900 python bench.py --profile-json gpurun_out/profile.json
cat gpurun_out/summary.txt
