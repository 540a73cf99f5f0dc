#!/bin/bash
# quick single-GPU visit: GPU tests + default bench + in-graph kernel profile
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-240)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
run pytest_gpu 1800 python -m pytest tests/ -x -q -m gpu
run bench 600 python bench.py --no-cpu-baseline --profile-json gpurun_out/profile.json
ECGVIT_PDL=0 ECGVIT_WGRAD_STREAM=0 run prof 600 python scripts/profile_step.py 0.1 gpurun_out/profiler_step_p01.json
cat gpurun_out/summary.txt
