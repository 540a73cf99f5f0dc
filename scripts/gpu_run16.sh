#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-200)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
ECGVIT_PDL=0 run prof_p01 600 python scripts/profile_step.py 0.1 gpurun_out/profiler_step_p01.json
ECGVIT_PDL=0 run prof_p0 600 python scripts/profile_step.py 0.0 gpurun_out/profiler_step_p0.json
cat gpurun_out/summary.txt
