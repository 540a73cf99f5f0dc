#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
run bench_p01 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/profile_p01.json
run bench_p0 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --dropout 0 --profile-json gpurun_out/profile_p0.json
cat gpurun_out/summary.txt
