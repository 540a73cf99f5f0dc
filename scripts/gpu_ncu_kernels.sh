#!/bin/bash
# ncu --set full on named non-GEMM kernels of one eager step: bash scripts/gpu_ncu_kernels.sh <regex> <count> <outname>
mkdir -p gpurun_out
REGEX=${1:-attention}; COUNT=${2:-4}; OUT=${3:-prof_kernels}
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -c "$COUNT" \
    -o gpurun_out/$OUT -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/$OUT.log 2>&1
echo "ncu $OUT exit=$?" | tee -a gpurun_out/summary.txt
