#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(grep -E '^\{"metric"' gpurun_out/$name.log | tail -n 1 | cut -c1-200)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
nvidia-smi -L > gpurun_out/gpus.txt
run scale1 300 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline
run scale8 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 30 --warmup 3
run ref8 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 1 --warmup 1
cat gpurun_out/summary.txt
