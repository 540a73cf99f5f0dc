#!/bin/bash
# 2-GPU visit: NCCL data-parallel tests (eager + graph) and the N=1 / N=2 bench pair
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(grep -E '^\{"metric"|passed|failed' gpurun_out/$name.log | tail -n 1 | cut -c1-220)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
run ddp_test 600 python -m pytest -q --tb=short -p no:cacheprovider tests/test_gpu_ddp.py -m gpu
run bench1 300 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/profile.json
run bench2 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 30 --warmup 3
cat gpurun_out/summary.txt
