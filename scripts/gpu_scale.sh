#!/bin/bash
# 8-GPU visit: NCCL data-parallel tests, then the weak-scaling curve N = 1, 2, 4, 8 and the reference arm under torchrun
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(grep -E '^\{"metric"|^\{"impl"|passed|failed' gpurun_out/$name.log | tail -n 1 | cut -c1-200)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
nvidia-smi -L > gpurun_out/gpus.txt
run ddp_test 600 python -m pytest -q --tb=short -p no:cacheprovider tests/test_gpu_ddp.py -m gpu
run scale1 300 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline
for n in 2 4 8; do
  run scale$n 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) bench.py --gpus $n --steps 30 --warmup 3
done
run ref8 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 8 --steps 1 --warmup 1
cat gpurun_out/summary.txt
