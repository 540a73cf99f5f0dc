#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
nvidia-smi -L > gpurun_out/gpus.txt
PT="python -m pytest -q --tb=short -p no:cacheprovider"
run ddp_test 900 $PT tests/test_gpu_ddp.py -m gpu
run bench1 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline
run bench2 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3
run bench2_nograph 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 --no-graph
run bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
cat gpurun_out/summary.txt
