#!/bin/bash
# scaling visit on N GPUs (gpurun --gpus N): weak scaling (256 per GPU), BASELINE configs[2] as written (global 2048),
# and at N = 8 also configs[4] (large, global 4096) and an NCCL CTA-cap variant.   bash scripts/gpu_scale_r02.sh N
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
S=gpurun_out/scale_n${N}_summary.txt
: > $S
run() { local name=$1; shift; timeout -s KILL 400 "$@" > gpurun_out/$name.log 2>&1; echo "== $name rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$name.log | head -1) $(grep -o '"value": [0-9.]*' gpurun_out/$name.log | head -1)" | tee -a $S; grep '^{' gpurun_out/$name.log | tail -1 > gpurun_out/$name.json; }
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline"
run s${N}_n1 python $B --gpus 1
run s${N}_weak $TR --master-port 29801 $B --gpus $N
run s${N}_cfg3_global2048 $TR --master-port 29802 $B --gpus $N --global-batch 2048
if [ "$N" = "8" ]; then
  run s8_weak_ctas16 $TR --master-port 29803 $B --gpus 8 --nccl-max-ctas 16
  run s8_weak_fp32reduce $TR --master-port 29804 $B --gpus 8 --grad-reduce fp32
  run s8_cfg5_global4096 $TR --master-port 29805 $B --gpus 8 --config cfg5
  run s8_cfg5_n1 python $B --gpus 1 --config cfg5
  $TR --master-port 29807 scripts/ddp_timeline.py gpurun_out/tl_n8_bf16.json > gpurun_out/tl_n8_bf16.log 2>&1
fi
cat $S
