#!/bin/bash
# round-2 ncu evidence, part b: --set full of the GEMM flavours (wgrad <256,T,T,4>, both GELU epilogues, plain store),
# the long-geometry attention kernels, and a re-capture of the LayerNorm kernels after their rewrite
mkdir -p gpurun_out
GEMM_BENCH_NO_CUBLAS=1 timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -c 10 \
    -o gpurun_out/r02_ncu_gemm -f python scripts/gemm_bench.py 1 qkv_fwd,ff1_fwd,ff2_dgrad,ff1_wgrad,qkv_wgrad > gpurun_out/r02_ncu_gemm.log 2>&1
echo "gemm exit=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 4 -c 2 \
    -o gpurun_out/r02_ncu_attn_long -f python scripts/attn_one.py 4 2401 12 0.1 bwd > gpurun_out/r02_ncu_attn_long.log 2>&1
echo "attn long exit=$?"
cap() { local name=$1 regex=$2 skip=$3; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c 1 \
    -o gpurun_out/$name -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/$name.log 2>&1; echo "ncu $name exit=$?"; }
cap r02_ncu_ln_fwd 'layernorm_fwd' 100
cap r02_ncu_ln_bwd 'layernorm_bwd_kernel' 100
cap r02_ncu_attn_fwd 'attn_tc_fwd' 50
cap r02_ncu_attn_bwd 'attn_tc_bwd' 50
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 1100 -c 290 --csv --log-file gpurun_out/r02_step_metrics.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/r02_step.log 2>&1
echo "step list exit=$?"
ls -la gpurun_out/r02_*.ncu-rep
