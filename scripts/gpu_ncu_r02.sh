#!/bin/bash
# round-2 ncu evidence (1 GPU): --set full of the kernels VERDICT r01 asked for, plus the launch list of one eager step
mkdir -p gpurun_out
cap() { local name=$1 regex=$2 skip=$3; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c 1 \
    -o gpurun_out/$name -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/$name.log 2>&1; echo "ncu $name exit=$?"; }
cap r02_ncu_ln_fwd 'layernorm_fwd' 100
cap r02_ncu_ln_bwd 'layernorm_bwd_kernel' 100
cap r02_ncu_attn_fwd 'attn_tc_fwd' 50
cap r02_ncu_attn_bwd 'attn_tc_bwd' 50
cap r02_ncu_wgrad 'gemm_tc2_kernel<256, true, true' 200
cap r02_ncu_gelu_fwd 'gemm_tc2_kernel<256, false, false, 2' 50
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 1100 -c 290 --csv --log-file gpurun_out/r02_step_metrics.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/r02_step.log 2>&1
echo "step list exit=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_step_metrics.csv
