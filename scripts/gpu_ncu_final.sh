#!/bin/bash
# final ncu evidence: --set full on the dominant kernel family (three GEMM flavours), plus one full-step launch list
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-200)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
GEMM_BENCH_NO_CUBLAS=1 run ncu_gemm 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -c 8 -o gpurun_out/prof_gemm_final python scripts/gemm_bench.py 1 qkv_fwd,ff1_fwd,ff2_fwd,ff1_wgrad
run ncu_step 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 1400 -c 340 --csv --log-file gpurun_out/step_metrics.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline
cat gpurun_out/summary.txt
