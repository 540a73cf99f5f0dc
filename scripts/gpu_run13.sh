#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1; echo "exit=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
PT="python -m pytest -q --tb=short -p no:cacheprovider"
run parity_sizes 900 $PT tests/test_gpu_parity.py -m gpu -k "named_sizes"
# one full step (warm-up: 3 steps + 2 e2e... skip with -s), every launch with time + DRAM bytes + tensor-pipe activity
run ncu_step 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 1300 -c 330 --csv --log-file gpurun_out/step_metrics.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline
cat gpurun_out/summary.txt
