#!/bin/bash
# One GPU-box visit: kernel tests (isolated per risky group), parity tests, smoke, short bench. Logs -> gpurun_out/.
# Usage: scripts/gpu_check.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -s KILL $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
PT="python -m pytest -q --tb=short -p no:cacheprovider"
run k_simt 900 $PT tests/test_gpu_kernels.py -m gpu -k "not tcgen05 and not wgrad and not persistent and not epilogues"
for lay in 1-1 1-0 0-0 0-1; do
  run k_tc_$lay 300 $PT tests/test_gpu_kernels.py -m gpu -k "tcgen05_layouts and $lay"
done
run k_tc_rest 600 $PT tests/test_gpu_kernels.py -m gpu -k "wgrad or persistent or epilogues"
run parity_fp32 900 $PT tests/test_gpu_parity.py -m gpu -k "fp32 or golden or shorter or reference_style or accumulation"
run parity_bf16 900 $PT tests/test_gpu_parity.py -m gpu -k "bf16 or stock or graph"
run parity_base 900 $PT tests/test_gpu_parity.py -m gpu -k "base_model"
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
if [ "$1" != "quick" ]; then
  run bench_nograph 900 python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --profile-json gpurun_out/profile_nograph.json
  run bench 900 python bench.py --steps 10 --warmup 3 --profile-json gpurun_out/profile.json
fi
cat gpurun_out/summary.txt
