#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests: memcheck on every kernel family (tcgen05 attention and GEMM included),
# racecheck on the attention kernels (shared-memory staging between the softmax warps, the MMA issuer and the TMA warp)
mkdir -p gpurun_out
san() { local name=$1 tool=$2; shift 2; timeout -s KILL 1200 compute-sanitizer --tool $tool --error-exitcode 9 "$@" > gpurun_out/sanitize_$name.log 2>&1; echo "== $name ($tool) exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$name.log | tail -1) | $(tail -n 1 gpurun_out/sanitize_$name.log)"; }
san mem_kernels memcheck python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention or layernorm or head or patchify or adamw or sumsq"
san mem_attn_tc memcheck python -m pytest tests/test_gpu_attention_tc.py -x -q -m gpu -k "3-51-12 or 3-51-3 or 2-1-2 or 7-17-5 or 2-65-3 or 1-129-1 or 1-321-1 or 3-51-4 or 1-200-2"
san mem_gemm memcheck python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm"
san race_attn_tc racecheck python -m pytest tests/test_gpu_attention_tc.py -x -q -m gpu -k "3-51-3 or 7-17-5 or 2-65-3 or 1-129-1"
