#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests (memcheck on the non-GEMM kernels, racecheck on attention)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention or layernorm or head or patchify" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_fwd_bwd and 1-" 2>&1 | tail -3
