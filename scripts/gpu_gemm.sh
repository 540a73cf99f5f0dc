#!/bin/bash
# GEMM-only visit: kernel tests, then every GEMM shape of the step in isolation (with the cuBLAS time beside it)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" 2>&1 | tail -3
timeout 600 python scripts/gemm_bench.py 30 2>&1 | grep "'name'" | cut -c1-170
