#!/usr/bin/env python
"""LayerNorm fwd / bwd alone at the cfg2 shape (CUDA-graph replay over 12 buffer sets, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ecg_b200 import _lib
peaks = bench.measured_peaks()
cfg = dict(bench.BASE_CFG)
rows = bench.hbm_kernel_rooflines(torch, _lib, cfg, 256, 0.1, 85584455, peaks, {})
for r in rows:
    print(f"{r['kernel'][:44]:44s} {r['ms_per_launch']*1e3:8.2f} us  {r['achieved']:7.0f} GB/s  frac {r['frac']:.3f}")
