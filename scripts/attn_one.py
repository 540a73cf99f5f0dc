#!/usr/bin/env python
"""One attention forward (+ backward) launch at a given shape, for ncu:  attn_one.py B N H [p] [bwd]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecg_b200 import _lib as L

B, N, H = (int(a) for a in sys.argv[1:4])
p = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
bwd = len(sys.argv) > 5
dh = 64
lib = L.load()
inner = H * dh
st = torch.cuda.current_stream().cuda_stream
qkv = torch.randn(B * N, 3 * inner, device='cuda').bfloat16()
o = torch.empty(B * N, inner, device='cuda', dtype=torch.bfloat16)
d_o = torch.randn(B * N, inner, device='cuda').bfloat16()
dqkv = torch.empty_like(qkv)
lse = torch.empty(B, H, N, device='cuda')
seed = torch.tensor([1234], dtype=torch.int32, device='cuda')
n_scr = int(lib.ecgvit_attention_bwd_scratch_floats(B, N, H, dh, L.BF16))
scr = torch.empty(max(n_scr, 1), device='cuda')
for _ in range(3):
    L.check(lib.ecgvit_attention_fwd(qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, N, H, dh, dh ** -0.5, p, 1,
                                     seed.data_ptr() if p > 0 else None, L.BF16, st), 'attn')
    if bwd:
        L.check(lib.ecgvit_attention_bwd(qkv.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                         scr.data_ptr() if n_scr else None, B, N, H, dh, dh ** -0.5, p, 1,
                                         seed.data_ptr() if p > 0 else None, L.BF16, st), 'attn_bwd')
torch.cuda.synchronize()
