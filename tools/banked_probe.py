#!/usr/bin/env python
"""a few launches of the two bank-factorised relation kernels at config-2 size (for ncu -k regex:banked)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtos_b200 import _lib, ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
g = synthetic.make_graph_batch(64, 40, max_path_len=4)
idx = g["relation"].to(dev)
N, B, D, H = g["N"], 64, 512, 8
R = g["relation_bank"].shape[1]
st = torch.cuda.current_stream().cuda_stream
bank = torch.randn(R, D, device=dev)
W = torch.randn(2 * D, D, device=dev) * 0.02
Wperm, _ = ops.weight_prep(W, rel_heads=H)
_, PB = ops.gemm_tn(ops.cast_bf16(bank), Wperm, 2 * D, f32=False, bf16=True)
NB = N * B
qk = torch.randn(NB, 2 * D, device=dev).to(torch.bfloat16)
v = torch.randn(NB, D, device=dev)
pad = (torch.arange(N, device=dev).unsqueeze(1) >= (g["node_counts"].to(dev) + 1).unsqueeze(0)).to(torch.uint8).contiguous()
probs = torch.empty(B, H, N, N, device=dev)
att = torch.empty(NB, D, device=dev)
attb = torch.empty(NB, D, dtype=torch.bfloat16, device=dev)
seed = ops.rng_state(dev)
ds = torch.randn(B, H, N, N, device=dev)
G = torch.empty(ops.rel_tiling(N, B, D, H)["tiles"] * 128, 2 * D, dtype=torch.bfloat16, device=dev)
for _ in range(4):
    _lib.check(lib.gtos_rel_attn_banked_fwd(PB.data_ptr(), PB.stride(0), idx.data_ptr(), qk.data_ptr(), qk.data_ptr() + 2 * D, 2 * D,
                                            v.data_ptr(), D, pad.data_ptr(), None, 0.2, seed.data_ptr(), 12345, probs.data_ptr(), None,
                                            att.data_ptr(), D, attb.data_ptr(), N, B, D, H, R, st))
    _lib.check(lib.gtos_rel_grad_banked(PB.data_ptr(), PB.stride(0), idx.data_ptr(), qk.data_ptr(), qk.data_ptr() + 2 * D, 2 * D,
                                        ds.data_ptr(), G.data_ptr(), N, B, D, H, R, st))
torch.cuda.synchronize()
print("done")
