#!/usr/bin/env python
"""A few launches of each kernel the round-2 profiles cover, at config-2 size, for
    ncu --set full --clock-control none --import-source on -k regex:<name> -s <skip> -c <n> python tools/kernel_probe.py
Kernels: the fused dense relation attention (gemm_tn_kernel<256, 2, 2> with the attention tail), gtos_rel_grad, the plain
projection GEMM (gemm_tn_kernel<128, 0, 1>) at [N*B, D] x [D, D]^T, the weight-gradient GEMM (gemm_nn_kernel), the bank
gather kernels and the sorted bank segment sum."""
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
NB = N * B
bank = torch.randn(R, D, device=dev)
br = ops.BankedRelation(bank, idx)
relb = br.relb
W = torch.randn(2 * D, D, device=dev) * 0.02
Wperm, WpermT = ops.weight_prep(W, rel_heads=H)
_, PB = ops.gemm_tn(br.bankb, Wperm, 2 * D, f32=False, bf16=True)
qkv = torch.randn(NB, 3 * D, device=dev).to(torch.bfloat16)
vf = torch.randn(NB, D, device=dev)
pad = (torch.arange(N, device=dev).unsqueeze(1) >= (g["node_counts"].to(dev) + 1).unsqueeze(0)).to(torch.uint8).contiguous()
probs = torch.empty(B, H, N, N, device=dev)
att = torch.empty(NB, D, device=dev)
attb = torch.empty(NB, D, dtype=torch.bfloat16, device=dev)
seed = ops.rng_state(dev)
ds = torch.randn(B, H, N, N, device=dev)
G = torch.empty(ops.rel_tiling(N, B, D, H)["tiles"] * 128, 2 * D, dtype=torch.bfloat16, device=dev)
xa = torch.randn(NB, D, device=dev).to(torch.bfloat16)
wb = torch.randn(D, D, device=dev).to(torch.bfloat16)
yo = torch.empty(NB, D, device=dev)
dW = torch.empty(D, D, device=dev)
drel = torch.randn(N * N * B, D, device=dev)
keys, order = torch.sort(idx.view(-1))
dbank = torch.empty(R, D, device=dev)
scores = torch.empty(B, H, N, N, device=dev)
fused = lib.gtos_rel_attn_fusable(N, B, D, H) == 1           # needs GTOS_REL_FUSED_FWD=1 (full-row tiles)
print("fused dense relation attention:", fused, ops.rel_tiling(N, B, D, H))
for _ in range(3):
    if fused:
        _lib.check(lib.gtos_rel_attn_fwd(relb.data_ptr(), Wperm.data_ptr(), qkv.data_ptr(), qkv.data_ptr() + 2 * D, 3 * D,
                                         qkv.data_ptr() + 4 * D, 3 * D, pad.data_ptr(), 0.2, seed.data_ptr(), 12345,
                                         probs.data_ptr(), None, att.data_ptr(), D, attb.data_ptr(), N, B, D, H, st), "rel_attn_fwd")
    else:
        _lib.check(lib.gtos_rel_score(relb.data_ptr(), Wperm.data_ptr(), qkv.data_ptr(), qkv.data_ptr() + 2 * D, 3 * D,
                                      scores.data_ptr(), N, B, D, H, st), "rel_score")
    _lib.check(lib.gtos_rel_grad(relb.data_ptr(), Wperm.data_ptr(), qkv.data_ptr(), qkv.data_ptr() + 2 * D, 3 * D, ds.data_ptr(),
                                 G.data_ptr(), N, B, D, H, st), "rel_grad")
    ops.gemm_tn(xa, wb, D, out=yo)
    ops.gemm_nn(xa, xa, D, D, out=dW)
    _lib.check(lib.gtos_rel_attn_banked_fwd(PB.data_ptr(), PB.stride(0), idx.data_ptr(), qkv.data_ptr(), qkv.data_ptr() + 2 * D, 3 * D,
                                            vf.data_ptr(), D, pad.data_ptr(), None, 0.2, seed.data_ptr(), 12345, probs.data_ptr(), None,
                                            att.data_ptr(), D, attb.data_ptr(), N, B, D, H, R, st), "banked_fwd")
    _lib.check(lib.gtos_rel_grad_banked(PB.data_ptr(), PB.stride(0), idx.data_ptr(), qkv.data_ptr(), qkv.data_ptr() + 2 * D, 3 * D,
                                        ds.data_ptr(), G.data_ptr(), N, B, D, H, R, st), "banked_grad")
    _lib.check(lib.gtos_bank_segsum(drel.data_ptr(), order.data_ptr(), keys.data_ptr(), N * N * B, D, dbank.data_ptr(), R, st),
               "bank_segsum")
torch.cuda.synchronize()
print("done")
