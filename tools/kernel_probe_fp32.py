#!/usr/bin/env python
"""One encoder layer forward + backward in fp32 mode at config-2 size (relation attention on split operands), for
    ncu --set full --clock-control none --import-source on -k regex:<name> python tools/kernel_probe_fp32.py
Kernels: the K-tripled projection GEMM of the relation rows (gemm_tn_kernel on [P, 3D] x [2D, 3D]^T), gtos_split3,
gtos_rel_score_f32 / gtos_rel_grad_f32 / gtos_rel_dqk_f32, the three-pass attention core."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtos_b200 import ops  # noqa: E402
from gtos_b200.graph_transformer import GraphTransformer  # noqa: E402

dev = torch.device("cuda:0")
N, B, D, H = 41, 64, 512, 8
torch.manual_seed(0)
m = GraphTransformer(1, D, 1024, H, 0.0).to(dev)
x = torch.randn(N, B, D, device=dev, requires_grad=True)
rel = (torch.randn(N, N, B, D, device=dev) * 0.5).requires_grad_()
mask = torch.zeros(N, B, dtype=torch.bool, device=dev)
with ops.precision_mode("fp32"):
    for _ in range(2):
        out = m(x, rel, self_padding_mask=mask)
        out.sum().backward()
torch.cuda.synchronize()
print("done")
