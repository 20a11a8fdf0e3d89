#!/usr/bin/env python
"""CUDA-graph back-to-back time of the small non-GEMM kernels of the step at config-2 shapes (token NLL forward, column
sums, LayerNorm parameter gradients)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtos_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


def graph_time(fn, n=30):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


T, B, V, S = 60, 64, 10000, 40
logits = torch.randn(T, B, V, device=dev)
gate = torch.randn(T, B, 2, device=dev)
align = torch.softmax(torch.randn(T, B, S, device=dev), -1)
copy_seq = torch.randint(2, V + 16, (S, B), device=dev)
target = torch.randint(2, V, (T, B), device=dev)
print(f"token_nll fwd [3840, 10000]: {graph_time(lambda: ops.token_nll(logits, gate, align, copy_seq, target, 0)):.1f} us "
      f"(154 MB read once = {154e6 / 6.5e12 * 1e6:.0f} us at 6.5 TB/s)")
for rows, cols in [(2624, 512), (2624, 1536), (3840, 1536), (3840, 10000)]:
    x = torch.randn(rows, cols, device=dev)
    out = torch.empty(cols, device=dev)
    print(f"colsum [{rows}, {cols}]: {graph_time(lambda: ops.colsum(x, out=out)):.1f} us")
rows, D = 2624, 512
dy, z = torch.randn(rows, D, device=dev), torch.randn(rows, D, device=dev)
mean, rstd = torch.randn(rows, device=dev), torch.rand(rows, device=dev)
dg, db = torch.empty(D, device=dev), torch.empty(D, device=dev)
st = torch.cuda.current_stream


def lnpg():
    _lib.check(lib.gtos_ln_param_grad(dy.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                      rows, D, torch.cuda.current_stream().cuda_stream), "ln_param_grad")


print(f"ln_param_grad [{rows}, {D}]: {graph_time(lnpg):.1f} us")
