#!/usr/bin/env python
"""CUDA-event timing of gtos_attn_decode / gtos_token_topk at the config-5 shapes (2048 hypotheses)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtos_b200 import decode, ops          # noqa: E402


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = torch.device("cuda:0")
    Hyp, B, K = 2048, 256, 8
    for name, L, H, hd, mode in [("cross S=40", 40, 8, 64, "cross"), ("self t=16", 17, 8, 64, "self"), ("self t=31", 32, 8, 64, "self"),
                                 ("align S=40", 40, 1, 512, "cross")]:
        D = H * hd
        q = torch.randn(Hyp, D, device=dev)
        if mode == "cross":
            cache = torch.randn(L, B, 2 * D, device=dev).to(torch.bfloat16)
            slot = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(K)
            pad = torch.zeros(L, B, dtype=torch.uint8, device=dev)
            fn = lambda: decode.attn_decode(q, (cache, 0), 2 * D, D, B, L, H, hd, slot, 0, pad, B, hd ** -0.5, H == 1)
            byts = cache.numel() * 2
        else:
            cache = torch.randn(L, Hyp, 2 * D, device=dev).to(torch.bfloat16)
            anc = (torch.arange(Hyp, device=dev, dtype=torch.int32) // K * K).unsqueeze(0).repeat(L, 1).contiguous()
            anc = anc + torch.randint(0, K, (L, Hyp), device=dev, dtype=torch.int32)
            fn = lambda: decode.attn_decode(q, (cache, 0), 2 * D, D, Hyp, L, H, hd, anc, Hyp, None, 0, hd ** -0.5)
            byts = cache.numel() * 2
        us = timeit(fn)
        print(f"attn_decode {name:12s}: {us:7.1f} us  (cache {byts / 1e6:.0f} MB -> {byts / us / 1e3:.0f} GB/s if read once)")
    V, S = 10000, 40
    W = V + 16
    logits = torch.randn(Hyp, V, device=dev) * 3
    gate = torch.randn(Hyp, 2, device=dev)
    align = torch.softmax(torch.randn(Hyp, S, device=dev), -1)
    copy_seq = torch.randint(2, W, (S, B), device=dev)
    slot = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(K)
    us = timeit(lambda: ops.token_topk(logits, gate, align, copy_seq, slot, W, K))
    print(f"token_topk  [{Hyp} x {W}] k={K}: {us:7.1f} us  ({logits.numel() * 4 / us / 1e3:.0f} GB/s of logits)")
    us = timeit(lambda: ops.token_logprob(logits, gate, align, copy_seq, slot, W))
    print(f"token_logprob (table)        : {us:7.1f} us")
    tab = ops.token_logprob(logits, gate, align, copy_seq, slot, W)
    us = timeit(lambda: torch.topk(tab, K, dim=1))
    print(f"torch.topk on the table      : {us:7.1f} us")


if __name__ == "__main__":
    main()
