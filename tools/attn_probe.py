#!/usr/bin/env python
"""Time the attention core (gtos_attn_fwd / gtos_attn_bwd) at the config-2 shapes and print the phase timestamps of CTA 0."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtos_b200 import _lib, ops          # noqa: E402


def timeit(fn, n=30):
    for _ in range(3):
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
    lib = _lib.load()
    buf = (ctypes.c_uint64 * 48)()
    for name, T, S, B, H, hd, enc in [("encoder N=41", 41, 41, 64, 8, 64, True), ("decoder self T=60", 60, 60, 64, 8, 64, False),
                                      ("decoder cross 60x40", 60, 40, 64, 8, 64, False), ("align 60x40 1 head", 60, 40, 64, 1, 512, False)]:
        D = H * hd
        q = torch.randn(T * B, D, device=dev)
        k = torch.randn(S * B, 2 * D, device=dev)
        scores = torch.randn(B, H, S, T, device=dev)
        probs = torch.empty(B, H, T, S, device=dev)
        out = torch.empty(T * B, D, device=dev)
        outb = torch.empty(T * B, D, dtype=torch.bfloat16, device=dev)
        pad = torch.zeros(S, B, dtype=torch.uint8, device=dev)
        seed = ops.rng_state(dev)
        dout = torch.randn(T * B, D, device=dev)
        ds_jt = torch.empty(B, H, S, T, device=dev)
        ds_ts = torch.empty(B, H, T, S, device=dev)
        dq = torch.empty(T * B, D, device=dev)
        dkv = torch.empty(S * B, 2 * D, device=dev)

        def desc():
            d = ops._attn_desc(T, S, B, H, hd)
            d.v, d.ldv = k.data_ptr() + 4 * D, 2 * D
            d.scale, d.p_drop = (1.0 if enc else hd ** -0.5), 0.2
            d.key_pad = pad.data_ptr()
            d.seed_ptr, d.seed_off = seed.data_ptr(), 12345
            d.probs = probs.data_ptr()
            d.out, d.ldo, d.out_bf16 = out.data_ptr(), D, outb.data_ptr()
            if enc:
                d.scores_jt = scores.data_ptr()
            else:
                d.q, d.ldq, d.k, d.ldk = q.data_ptr(), D, k.data_ptr(), 2 * D
            return d

        def fwd():
            d = desc()
            _lib.check(lib.gtos_attn_fwd(ctypes.byref(d), torch.cuda.current_stream().cuda_stream))

        def bwd():
            d = desc()
            d.dout, d.lddo = dout.data_ptr(), D
            d.dscores_ts = ds_ts.data_ptr()
            d.dv, d.lddv = dkv.data_ptr() + 4 * D, 2 * D
            if enc:
                d.dscores_jt = ds_jt.data_ptr()
            else:
                d.dq, d.lddq, d.dk, d.lddk = dq.data_ptr(), D, dkv.data_ptr(), 2 * D
            _lib.check(lib.gtos_attn_bwd(ctypes.byref(d), torch.cuda.current_stream().cuda_stream))

        once = os.environ.get("ATTN_PROBE_ONCE")       # "<shape substring>": one fwd + bwd of that shape (for ncu)
        if once is not None:
            if once in name:
                fwd(); bwd()
                torch.cuda.synchronize()
            continue
        tf, tb = timeit(fwd), timeit(bwd)
        lib.gtos_debug_attn_trace(ctypes.cast(buf, ctypes.c_void_p), 1)
        fwd(); bwd()
        torch.cuda.synchronize()
        lib.gtos_debug_attn_trace(ctypes.cast(buf, ctypes.c_void_p), 0)
        tr = np.frombuffer(buf, dtype=np.uint64).reshape(3, 16).astype(np.int64)
        ph = [" ".join(str(int(v - tr[kk, 0])) for v in tr[kk, 1:6]) for kk in range(3)]
        print(f"{name:22s}: fwd {tf:6.1f} us, bwd (q + kv) {tb:6.1f} us | CTA-0 cycles fwd [{ph[0]}] bwd_q [{ph[1]}] bwd_kv [{ph[2]}]")


if __name__ == "__main__":
    main()
