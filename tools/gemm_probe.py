#!/usr/bin/env python
"""Time the plain tcgen05 GEMM (gtos_gemm_tn / gtos_gemm_nn) at the hot path's small shapes with CUDA events,
back to back and with an L2 flush in between.  Usage: python tools/gemm_probe.py [--ncu]  (--ncu: one launch of each
shape between cudaProfilerStart/Stop for `ncu --profile-from-start off`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtos_b200 import _lib, ops          # noqa: E402

SHAPES = [  # (M, N, K, what)
    (2624, 512, 512, "encoder out_proj / V proj"),
    (2624, 1024, 512, "encoder QK proj / fc1"),
    (2624, 512, 1024, "encoder fc2"),
    (3840, 1536, 512, "decoder QKV"),
    (3840, 512, 512, "decoder out_proj"),
    (2048, 512, 512, "decode step proj (Hyp=2048)"),
    (3840, 10000, 304, "vocabulary projection"),
    (24659, 256, 768, "GRU dh GEMM"),
]


def main():
    ncu = "--ncu" in sys.argv
    dev = torch.device("cuda:0")
    lib = _lib.load()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for M, N, K, what in SHAPES:
        A = torch.randn(M, K, device=dev).to(torch.bfloat16)
        B = torch.randn(N, K, device=dev).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev)

        def run():
            ops.gemm_tn(A, B, N, bias=bias, out=out)

        for _ in range(5):
            run()
        torch.cuda.synchronize()
        if os.environ.get("GTOS_DBG") == "2":
            import ctypes
            import numpy as np
            run()
            torch.cuda.synchronize()
            buf = (ctypes.c_uint64 * (148 * 16))()
            _lib.check(lib.gtos_debug_read_trace(ctypes.cast(buf, ctypes.c_void_p), 148 * 16))
            tr = np.frombuffer(buf, dtype=np.uint64).reshape(148, 16).astype(np.int64)
            rel = tr[:, 1:16] - tr[:, :1]
            names = ["prologue", "pdl_wait", "tma_issued", "first_full", "mma_done", "tfull", "epi_done", "store_drained", "exit",
                     "c1_start", "c1_tmem", "c1_bias", "c1_waitread", "c1_bar1", "c1_end"]
            for cta in (0, 40, 83):
                print(f"   {what} cta {cta}: " + ", ".join(f"{n}={int(v)}" for n, v in zip(names, rel[cta])))
            continue
        if ncu:
            torch.cuda.cudart().cudaProfilerStart()
            run()
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        b2b = e0.elapsed_time(e1) / n * 1e3
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            run()
            s.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(n):
                    run()
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        gr = e0.elapsed_time(e1) / n * 1e3
        tot = 0.0
        for _ in range(5):
            flush.zero_()
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1) * 1e3
        fl = 2.0 * M * N * K
        print(f"{what:32s} M={M:6d} N={N:6d} K={K:5d}: eager b2b {b2b:6.1f} us, graph b2b {gr:6.1f} us ({fl / gr / 1e6:6.0f} TF/s), "
              f"cold single {tot / 5:6.1f} us")


if __name__ == "__main__":
    main()
