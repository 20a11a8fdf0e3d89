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


STEP_SHAPES = [  # every plain-GEMM shape of the cfg2 step (M = rows of the activation, N, K)
    (2624, 1024, 512, "enc QK proj / fc1 / dh"), (2624, 512, 512, "enc V / out_proj / datt"), (2624, 512, 1024, "enc fc2 / FFN dx"),
    (2624, 512, 1536, "enc in_proj dx"), (3840, 1536, 512, "dec QKV"), (3840, 512, 512, "dec q / out_proj / datt"),
    (2560, 1024, 512, "dec memory K,V"), (3840, 1024, 512, "dec fc1 / dh"), (3840, 512, 1024, "dec fc2 / FFN dx"),
    (3840, 512, 1536, "dec self in_proj dx"), (2560, 512, 1024, "dec memory dx"), (3840, 304, 512, "transfer"),
    (3840, 10000, 304, "vocabulary projection"), (3840, 304, 10000, "d logits x W"), (3840, 512, 304, "transfer dx"),
    (24659, 512, 512, "bank out_proj"), (24659, 256, 768, "GRU dh GEMM"), (98636, 512, 1536, "GRU layer-1 dx"),
    (98636, 104, 1536, "GRU layer-0 dx"), (24659, 512, 512, "d finals"),
]


def sweep():
    """graph back-to-back time of every step shape at every tile width (GTOS_FORCE_BN) next to the chooser's own pick"""
    dev = torch.device("cuda:0")
    n = 40
    for M, N, K, what in STEP_SHAPES:
        A = torch.randn(M, K, device=dev).to(torch.bfloat16)
        B = torch.randn(N, K, device=dev).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev)
        res = {}
        os.environ.pop("GTOS_FORCE_BN", None)
        os.environ.pop("GTOS_FORCE_CG", None)
        ops.gemm_tn(A, B, N, bias=bias, out=out)
        ref = out.clone()
        for bn in ("auto", "64", "128", "256", "64x2", "128x2", "256x2"):
            os.environ.pop("GTOS_FORCE_CG", None)
            if bn == "auto":
                os.environ.pop("GTOS_FORCE_BN", None)
            else:
                os.environ["GTOS_FORCE_BN"] = bn.split("x")[0]
                if bn.endswith("x2"):
                    os.environ["GTOS_FORCE_CG"] = "2"          # CTA pair (cta_group::2)
            try:
                out.zero_()
                ops.gemm_tn(A, B, N, bias=bias, out=out)
                torch.cuda.synchronize()
                if not torch.equal(out, ref):
                    print(f"   !! {what} bn={bn}: result differs from the default tile, max abs {float((out - ref).abs().max()):.3e}")
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    ops.gemm_tn(A, B, N, bias=bias, out=out)
                    s.synchronize()
                    with torch.cuda.graph(g, stream=s):
                        for _ in range(n):
                            ops.gemm_tn(A, B, N, bias=bias, out=out)
                torch.cuda.synchronize()
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                res[bn] = e0.elapsed_time(e1) / n * 1e3
            except Exception as e:
                res[bn] = float("nan")
        os.environ.pop("GTOS_FORCE_BN", None)
        os.environ.pop("GTOS_FORCE_CG", None)
        best = min((v, k) for k, v in res.items() if k != "auto" and v == v)
        print(f"{what:28s} M={M:6d} N={N:6d} K={K:5d}: auto {res['auto']:7.1f} | 64: {res['64']:7.1f}  128: {res['128']:7.1f}  "
              f"256: {res['256']:7.1f} | pairs 64: {res.get('64x2', float('nan')):7.1f}  128: {res.get('128x2', float('nan')):7.1f}  "
              f"256: {res.get('256x2', float('nan')):7.1f} us  -> best {best[1]} ({100 * (res['auto'] - best[0]) / res['auto']:.0f}% under auto)")


NN_SHAPES = [  # weight-gradient GEMMs of the cfg2 step: dW [M, N] = dY[Kd, M]^T X[Kd, N]
    (512, 512, 2624, "enc out_proj / V"), (1536, 512, 2624, "enc in_proj"), (1024, 512, 2624, "enc fc1"), (512, 1024, 2624, "enc fc2"),
    (512, 512, 3840, "dec out_proj / q"), (1536, 512, 3840, "dec self in_proj"), (1024, 512, 2560, "dec memory K,V"),
    (1024, 512, 3840, "dec fc1"), (512, 1024, 3840, "dec fc2"), (304, 512, 3840, "transfer"), (10000, 304, 3840, "vocabulary"),
    (768, 104, 98636, "GRU w_ih l0"), (768, 512, 98636, "GRU w_ih l1"), (768, 256, 98636, "GRU w_hh"), (512, 512, 24659, "bank out_proj"),
]


def sweep_nn():
    dev = torch.device("cuda:0")
    n = 30
    for M, N, Kd, what in NN_SHAPES:
        A = torch.randn(Kd, (M + 7) // 8 * 8, device=dev).to(torch.bfloat16)
        B = torch.randn(Kd, (N + 7) // 8 * 8, device=dev).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev)
        res = {}
        for bn in ("auto", "64", "128", "256"):
            for sp in ("auto", "2", "4", "6", "9", "12", "18"):
                if (bn == "auto") != (sp == "auto"):
                    continue
                for k, v in (("GTOS_FORCE_NN_BN", bn), ("GTOS_FORCE_NN_SPLITS", sp)):
                    if v == "auto":
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
                try:
                    g = torch.cuda.CUDAGraph()
                    s = torch.cuda.Stream()
                    with torch.cuda.stream(s):
                        ops.gemm_nn(A, B, M, N, out=out)
                        s.synchronize()
                        with torch.cuda.graph(g, stream=s):
                            for _ in range(n):
                                ops.gemm_nn(A, B, M, N, out=out)
                    torch.cuda.synchronize()
                    g.replay()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    g.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    res[(bn, sp)] = e0.elapsed_time(e1) / n * 1e3
                except Exception:
                    pass
        for k in ("GTOS_FORCE_NN_BN", "GTOS_FORCE_NN_SPLITS"):
            os.environ.pop(k, None)
        auto = res.pop(("auto", "auto"), float("nan"))
        top = sorted((v, k) for k, v in res.items())[:4]
        print(f"{what:18s} M={M:5d} N={N:5d} Kd={Kd:6d}: auto {auto:6.1f} us | best " +
              ", ".join(f"bn{k[0]}/s{k[1]}: {v:5.1f}" for v, k in top))


def main():
    if "--sweep" in sys.argv:
        return sweep()
    if "--sweep-nn" in sys.argv:
        return sweep_nn()
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
