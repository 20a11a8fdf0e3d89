#!/usr/bin/env python
"""CUDA-event timing of the bank-factorised backward pieces at config-2 size (run on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gtos_b200 import _lib, ops, synthetic  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = synthetic.make_graph_batch(64, 40, max_path_len=4)
    idx = g["relation"].to(dev)
    N, B, D, H = g["N"], 64, 512, 8
    R = g["relation_bank"].shape[1]
    st = torch.cuda.current_stream().cuda_stream
    til = ops.rel_tiling(N, B, D, H)
    rows = til["tiles"] * 128
    bank = torch.randn(R, D, device=dev)
    br = ops.BankedRelation(bank, idx)
    br.prepare(H)
    G = (torch.randn(rows, 2 * D, device=dev) * 0.1).to(torch.bfloat16)
    S = torch.zeros(R, 2 * D, dtype=torch.bfloat16, device=dev)
    spill = torch.empty(R, 2 * D, device=dev)
    dW = torch.empty(2 * D, D, device=dev)
    dbank = torch.empty(R, D, device=dev)
    W = torch.randn(2 * D, D, device=dev) * 0.02
    Wperm, WpermT = ops.weight_prep(W, rel_heads=H)
    drel = torch.empty(N, N, B, D, device=dev)
    cnt = torch.bincount(br.keys[: N * N * B].long(), minlength=R)
    print(f"R={R} P={N*N*B} rows={rows} max pairs/row={int(cnt.max())} rows>32: {int((cnt > 32).sum())} empty: {int((cnt == 0).sum())}")

    def t(fn, n=20):
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

    res = {}
    res["prepare(keys+sort)"] = t(lambda: (setattr(br, "_heads", None), br.prepare(H)))
    res["BankedRelation(cast+bf16 gather)"] = t(lambda: ops.BankedRelation(bank, idx))
    res["bank_gather(fp32+bf16)"] = t(lambda: ops.bank_gather(bank, idx))
    res["segsum(3 launches)"] = t(lambda: _lib.check(lib.gtos_rel_segsum(G.data_ptr(), br.order.data_ptr(), br.keys.data_ptr(),
                                                                          N * N * B, 2 * D, S.data_ptr(), 2 * D, spill.data_ptr(), st)))
    res["rel_dw_bank"] = t(lambda: _lib.check(lib.gtos_rel_dw_bank(S.data_ptr(), 2 * D, br.bankb.data_ptr(), dW.data_ptr(), R, D, H, st)))
    res["dbank gemm_tn"] = t(lambda: ops.gemm_tn(S, WpermT, D, out=dbank, accumulate=True, K=2 * D))
    res["dbank gemm_tn (no acc)"] = t(lambda: ops.gemm_tn(S, WpermT, D, out=dbank, accumulate=False, K=2 * D))
    Scat = torch.zeros(R, 8 * D, dtype=torch.bfloat16, device=dev)
    Wcat = torch.zeros(D, 8 * D, dtype=torch.bfloat16, device=dev)
    res["dbank = [S1..S4][W1;..;W4] (one GEMM)"] = t(lambda: ops.gemm_tn(Scat, Wcat, D))
    res["rel_drel (dense)"] = t(lambda: _lib.check(lib.gtos_rel_drel(G.data_ptr(), WpermT.data_ptr(), drel.data_ptr(), 1, N, B, D, H, st)))
    res["rel_dw (dense)"] = t(lambda: _lib.check(lib.gtos_rel_dw(G.data_ptr(), br.relb.data_ptr(), dW.data_ptr(), None, 0, N, B, D, H, st)))
    dbk = torch.empty(R, D, device=dev)
    res["bank_scatter_add (dense)"] = t(lambda: _lib.check(lib.gtos_bank_scatter_add(drel.data_ptr(), idx.data_ptr(), N * N * B, D,
                                                                                     dbk.data_ptr(), R, st)))
    for k, v in res.items():
        print(f"{k:36s} {v:8.1f} us")


if __name__ == "__main__":
    main()
