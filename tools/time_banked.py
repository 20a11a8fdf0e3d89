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
    # ---- forward half: projected bank + gather kernels vs the P-row tcgen05 kernels ----
    NB = N * B
    hd = D // H
    qk = (torch.randn(NB, 2 * D, device=dev)).to(torch.bfloat16)
    v = torch.randn(NB, D, device=dev)
    pad = (torch.arange(N, device=dev).unsqueeze(1) >= (g["node_counts"].to(dev) + 1).unsqueeze(0)).to(torch.uint8).contiguous()
    scores = torch.empty(B, H, N, N, device=dev)
    probs = torch.empty(B, H, N, N, device=dev)
    att = torch.empty(NB, D, device=dev)
    attb = torch.empty(NB, D, dtype=torch.bfloat16, device=dev)
    seed = ops.rng_state(dev)
    relb = br.relb
    res["--- forward half ---"] = 0.0
    res["rel_score (P-row tcgen05 GEMM + score epilogue)"] = t(lambda: _lib.check(lib.gtos_rel_score(
        relb.data_ptr(), Wperm.data_ptr(), qk.data_ptr(), qk.data_ptr() + 2 * D, 2 * D, scores.data_ptr(), N, B, D, H, st)))
    import ctypes as C
    d = ops._attn_desc(N, N, B, H, hd)
    d.v, d.ldv = v.data_ptr(), D
    d.scale, d.p_drop = 1.0, 0.2
    d.scores_jt, d.key_pad = scores.data_ptr(), pad.data_ptr()
    d.seed_ptr, d.seed_off = seed.data_ptr(), 12345
    d.probs = probs.data_ptr()
    d.out, d.ldo, d.out_bf16 = att.data_ptr(), D, attb.data_ptr()
    res["attn_fwd (mask+softmax+dropout+PV, encoder)"] = t(lambda: _lib.check(lib.gtos_attn_fwd(C.byref(d), st)))
    PB = torch.empty(R, 2 * D, dtype=torch.bfloat16, device=dev)
    res["bank projection gemm_tn [R,D]x[D,2D] -> bf16"] = t(lambda: ops.gemm_tn(br.bankb, Wperm, 2 * D, f32=False, bf16=True))
    _, PB = ops.gemm_tn(br.bankb, Wperm, 2 * D, f32=False, bf16=True)
    res["rel_attn_banked_fwd (gather+score+softmax+PV)"] = t(lambda: _lib.check(lib.gtos_rel_attn_banked_fwd(
        PB.data_ptr(), PB.stride(0), idx.data_ptr(), qk.data_ptr(), qk.data_ptr() + 2 * D, 2 * D, v.data_ptr(), D, pad.data_ptr(), None,
        0.2, seed.data_ptr(), 12345, probs.data_ptr(), None, att.data_ptr(), D, attb.data_ptr(), N, B, D, H, R, st)))
    ds = torch.randn(B, H, N, N, device=dev)
    res["rel_grad (P-row recompute GEMM + G epilogue)"] = t(lambda: _lib.check(lib.gtos_rel_grad(
        relb.data_ptr(), Wperm.data_ptr(), qk.data_ptr(), qk.data_ptr() + 2 * D, 2 * D, ds.data_ptr(), G.data_ptr(), N, B, D, H, st)))
    res["rel_grad_banked (gather + G rows)"] = t(lambda: _lib.check(lib.gtos_rel_grad_banked(
        PB.data_ptr(), PB.stride(0), idx.data_ptr(), qk.data_ptr(), qk.data_ptr() + 2 * D, 2 * D, ds.data_ptr(), G.data_ptr(),
        N, B, D, H, R, st)))
    res["bank bf16 gather (dense operand of rel_score)"] = t(lambda: (setattr(br, "_relb", None), br.relb))
    pairs = N * N * B
    print(f"gather traffic per launch: {pairs * 2 * D * 2 / 1e6:.0f} MB of PB rows (table {R * 2 * D * 2 / 1e6:.0f} MB); "
          f"G written: {pairs * 2 * D * 2 / 1e6:.0f} MB")
    for k, v_ in res.items():
        print(f"{k:52s} {v_:8.1f} us")


if __name__ == "__main__":
    main()
