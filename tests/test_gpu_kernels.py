"""-m gpu: every C-ABI kernel of libgtos_b200.so against a plain PyTorch fp32 statement of the same
op on the same seeded inputs (bf16-rounded where the kernel consumes bf16 operands, so the only
difference is accumulation order).  Tolerances are stated per test."""
import ctypes as C

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

SEED = 19940117


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    from gtos_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.gtos_device_check(), "device_check")
    torch.manual_seed(SEED)
    return torch.device("cuda:0")


def bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 64, 128), (300, 512, 512), (2624, 1536, 512),
                                   (77, 130, 104), (1000, 768, 256), (5, 16, 8)])
def test_gemm_tn(dev, M, N, K):
    from gtos_b200 import ops
    A = bf(torch.randn(M, K, device=dev))
    B = bf(torch.randn(N, K, device=dev))
    bias = torch.randn(N, device=dev)
    ref = A.float() @ B.float().t() + bias
    out, outb = ops.gemm_tn(A, B, N, bias=bias, f32=True, bf16=True)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-5 * max(1, K ** 0.5)
    assert rel_err(outb[:, :N].float(), ref) < 1e-2
    out2, _ = ops.gemm_tn(A, B, N, bias=None, relu=True, out=out.clone(), accumulate=True)
    assert rel_err(out2, ref + torch.relu(ref - bias)) < 1e-4
    # single-output calls take the TMA-store epilogues (fp32 tiles / bf16 tiles); a bias that is only 4-byte aligned
    # takes the scalar bias path
    o32, _ = ops.gemm_tn(A, B, N, bias=bias)
    assert rel_err(o32, ref) < 2e-5 * max(1, K ** 0.5)
    _, o16 = ops.gemm_tn(A, B, N, bias=bias, f32=False, bf16=True, relu=True)
    assert rel_err(o16[:, :N].float(), torch.relu(ref)) < 1e-2
    if o16.shape[1] > N:
        assert float(o16[:, N:].float().abs().max()) == 0.0          # padding columns stay zero
    bias_u = torch.randn(N + 1, device=dev)[1:]
    ref_u = A.float() @ B.float().t() + bias_u
    o32u, _ = ops.gemm_tn(A, B, N, bias=bias_u)
    _, o16u = ops.gemm_tn(A, B, N, bias=bias_u, f32=False, bf16=True)
    assert rel_err(o32u, ref_u) < 2e-5 * max(1, K ** 0.5) and rel_err(o16u[:, :N].float(), ref_u) < 1e-2


@pytest.mark.parametrize("M,N,K", [(2624, 512, 512), (300, 304, 104), (129, 1024, 1536), (4000, 96, 64)])
@pytest.mark.parametrize("bn", ["64", "128", "256"])
def test_gemm_tn_every_tile_width_and_cta_pairs(dev, M, N, K, bn, monkeypatch):
    """the plain GEMM at a forced tile width, alone and as a CTA pair (cta_group::2: each CTA stages its own 128 rows of A
    and half of the weight rows; odd numbers of row blocks leave the last pair a zero-filled block).  The pair path only
    engages by itself for many-wave products (the GRU layer GEMMs, M ~ 10^5), so it is forced here: every epilogue (fp32 TMA
    store, bf16 TMA store, dual output, addend) must give bit-identical results to the default launch."""
    from gtos_b200 import _lib, ops
    A = bf(torch.randn(M, K, device=dev))
    B = bf(torch.randn(N, K, device=dev))
    bias = torch.randn(N, device=dev)
    add = torch.randn(M, N, device=dev)

    def run_all():
        o32, _ = ops.gemm_tn(A, B, N, bias=bias)
        _, o16 = ops.gemm_tn(A, B, N, bias=bias, f32=False, bf16=True, relu=True)
        d32, d16 = ops.gemm_tn(A, B, N, bias=bias, f32=True, bf16=True)
        oadd = torch.empty(M, N, device=dev)
        if N % 4 == 0:
            _lib.check(_lib.load().gtos_gemm_tn_add(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), None, add.data_ptr(), N,
                                                    oadd.data_ptr(), N, M, N, K, torch.cuda.current_stream().cuda_stream),
                       "gemm_tn_add")
        else:
            oadd.zero_()
        torch.cuda.synchronize()
        return o32, o16, d32, d16, oadd

    ref = run_all()
    assert rel_err(ref[0], A.float() @ B.float().t() + bias) < 2e-5 * max(1, K ** 0.5)
    monkeypatch.setenv("GTOS_FORCE_BN", bn)
    for cg in ("1", "2"):
        monkeypatch.setenv("GTOS_FORCE_CG", cg)
        got = run_all()
        for i, (a, b) in enumerate(zip(got, ref)):
            assert torch.equal(a, b), f"bn={bn} cg={cg} output {i}: max abs diff {float((a.float() - b.float()).abs().max()):.3e}"


@pytest.mark.parametrize("Kd,M,N", [(64, 128, 256), (128, 128, 64), (2624, 1536, 512), (1000, 512, 1024),
                                    (333, 200, 104), (4096, 768, 256)])
def test_gemm_nn(dev, Kd, M, N):
    from gtos_b200 import ops
    A = bf(torch.randn(Kd, (M + 7) // 8 * 8, device=dev))
    B = bf(torch.randn(Kd, (N + 7) // 8 * 8, device=dev))
    ref = A.float()[:, :M].t() @ B.float()[:, :N]
    out = ops.gemm_nn(A, B, M, N)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-5 * max(1, Kd ** 0.5)


def _rel_inputs(dev, N, B, D, H, wscale):
    x_q = bf(torch.randn(N * B, 2 * D, device=dev))                 # projected [q | k], staged as bf16
    rel = torch.randn(N, N, B, D, device=dev) * 0.5
    Wr = torch.randn(2 * D, D, device=dev) * wscale
    return x_q, rel, Wr


def _rel_scores_ref(qkv, relb, Wr, N, B, D, H):
    """scores[b,h,j,i] = hd^-1/2 < q_i + Wa r[j,i] , k_j + Wb r[j,i] >  from the bf16-rounded operands"""
    hd = D // H
    q = qkv.float()[:, :D].reshape(N, B, H, hd)
    k = qkv.float()[:, D:2 * D].reshape(N, B, H, hd)
    r = relb.float()
    W = bf(Wr).float()
    ra = (r @ W[:D].t()).view(N, N, B, H, hd)                        # [j,i,b,h,d]
    rb = (r @ W[D:].t()).view(N, N, B, H, hd)
    s = ((q.unsqueeze(0) + ra) * (k.unsqueeze(1) + rb)).sum(-1) * hd ** -0.5   # [j,i,b,h]
    return s.permute(2, 3, 0, 1).contiguous()                        # [b,h,j,i]


@pytest.mark.parametrize("N,B,D,H", [(17, 8, 128, 8), (41, 4, 512, 8), (9, 3, 256, 4), (130, 2, 128, 2)])
def test_rel_score(dev, N, B, D, H):
    from gtos_b200 import _lib, ops
    qkv, rel, Wr = _rel_inputs(dev, N, B, D, H, 0.05)
    relb = ops.relation_to_bf16(rel)
    Wperm, _ = ops.weight_prep(Wr, rel_heads=H)
    scores = torch.full((B, H, N, N), float("nan"), device=dev)
    _lib.check(_lib.load().gtos_rel_score(relb.data_ptr(), Wperm.data_ptr(), qkv.data_ptr(), qkv.data_ptr() + 2 * D,
                                          2 * D, scores.data_ptr(), N, B, D, H,
                                          torch.cuda.current_stream().cuda_stream), "rel_score")
    torch.cuda.synchronize()
    ref = _rel_scores_ref(qkv, relb, Wr, N, B, D, H)
    assert torch.isfinite(scores).all()
    assert rel_err(scores, ref) < 1e-4


@pytest.mark.parametrize("N,B,D,H", [(17, 8, 128, 8), (41, 4, 512, 8)])
def test_rel_backward_pieces(dev, N, B, D, H):
    """G, d_relation, dW_rel, dq, dk from the tcgen05 kernels vs autograd of the fp32 formula."""
    from gtos_b200 import _lib, ops
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    qkv, rel, Wr = _rel_inputs(dev, N, B, D, H, 0.05)
    relb = ops.relation_to_bf16(rel)
    Wperm, WpermT = ops.weight_prep(Wr, rel_heads=H)
    ds = torch.randn(B, H, N, N, device=dev)
    # reference through autograd on the bf16-rounded operands
    r32 = relb.float().requires_grad_()
    W32 = bf(Wr).float().requires_grad_()
    q32 = qkv.float().requires_grad_()
    hd = D // H
    q = q32[:, :D].reshape(N, B, H, hd)
    k = q32[:, D:2 * D].reshape(N, B, H, hd)
    ra = (r32 @ W32[:D].t()).view(N, N, B, H, hd)
    rb = (r32 @ W32[D:].t()).view(N, N, B, H, hd)
    s = (((q.unsqueeze(0) + ra) * (k.unsqueeze(1) + rb)).sum(-1) * hd ** -0.5).permute(2, 3, 0, 1)
    (s * ds).sum().backward()
    tiles = ops.rel_tiling(N, B, D, H)["tiles"]
    G = torch.empty(tiles * 128, 2 * D, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.gtos_rel_grad(relb.data_ptr(), Wperm.data_ptr(), qkv.data_ptr(), qkv.data_ptr() + 2 * D, 2 * D,
                                 ds.data_ptr(), G.data_ptr(), N, B, D, H, st), "rel_grad")
    dqkv = torch.zeros(N * B, 3 * D, device=dev)
    _lib.check(lib.gtos_rel_dqk(G.data_ptr(), dqkv.data_ptr(), dqkv.data_ptr() + 4 * D, 3 * D, None, None, N, B, D, H, st), "dqk")
    drel = torch.full((N, N, B, D), float("nan"), device=dev)
    _lib.check(lib.gtos_rel_drel(G.data_ptr(), WpermT.data_ptr(), drel.data_ptr(), 0, N, B, D, H, st), "drel")
    ws_n = lib.gtos_rel_dw_workspace(N, B, D, H)
    ws = torch.empty(max(ws_n, 1), device=dev)
    dW = torch.full((2 * D, D), float("nan"), device=dev)
    _lib.check(lib.gtos_rel_dw(G.data_ptr(), relb.data_ptr(), dW.data_ptr(), ws.data_ptr(), ws_n, N, B, D, H, st), "dw")
    torch.cuda.synchronize()
    assert torch.isfinite(G.float()).all()
    # G is rounded to bf16 once, so downstream tolerances are bf16-level (1e-2 relative)
    assert rel_err(dqkv[:, :D], q32.grad[:, :D]) < 1e-2
    assert rel_err(dqkv[:, D:2 * D], q32.grad[:, D:2 * D]) < 1e-2
    # optional bf16 operand copies: the same sums, rounded once
    dqkv2 = torch.zeros(N * B, 3 * D, device=dev)
    dqkv_b = torch.zeros(N * B, 3 * D, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.gtos_rel_dqk(G.data_ptr(), dqkv2.data_ptr(), dqkv2.data_ptr() + 4 * D, 3 * D, dqkv_b.data_ptr(),
                                dqkv_b.data_ptr() + 2 * D, N, B, D, H, st), "dqk(bf16)")
    torch.cuda.synchronize()
    assert torch.equal(dqkv2, dqkv)
    assert torch.equal(dqkv_b[:, :2 * D], bf(dqkv[:, :2 * D])) and float(dqkv_b[:, 2 * D:].float().abs().max()) == 0.0
    assert rel_err(drel, r32.grad) < 1e-2
    assert rel_err(dW, W32.grad) < 1e-2


def _attn_ref(q, k, v, scale, key_pad, attn_mask, H):
    T, B, D = q.shape
    S = k.shape[0]
    hd = D // H
    qh, kh, vh = (t.view(-1, B, H, hd) for t in (q, k, v))
    s = torch.einsum("tbhd,sbhd->bhts", qh, kh) * scale
    if attn_mask is not None:
        s = s.masked_fill(attn_mask.bool()[None, None], float("-inf"))
    if key_pad is not None:
        s = s.masked_fill(key_pad.bool().t()[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    return torch.einsum("bhts,sbhd->tbhd", p, vh).reshape(T, B, D), p


@pytest.mark.parametrize("T,S,B,H,hd,causal", [(6, 8, 3, 4, 8, False), (40, 40, 16, 8, 64, True), (33, 70, 5, 1, 512, False),
                                                (1, 40, 64, 8, 64, False), (1, 40, 2048, 8, 64, False)])
def test_attention_core(dev, T, S, B, H, hd, causal):
    from gtos_b200 import _lib, ops
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    D = H * hd
    q = torch.randn(T, B, D, device=dev, requires_grad=True)
    k = torch.randn(S, B, D, device=dev, requires_grad=True)
    v = torch.randn(S, B, D, device=dev, requires_grad=True)
    lens = torch.randint(max(1, S // 2), S + 1, (B,), device=dev)
    key_pad = (torch.arange(S, device=dev).unsqueeze(1) >= lens.unsqueeze(0))
    am = torch.ones(T, S, dtype=torch.bool, device=dev).triu_(1) if causal else None
    if causal:
        key_pad = None
    scale = hd ** -0.5
    ref, pref = _attn_ref(q, k, v, scale, key_pad, am, H)
    dout = torch.randn(T, B, D, device=dev)          # NOT randn_like: ref may be a strided view
    dw = torch.randn(B, H, T, S, device=dev)
    (ref * dout).sum().backward(retain_graph=True)
    gq, gk, gv = q.grad.clone(), k.grad.clone(), v.grad.clone()
    probs = torch.empty(B, H, T, S, device=dev)
    out = torch.empty(T * B, D, device=dev)
    d = ops._attn_desc(T, S, B, H, hd)
    d.q, d.ldq, d.k, d.ldk, d.v, d.ldv = q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D
    d.scale, d.p_drop = scale, 0.0
    d.key_pad = ops.as_u8(key_pad).data_ptr() if key_pad is not None else None
    d.attn_mask = ops.as_u8(am).data_ptr() if am is not None else None
    d.probs, d.out, d.ldo = probs.data_ptr(), out.data_ptr(), D
    _lib.check(lib.gtos_attn_fwd(C.byref(d), st), "attn_fwd")
    torch.cuda.synchronize()
    # the three products run on bf16 tensor-core tiles (fp32 accumulate): bf16-level tolerances
    assert rel_err(probs, pref) < 1e-2
    assert rel_err(out.view(T, B, D), ref) < 1e-2
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ds = torch.empty(B, H, T, S, device=dev)
    d.dout, d.lddo = dout.data_ptr(), D
    d.dscores_ts = ds.data_ptr()
    d.dq, d.lddq, d.dk, d.lddk, d.dv, d.lddv = dq.data_ptr(), D, dk.data_ptr(), D, dv.data_ptr(), D
    _lib.check(lib.gtos_attn_bwd(C.byref(d), st), "attn_bwd")
    torch.cuda.synchronize()
    assert rel_err(dq, gq) < 2e-2 and rel_err(dk, gk) < 2e-2 and rel_err(dv, gv) < 2e-2
    # optional bf16 operand copies of dq / dk / dv: bit-identical fp32 outputs, copies = those values rounded once
    dq2, dk2, dv2 = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    dqb, dkb, dvb = (torch.empty(t.shape, dtype=torch.bfloat16, device=dev) for t in (q, k, v))
    d.dq, d.dk, d.dv = dq2.data_ptr(), dk2.data_ptr(), dv2.data_ptr()
    d.dq_bf16, d.dk_bf16, d.dv_bf16 = dqb.data_ptr(), dkb.data_ptr(), dvb.data_ptr()
    _lib.check(lib.gtos_attn_bwd(C.byref(d), st), "attn_bwd(bf16 copies)")
    torch.cuda.synchronize()
    assert torch.equal(dq2, dq) and torch.equal(dk2, dk) and torch.equal(dv2, dv)
    assert torch.equal(dqb, bf(dq)) and torch.equal(dkb, bf(dk)) and torch.equal(dvb, bf(dv))
    d.dq, d.dk, d.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    d.dq_bf16 = d.dk_bf16 = d.dv_bf16 = None
    # extra gradient flowing into the returned weights
    q.grad = k.grad = v.grad = None
    ((ref * dout).sum() + (pref * dw).sum()).backward()
    d.dprobs_extra = dw.data_ptr()
    _lib.check(lib.gtos_attn_bwd(C.byref(d), st), "attn_bwd")
    torch.cuda.synchronize()
    assert rel_err(dq, q.grad) < 2e-2 and rel_err(dk, k.grad) < 2e-2 and rel_err(dv, v.grad) < 2e-2


def test_add_ln_and_ffn(dev):
    from gtos_b200 import ops
    M, D, Fd = 300, 128, 256
    x = torch.randn(M, D, device=dev, requires_grad=True)
    res = torch.randn(M, D, device=dev, requires_grad=True)
    g = (1 + 0.1 * torch.randn(D, device=dev)).requires_grad_()
    b = (0.1 * torch.randn(D, device=dev)).requires_grad_()
    y, yb = ops.add_layer_norm(x, res, g, b, 0.0)
    ref = torch.nn.functional.layer_norm(x + res, (D,), g, b)
    assert rel_err(y, ref) < 1e-5 and rel_err(yb.float(), ref) < 1e-2
    w = torch.randn_like(y)
    gs = torch.autograd.grad((y * w).sum(), [x, res, g, b])
    rs = torch.autograd.grad((ref * w).sum(), [x, res, g, b])
    for a, r in zip(gs, rs):
        assert rel_err(a, r) < 1e-4
    W1 = (torch.randn(Fd, D, device=dev) * 0.1).requires_grad_()
    b1 = (torch.randn(Fd, device=dev) * 0.1).requires_grad_()
    W2 = (torch.randn(D, Fd, device=dev) * 0.1).requires_grad_()
    b2 = (torch.randn(D, device=dev) * 0.1).requires_grad_()
    h = ops.ffn(x, None, W1, b1, W2, b2, 0.0)
    href = torch.relu(x @ W1.t() + b1) @ W2.t() + b2
    assert rel_err(h, href) < 1e-2
    # gradients against the same-rounding reference (bf16 operands): otherwise ReLU units whose
    # pre-activation is within bf16 noise of zero flip and dominate the max-norm error
    xr, W1r, W2r = (bf(t.detach()).float().requires_grad_() for t in (x, W1, W2))
    b1r, b2r = b1.detach().clone().requires_grad_(), b2.detach().clone().requires_grad_()
    hr = torch.relu(xr @ W1r.t() + b1r) @ W2r.t() + b2r
    gs = torch.autograd.grad((h * w).sum(), [x, W1, b1, W2, b2])
    rs = torch.autograd.grad((hr * w).sum(), [xr, W1r, b1r, W2r, b2r])
    for a, r in zip(gs, rs):
        assert rel_err(a, r) < 1e-2


def test_ln_param_grad_alone_and_tagged_gradient(dev):
    """gtos_ln_param_grad (the LayerNorm weight / bias gradients as their own entry point) vs torch, and the bf16
    operand copy the LayerNorm backward attaches to its result: exactly bf16(dx), dropped when the tensor is modified."""
    from gtos_b200 import _lib, ops
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    rows, D = 777, 256
    x = torch.randn(rows, D, device=dev)
    g = (1 + 0.1 * torch.randn(D, device=dev)).requires_grad_()
    b = (0.1 * torch.randn(D, device=dev)).requires_grad_()
    dy = torch.randn(rows, D, device=dev)
    mean, var = x.mean(1), x.var(1, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    dgamma, dbeta = torch.full((D,), float("nan"), device=dev), torch.full((D,), float("nan"), device=dev)
    _lib.check(lib.gtos_ln_param_grad(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dgamma.data_ptr(),
                                      dbeta.data_ptr(), rows, D, st), "ln_param_grad")
    ref = torch.nn.functional.layer_norm(x, (D,), g, b)
    rg, rb = torch.autograd.grad((ref * dy).sum(), [g, b])
    torch.cuda.synchronize()
    assert rel_err(dgamma, rg) < 1e-4 and rel_err(dbeta, rb) < 1e-4
    # tagged gradient
    xi = x.clone().requires_grad_()
    res = torch.randn(rows, D, device=dev, requires_grad=True)
    seen = {}

    class Probe(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            return t.clone()

        @staticmethod
        def backward(ctx, gr):
            seen["tag"] = getattr(gr, "_gtos_bf16", None)
            seen["grad"] = gr
            return gr

    y, _ = ops.add_layer_norm(Probe.apply(xi), res, g, b, 0.0)
    (y * dy).sum().backward()
    torch.cuda.synchronize()
    assert seen["tag"] is not None, "the LayerNorm backward did not tag its result"
    copy, ver = seen["tag"]
    assert ver == seen["grad"]._version and torch.equal(copy.view_as(seen["grad"]), bf(seen["grad"]))
    d2 = seen["grad"].view(rows, D)
    for k in ops.stats:
        ops.stats[k] = 0
    ops.grad_operand(seen["grad"], d2)[2].join()
    seen["grad"].add_(1.0)                                  # a modified tensor must not use the stale copy
    ops.grad_operand(seen["grad"], d2)[2].join()
    torch.cuda.synchronize()
    assert ops.stats == {"grad_operand_tagged": 1 if ops._side_enabled and ops._grad_tags else 0,
                         "grad_operand_cast": 1 if ops._side_enabled and ops._grad_tags else 2}


def test_dropout_statistics(dev):
    """p > 0 can only be tested statistically: keep rate, 1/(1-p) scaling, determinism per seed."""
    from gtos_b200 import ops
    x = torch.ones(1 << 20, device=dev)
    seed = ops.rng_state(dev)
    off = ops.new_seed_off()
    y = ops.dropout_f32(x, 0.2, seed, off)
    y2 = ops.dropout_f32(x, 0.2, seed, off)
    keep = (y != 0).float().mean().item()
    assert abs(keep - 0.8) < 5e-3
    assert torch.equal(y, y2)
    assert abs(y.max().item() - 1.25) < 1e-6
    y3 = ops.dropout_f32(x, 0.2, seed, ops.new_seed_off())
    assert not torch.equal(y, y3)


def test_bank_gather_and_scatter(dev):
    from gtos_b200 import ops
    R, D, N, B = 300, 128, 9, 4
    bank = torch.randn(R, D, device=dev, requires_grad=True)
    idx = torch.randint(0, R, (N, N, B), device=dev)
    idx[0, :, :] = 2                                     # a heavily repeated row (contended reductions)
    rel = ops.bank_gather(bank, idx)
    ref = bank.index_select(0, idx.reshape(-1)).view(N, N, B, D)
    assert torch.equal(rel, ref)
    assert torch.equal(ops.staged_relation_bf16(rel), ref.to(torch.bfloat16))
    w = torch.randn_like(ref)
    (g,) = torch.autograd.grad((rel * w).sum(), bank)
    (gr,) = torch.autograd.grad((ref * w).sum(), bank)
    assert rel_err(g, gr) < 1e-5


@pytest.mark.parametrize("sorted_bwd", [True, False])
@pytest.mark.parametrize("R,D,N,B", [(300, 128, 9, 4), (5000, 512, 41, 6), (50, 100, 7, 3)])
def test_bank_tensor_index_select_is_the_callers_gather(dev, R, D, N, B, sorted_bwd, monkeypatch):
    """RelationEncoder returns a BankTensor: the reference caller's own line (generator.py:79)
    `relation.index_select(0, idx.view(-1)).view(*idx.size(), -1)` then runs gtos_bank_gather - bit-identical values, the bf16
    operand copy riding along through .view - and its backward is the sorted segmented sum (or the plain scatter),
    equal to ATen's index_add_ up to fp32 summation order.  Hot rows, unused rows and a width that is not a multiple of
    128 are covered."""
    from gtos_b200 import ops
    monkeypatch.setattr(ops, "_bank_sorted_bwd", sorted_bwd)
    g = torch.Generator().manual_seed(SEED + R)
    plain = torch.randn(R, D, generator=g).to(dev).requires_grad_()
    idx = torch.randint(0, R // 2, (N, N, B), generator=g)
    idx[torch.rand(N, N, B, generator=g) < 0.4] = 5                        # a hot row (the <TL> path)
    idx[0] = 2
    idx = idx.to(dev)
    bank = ops.as_bank_tensor(plain * 1.0)
    assert isinstance(bank, ops.BankTensor)
    rel = bank.index_select(0, idx.view(-1)).view(*idx.size(), -1)          # the caller's line, verbatim
    ref = (plain * 1.0).as_subclass(torch.Tensor).index_select(0, idx.view(-1)).view(*idx.size(), -1)
    assert torch.equal(rel.as_subclass(torch.Tensor), ref)
    relb = ops.staged_relation_bf16(rel)
    assert relb is not None and torch.equal(relb, ref.to(torch.bfloat16))
    w = torch.randn(ref.shape, generator=g).to(dev)
    (ga,) = torch.autograd.grad((rel * w).sum(), plain)
    (gb,) = torch.autograd.grad((ref * w).sum(), plain)
    assert rel_err(ga, gb) < 1e-5
    assert float(ga[R // 2:].abs().max()) == 0.0                           # rows no pair uses
    # translator/generator.py:73 spelling, and an in-place write invalidates the bf16 tag
    rel2 = bank.index_select(0, idx.contiguous().view(-1)).contiguous().view(*idx.size(), -1)
    assert ops.staged_relation_bf16(rel2) is not None
    rel2.add_(1.0)
    assert ops.staged_relation_bf16(rel2) is None
    # opt-in provenance: the dense tensor is recognised as bank[idx] while nobody has written to it
    monkeypatch.setattr(ops, "_rel_provenance", True)
    rel3 = bank.index_select(0, idx.view(-1)).view(*idx.size(), -1)
    src = ops.factorised_source(rel3) if D % 128 == 0 else None
    if D % 128 == 0:
        assert src is not None and src.bank.data_ptr() == bank.data_ptr() and torch.equal(src.idx, idx)
    rel3.mul_(2.0)
    assert ops.factorised_source(rel3) is None
    monkeypatch.setattr(ops, "_rel_provenance", False)
    assert ops.factorised_source(bank.index_select(0, idx.view(-1)).view(*idx.size(), -1)) is None
    # anything else behaves like a plain tensor
    assert type(bank * 2) is torch.Tensor and torch.equal(bank[idx[..., 0]].as_subclass(torch.Tensor), plain.detach()[idx[..., 0]])


@pytest.mark.parametrize("N,B,D,H,R", [(9, 4, 128, 8, 300), (41, 6, 512, 8, 5000), (17, 3, 128, 4, 40)])
def test_rel_segsum_and_bank_weight_grad(dev, N, B, D, H, R):
    """§8 f-0 backward pieces: keys of the tile-major G rows, segmented sum over the row-sorted pair list
    (S_r = sum of the G rows whose pair uses bank row r) and dW = S^T bank in reference row order, against
    index_add / matmul in fp32 on the same bf16 inputs.  Hot rows (thousands of pairs), rows without pairs and
    tile padding rows are all present."""
    from gtos_b200 import _lib, ops
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cpu").manual_seed(SEED)
    idx = torch.randint(0, R // 2, (N, N, B), generator=g)              # rows >= R/2 never used
    idx[torch.rand(N, N, B, generator=g) < 0.3] = 5                       # one very hot row
    idx[0] = 2
    idx = idx.to(dev)
    til = ops.rel_tiling(N, B, D, H)
    rows = til["tiles"] * 128
    raw = torch.empty(rows, dtype=torch.int32, device=dev)
    _lib.check(lib.gtos_rel_pair_keys(idx.data_ptr(), N, B, D, H, R, raw.data_ptr(), st), "pair_keys")
    assert int((raw < R).sum()) == N * N * B
    # the key of G row (tile, jj*bi+ii) is idx[j, i, b]: rebuild the expected map from the tiling
    bi, bj, ni, nj = til["bi"], til["bj"], til["ni_blk"], til["nj_blk"]
    gidx = torch.arange(rows, device=dev)
    tile, r = gidx // 128, gidx % 128
    ib, jb, b = tile % ni, (tile // ni) % nj, tile // (ni * nj)
    jj, ii = r // bi, r % bi
    j, i = jb * bj + jj, ib * bi + ii
    valid = (jj < bj) & (i < N) & (j < N)
    exp = torch.full((rows,), R, dtype=torch.int64, device=dev)
    exp[valid] = idx[j[valid], i[valid], b[valid]]
    assert torch.equal(raw.long(), exp)
    keys, order = torch.sort(raw)
    order = order.to(torch.int32)
    G = (torch.randn(rows, 2 * D, device=dev) * 0.3).to(torch.bfloat16)
    S = torch.zeros(R, 2 * D, dtype=torch.bfloat16, device=dev)
    spill = torch.full((R, 2 * D), float("nan"), device=dev)             # must be zeroed by the kernel where used
    _lib.check(lib.gtos_rel_segsum(G.data_ptr(), order.data_ptr(), keys.data_ptr(), N * N * B, 2 * D, S.data_ptr(),
                                   2 * D, spill.data_ptr(), st), "segsum")
    ref = torch.zeros(R + 1, 2 * D, device=dev).index_add_(0, exp, G.float())[:R]
    assert torch.isfinite(S.float()).all()
    assert rel_err(S.float(), ref) < 2 ** -8                               # one bf16 rounding of an fp32 sum
    assert (S[R // 2:] == 0).all()
    bank = torch.randn(R, D, device=dev).to(torch.bfloat16)
    dW = torch.empty(2 * D, D, device=dev)
    _lib.check(lib.gtos_rel_dw_bank(S.data_ptr(), 2 * D, bank.data_ptr(), dW.data_ptr(), R, D, H, st), "dw_bank")
    hd = D // H
    perm = torch.tensor([(p // (2 * hd)) * hd + (p % (2 * hd)) if (p % (2 * hd)) < hd
                         else D + (p // (2 * hd)) * hd + (p % (2 * hd)) - hd for p in range(2 * D)], device=dev)
    dref = torch.zeros(2 * D, D, device=dev)
    dref[perm] = S.float().t() @ bank.float()
    assert rel_err(dW, dref) < 1e-4
