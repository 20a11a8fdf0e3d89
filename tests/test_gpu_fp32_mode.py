"""-m gpu: fp32 mode (ops.set_precision("fp32"), BASELINE north star "within 1e-3 fp32").

Every assertion here is against the TRUE fp32 oracle (oracle/gtos_oracle.py in its default "fp32" matmul mode - the
reference's own arithmetic, pinned by tests/golden), outputs AND gradients, at 1e-3: max-norm error relative to the largest
reference value, and relative L2.  Weights are inflated x2 / x3 like the bf16-mode tests (errors hide at the default
std 0.02).  No bf16-emulation oracle is involved anywhere in this file.

ReLU sub-gradients.  Outputs need no care.  Gradients do: relu'(0) is a jump, so two correct fp32 implementations whose
FFN pre-activations differ by 1e-5 of the layer's scale take different sub-gradients at the handful of units (about one in
10^5) that sit that close to zero, and one such unit moves its row of the input gradient by percents (first run of this
file: outputs at 1e-5, gradients at 2e-3 L2 / 1e-2 max-norm on exactly the D = 512, x3-weight cases, one row each).  The
tests therefore fix the CHOICE, not the arithmetic (class Kinks): the GPU run reports its activation pattern; the oracle
takes the GPU's choice for units with |pre-activation| < 1e-4 x the layer's largest one, its own everywhere else; the test
asserts that the two patterns agree on every unit outside that band and that the band holds < 0.1 % of the units.  Given
the same choice, every gradient has to agree at 1e-3.
"""
import pytest
import torch

from conftest import l2_err, rel_err
from oracle import gtos_oracle as O
from oracle import hotpath_oracle as HO
from test_gpu_modules import _V, boost, compare_grads, oracle_params, pad_mask

pytestmark = pytest.mark.gpu
TOL32 = 1e-3
SEED = 19940117


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    from gtos_b200 import _lib
    _lib.check(_lib.load().gtos_device_check(), "device_check")
    return torch.device("cuda:0")


class Kinks:
    """the ReLU sub-gradient protocol described in the module docstring"""
    BAND = 1e-4

    def __init__(self):
        self.masks, self.units, self.in_band, self.flipped, self.outside = [], 0, 0, 0, 0

    def gpu(self):
        return _KinkRecord(self)

    def oracle(self):
        return _KinkReplay(self)

    def reference(self, width):
        """the same protocol for the REAL reference modules (oracle/_ref): torch.nn.functional.relu is wrapped while the
        reference runs; calls whose last dimension is `width` (the FFN hidden size: graph_transformer.py:61,
        transformer.py:67) go through the hook, every other relu (the char-CNN front-end) is untouched"""
        return _KinkReplay(self, functional_width=width)

    def check(self):
        assert self.units > 0 and not self.masks, "the oracle consumed a different number of FFN calls than the GPU ran"
        assert self.outside == 0, f"{self.outside} ReLU units disagree OUTSIDE the rounding band"
        assert self.in_band <= max(3, 1e-3 * self.units), (self.in_band, self.units)


class _KinkRecord:
    def __init__(self, k):
        self.k = k

    def __enter__(self):
        from gtos_b200 import ops32
        ops32.relu_trace = self.k.masks

    def __exit__(self, *a):
        from gtos_b200 import ops32
        ops32.relu_trace = None


class _KinkReplay:
    def __init__(self, k, functional_width=None):
        self.k = k
        self.width = functional_width
        self._orig = None

    def _hook(self, pre):
        k = self.k
        m_gpu = k.masks.pop(0).view_as(pre)
        own = pre.detach() > 0
        band = pre.detach().abs() < k.BAND * pre.detach().abs().max()
        k.units += pre.numel()
        k.in_band += int(band.sum())
        k.flipped += int(((m_gpu != own) & band).sum())
        k.outside += int(((m_gpu != own) & ~band).sum())
        return pre * torch.where(band, m_gpu, own).to(pre.dtype)

    def __enter__(self):
        if self.width is None:
            O._RELU_HOOK = self._hook
            return
        import torch.nn.functional as F
        self._orig = F.relu

        def relu(x, inplace=False):
            if x.dim() == 3 and x.shape[-1] == self.width and self.k.masks:
                return self._hook(x)
            return self._orig(x, inplace=inplace)

        F.relu = relu

    def __exit__(self, *a):
        if self.width is None:
            O._RELU_HOOK = None
        else:
            import torch.nn.functional as F
            F.relu = self._orig


@pytest.fixture()
def fp32_mode():
    from gtos_b200 import ops
    with ops.precision_mode("fp32"):
        yield


def test_split_operand_gemms_match_float64(dev):
    """the K-tripled GEMMs (forward form, transposed-weight form, row-stacked weight-gradient form) against float64"""
    from gtos_b200 import ops, ops32
    gen = torch.Generator().manual_seed(SEED)
    for M, K, N in [(300, 100, 260), (2624, 512, 1536), (77, 300, 1000)]:
        x = (torch.randn(M, K, generator=gen) * 3).to(dev)
        W = torch.randn(N, K, generator=gen).to(dev)
        b = torch.randn(N, generator=gen).to(dev)
        dy = torch.randn(M, N, generator=gen).to(dev)
        y = ops32.mm3(x, W, b)
        ref = (x.double() @ W.double().t() + b.double())
        assert rel_err(y, ref) < 2e-5, (M, K, N, rel_err(y, ref))
        # plain bf16 operands for scale: three orders of magnitude worse
        yb, _ = ops.gemm_tn(ops.cast_bf16(x), ops.cast_bf16(W), N, bias=b)
        assert rel_err(yb, ref) > 20 * rel_err(y, ref)
        xs, dys = ops32.split3(x, 0), ops32.split3(dy, 1)
        dx, _ = ops.gemm_tn(dys, ops32.split3(W.t(), 0), K)
        assert rel_err(dx, dy.double() @ W.double()) < 2e-5
        dW = ops32._wgrad(dys, xs, N, K)
        assert rel_err(dW, dy.double().t() @ x.double()) < 2e-5
    # layout of the staged operand
    x = torch.randn(5, 12, generator=gen).to(dev)
    s0, s1 = ops32.split3(x, 0).float(), ops32.split3(x, 1).float()
    hi = x.to(torch.bfloat16).float()
    lo = (x - hi).to(torch.bfloat16).float()
    z = torch.zeros(5, 4, device=dev)
    assert torch.equal(s0, torch.cat([hi, z, lo, z, hi, z], 1)) and torch.equal(s1, torch.cat([hi, z, hi, z, lo, z], 1))
    assert (x - hi - lo).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()


@pytest.mark.parametrize("T,S,B,H,hd,causal", [(7, 9, 3, 4, 8, False), (30, 41, 5, 8, 64, False), (33, 33, 2, 8, 64, True),
                                               (60, 257, 2, 1, 128, False)])
def test_attention_core_three_pass_mode_vs_float64(dev, T, S, B, H, hd, causal):
    """gtos_attn_fwd / gtos_attn_bwd with desc.precise = 1 (decoder mode: scores from q, k) against float64 torch"""
    import ctypes as C
    from gtos_b200 import _lib, ops
    gen = torch.Generator().manual_seed(SEED + T)
    D = H * hd
    q, k, v, do = (torch.randn(n, B, D, generator=gen).to(dev) * s for n, s in ((T, 2.0), (S, 2.0), (S, 1.0), (T, 1.0)))
    lens = [S] + [int(x) for x in torch.randint(S // 2, S + 1, (B - 1,), generator=gen)]
    kp = pad_mask(lens, S).to(dev)
    am = torch.ones(T, S, dtype=torch.bool).triu_(1).to(dev) if causal else None
    probs = torch.empty(B, H, T, S, device=dev)
    out = torch.empty(T * B, D, device=dev)
    d = ops._attn_desc(T, S, B, H, hd)
    d.precise = 1
    d.q, d.ldq, d.k, d.ldk, d.v, d.ldv = q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D
    d.scale, d.p_drop = hd ** -0.5, 0.0
    d.key_pad = ops.as_u8(kp).data_ptr()
    d.attn_mask = ops.as_u8(am).data_ptr() if am is not None else None
    d.probs, d.out, d.ldo = probs.data_ptr(), out.data_ptr(), D
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.load().gtos_attn_fwd(C.byref(d), st), "attn_fwd")
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ds = torch.empty(B, H, T, S, device=dev)
    d.dout, d.lddo, d.dscores_ts = do.data_ptr(), D, ds.data_ptr()
    d.dq, d.lddq, d.dk, d.lddk, d.dv, d.lddv = dq.data_ptr(), D, dk.data_ptr(), D, dv.data_ptr(), D
    _lib.check(_lib.load().gtos_attn_bwd(C.byref(d), st), "attn_bwd")
    # float64 reference
    q6, k6, v6 = (t.double().requires_grad_() for t in (q, k, v))
    qh = q6.view(T, B, H, hd).permute(1, 2, 0, 3)
    kh = k6.view(S, B, H, hd).permute(1, 2, 0, 3)
    vh = v6.view(S, B, H, hd).permute(1, 2, 0, 3)
    w = qh @ kh.transpose(-1, -2) * hd ** -0.5
    w = w.masked_fill(kp.t()[:, None, None, :], float("-inf"))
    if am is not None:
        w = w.masked_fill(am[None, None], float("-inf"))
    pr = torch.softmax(w, -1)
    o = (pr @ vh).permute(2, 0, 1, 3).reshape(T * B, D)
    gq, gk, gv = torch.autograd.grad((o * do.double().view(T * B, D)).sum(), [q6, k6, v6])
    # split operands carry 16 significand bits: score errors of ~1e-5 absolute, amplified by p (1 - p) in the softmax
    assert rel_err(probs, pr) < 1e-4 and rel_err(out, o) < 1e-4, (rel_err(probs, pr), rel_err(out, o))
    for a, b_ in ((dq, gq), (dk, gk), (dv, gv)):
        assert rel_err(a, b_) < 2e-4 and l2_err(a, b_) < 2e-4, (rel_err(a, b_), l2_err(a, b_))


@pytest.mark.parametrize("N,B,D,H,F,L,wf", [(17, 8, 128, 8, 256, 2, 1.0), (17, 8, 128, 8, 256, 2, 3.0),
                                            (41, 6, 512, 8, 1024, 2, 2.0), (61, 3, 512, 8, 1024, 1, 3.0)])
def test_graph_transformer_fp32_mode_vs_fp32_oracle(dev, fp32_mode, N, B, D, H, F, L, wf):
    from gtos_b200.graph_transformer import GraphTransformer
    gen = torch.Generator().manual_seed(SEED)
    m = GraphTransformer(L, D, F, H, 0.0)
    boost(m, wf, gen)
    x = torch.randn(N, B, D, generator=gen)
    rel = torch.randn(N, N, B, D, generator=gen) * 0.5
    lens = [N] + [int(v) for v in torch.randint(N // 2, N + 1, (B - 1,), generator=gen)]
    mask = pad_mask(lens, N)
    wo = torch.randn(N, B, D, generator=gen)
    P = oracle_params(m)
    xc, rc = x.clone().requires_grad_(), rel.clone().requires_grad_()
    plain = O.graph_transformer(P, "", xc, rc, L, H, self_padding_mask=mask).detach()     # no protocol: plain relu
    m = m.to(dev)
    xg, rg = x.to(dev).requires_grad_(), rel.to(dev).requires_grad_()
    kinks = Kinks()
    with kinks.gpu():
        out = m(xg, rg, self_padding_mask=mask.to(dev))
    with kinks.oracle():
        ref = O.graph_transformer(P, "", xc, rc, L, H, self_padding_mask=mask)
    kinks.check()
    assert rel_err(out, plain) < TOL32 and rel_err(out, ref) < TOL32, (rel_err(out, plain), rel_err(out, ref))
    compare_grads(m, P, (out * wo.to(dev)).sum(), (ref * wo).sum(), [xg, rg], [xc, rc], tol=TOL32, tol_max=TOL32)
    with torch.no_grad():
        attn = m.get_attn_weights(xg, rg, self_padding_mask=mask.to(dev))
    aref = O.graph_transformer(P, "", xc, rc, L, H, self_padding_mask=mask, return_weights=True)
    assert attn.shape == aref.shape and rel_err(attn, aref) < TOL32


def test_graph_transformer_fp32_mode_factorised_relation_and_weights_grad(dev, fp32_mode):
    """a BankedRelation argument (gathered densely in this mode): gradient reaches the bank; plus the need_weights path
    of RelationMultiheadAttention with a gradient flowing into the returned weights"""
    from gtos_b200 import ops
    from gtos_b200.graph_transformer import GraphTransformer, RelationMultiheadAttention
    N, B, D, H, F, L, R = 17, 5, 128, 8, 256, 2, 150
    gen = torch.Generator().manual_seed(SEED + 7)
    m = GraphTransformer(L, D, F, H, 0.0)
    boost(m, 2.0, gen)
    x = torch.randn(N, B, D, generator=gen)
    bank = torch.randn(R, D, generator=gen) * 0.5
    idx = torch.randint(0, R, (N, N, B), generator=gen)
    mask = pad_mask([N, 9, 12, 15, 17], N)
    wo = torch.randn(N, B, D, generator=gen)
    P = oracle_params(m)
    xc, bc = x.clone().requires_grad_(), bank.clone().requires_grad_()
    m = m.to(dev)
    xg, bg = x.to(dev).requires_grad_(), bank.to(dev).requires_grad_()
    kinks = Kinks()
    with kinks.gpu():
        out = m(xg, ops.BankedRelation(bg, idx.to(dev)), self_padding_mask=mask.to(dev))
    with kinks.oracle():
        ref = O.graph_transformer(P, "", xc, O.bank_to_dense(bc, idx), L, H, self_padding_mask=mask)
    kinks.check()
    assert rel_err(out, ref) < TOL32
    compare_grads(m, P, (out * wo.to(dev)).sum(), (ref * wo).sum(), [xg, bg], [xc, bc], tol=TOL32, tol_max=TOL32)

    a = RelationMultiheadAttention(D, H, 0.0)
    boost(a, 3.0, gen)
    rel = torch.randn(N, N, B, D, generator=gen) * 0.5
    ww = torch.randn(N, N, B, H, generator=gen)
    P = oracle_params(a)
    xc, rc = x.clone().requires_grad_(), rel.clone().requires_grad_()
    ref, wref = O.rel_mha(P, "", xc, xc, xc, rc, H, mask, need_weights=True)
    a = a.to(dev)
    xg, rg = x.to(dev).requires_grad_(), rel.to(dev).requires_grad_()
    out, w = a(xg, xg, xg, rg, key_padding_mask=mask.to(dev), need_weights=True)
    assert rel_err(out, ref) < TOL32 and rel_err(w, wref) < TOL32
    compare_grads(a, P, (out * wo.to(dev)).sum() + (w * ww.to(dev)).sum(), (ref * wo).sum() + (wref * ww).sum(),
                  [xg, rg], [xc, rc], tol=TOL32, tol_max=TOL32)


@pytest.mark.parametrize("T,S,B,D,H,F,L", [(6, 8, 3, 32, 4, 64, 2), (30, 40, 16, 512, 8, 1024, 1)])
def test_transformer_fp32_mode_vs_fp32_oracle(dev, fp32_mode, T, S, B, D, H, F, L):
    from gtos_b200.transformer import Transformer
    gen = torch.Generator().manual_seed(SEED + 2)
    m = Transformer(L, D, F, H, 0.0, with_external=True)
    boost(m, 3.0 if D < 100 else 2.0, gen)
    x, kv = torch.randn(T, B, D, generator=gen), torch.randn(T, B, D, generator=gen)
    mem = torch.randn(S, B, D, generator=gen)
    tl = [T] + [int(v) for v in torch.randint(T // 2, T + 1, (B - 1,), generator=gen)]
    sl = [S] + [int(v) for v in torch.randint(S // 2, S + 1, (B - 1,), generator=gen)]
    tmask, smask = pad_mask(tl, T), pad_mask(sl, S)
    cm = O.causal_mask(T)
    wo = torch.randn(T, B, D, generator=gen)
    P = oracle_params(m)
    xc, kc, mc = (t.clone().requires_grad_() for t in (x, kv, mem))
    m = m.to(dev)
    xg, kg, mg = (t.to(dev).requires_grad_() for t in (x, kv, mem))
    kinks = Kinks()
    with kinks.gpu():
        out = m(xg, kv=kg, self_padding_mask=tmask.to(dev), self_attn_mask=cm.to(dev), external_memories=mg,
                external_padding_mask=smask.to(dev))
    with kinks.oracle():
        ref = O.transformer(P, "", xc, L, H, kv=kc, self_padding_mask=tmask, self_attn_mask=cm, external_memories=mc,
                            external_padding_mask=smask, with_external=True)
    kinks.check()
    assert rel_err(out, ref) < TOL32
    compare_grads(m, P, (out * wo.to(dev)).sum(), (ref * wo).sum(), [xg, kg, mg], [xc, kc, mc], tol=TOL32, tol_max=TOL32)
    kinks = Kinks()
    with kinks.gpu():
        out2 = m(xg, self_padding_mask=tmask.to(dev), self_attn_mask=cm.to(dev), external_memories=mg,
                 external_padding_mask=smask.to(dev))
    with kinks.oracle():
        ref2 = O.transformer(P, "", xc, L, H, self_padding_mask=tmask, self_attn_mask=cm, external_memories=mc,
                             external_padding_mask=smask, with_external=True)
    kinks.check()
    assert rel_err(out2, ref2) < TOL32
    compare_grads(m, P, (out2 * wo.to(dev)).sum(), (ref2 * wo).sum(), [xg, mg], [xc, mc], tol=TOL32, tol_max=TOL32)


@pytest.mark.parametrize("R,Lmax,rel_dim,hid,D,V", [(23, 4, 12, 16, 32, 19), (3000, 8, 100, 256, 512, 206)])
def test_relation_encoder_fp32_mode_vs_fp32_oracle(dev, fp32_mode, R, Lmax, rel_dim, hid, D, V):
    from gtos_b200.encoder import RelationEncoder
    gen = torch.Generator().manual_seed(SEED + 3)
    m = RelationEncoder(_V(V), rel_dim, D, hid, 2, 0.0)
    with torch.no_grad():
        m.rel_embed.weight.mul_(20.0)
    lengths = torch.randint(1, Lmax + 1, (R,), generator=gen)
    lengths[0] = Lmax
    tokens = torch.randint(2, V, (Lmax, R), generator=gen)
    tokens = tokens.masked_fill(torch.arange(Lmax).unsqueeze(1) >= lengths.unsqueeze(0), 0)
    wo = torch.randn(R, D, generator=gen)
    P = oracle_params(m)
    ref = O.relation_encoder(P, "", tokens, lengths, num_layers=2)
    m = m.to(dev)
    out = m(tokens.to(dev), lengths.to(dev))
    assert rel_err(out, ref) < TOL32
    compare_grads(m, P, (out * wo.to(dev)).sum(), (ref * wo).sum(), [], [], tol=TOL32, tol_max=TOL32)


@pytest.mark.parametrize("T,S,B,D,H,F,L,V,tok", [(6, 8, 3, 32, 4, 64, 2, 41, 24), (30, 40, 8, 512, 8, 1024, 3, 1000, 300)])
def test_decode_layer_fp32_mode_vs_fp32_oracle(dev, fp32_mode, T, S, B, D, H, F, L, V, tok):
    from gtos_b200.decoder import DecodeLayer
    gen = torch.Generator().manual_seed(SEED + 4)
    vocabs = {"predictable_token": _V(V)}
    m = DecodeLayer(vocabs, L, D, F, H, tok, 0, 0.0)
    boost(m, 3.0 if D < 100 else 2.0, gen)
    probe = torch.randn(1, B, D, generator=gen).expand(T, B, D).clone()
    graph, snt = torch.randn(S, B, D, generator=gen), torch.randn(T, B, D, generator=gen)
    tl = [T] + [int(v) for v in torch.randint(T // 2, T + 1, (B - 1,), generator=gen)]
    sl = [S] + [int(v) for v in torch.randint(S // 2, S + 1, (B - 1,), generator=gen)]
    tmask, smask = pad_mask(tl, T), pad_mask(sl, S)
    cm = O.causal_mask(T)
    copy_seq = torch.randint(2, V + 5, (S, B), generator=gen)
    target = torch.randint(2, V, (T, B), generator=gen).masked_fill(tmask, 0)
    P = oracle_params(m)
    pc, gc, sc = (t.clone().requires_grad_() for t in (probe, graph, snt))
    m = m.to(dev)
    pg, gg, sg = (t.to(dev).requires_grad_() for t in (probe, graph, snt))
    kinks = Kinks()
    with kinks.gpu():
        loss = m(pg, gg, sg, smask.to(dev), tmask.to(dev), cm.to(dev), copy_seq.to(dev), target=target.to(dev))
    with kinks.oracle():
        ref = O.decode_layer(P, "", pc, gc, sc, smask, tmask, cm, copy_seq, L, H, 0, target=target)
    kinks.check()
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item())
    compare_grads(m, P, loss, ref, [pg, gg, sg], [pc, gc, sc], tol=TOL32, tol_max=TOL32)
    with torch.no_grad():
        ll = m(pg, gg, sg, smask.to(dev), tmask.to(dev), cm.to(dev), copy_seq.to(dev), work=True)
    llref = O.decode_layer(P, "", pc, gc, sc, smask, tmask, cm, copy_seq, L, H, 0, work=True)
    assert rel_err(ll, llref) < TOL32


def test_hot_path_fp32_mode_loss_and_every_gradient_vs_fp32_oracle(dev, fp32_mode):
    """the assembled step (RelationEncoder -> index_select -> GraphTransformer -> snt Transformer -> DecodeLayer -> loss,
    generator.py:71-94,169-182) at the model's real widths on a small batch, weights x2.5: loss and the gradient of EVERY
    parameter against the fp32 oracle at 1e-3; then dropout on: the step runs and its loss / gradients are finite"""
    from gtos_b200 import hotpath, ops, synthetic
    cfg = hotpath.HotPathConfig(graph_layers=2, snt_layers=1, inference_layers=2, dropout=0.0, vocab_size=1000)
    D = cfg.embed_dim
    torch.manual_seed(SEED)
    model = hotpath.HotPath(cfg, relation_mode="index_select")
    gen = torch.Generator().manual_seed(5)
    boost(model, 2.5, gen)
    g = synthetic.make_batch(6, 20, D, T_max=14, T_min=7, V=1000, seed=SEED + 3)
    batch = hotpath.batch_tensors(g)
    P = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    names = [n for n, _ in model.named_parameters()]
    model = model.to(dev)
    dbatch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    kinks = Kinks()
    with kinks.gpu():
        loss = model(dbatch)
    with kinks.oracle():
        ref = HO.hotpath_loss(P, batch, cfg)
    kinks.check()
    g_cpu = torch.autograd.grad(ref, [P[n] for n in names], allow_unused=True)
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item()), (loss.item(), ref.item())
    g_gpu = torch.autograd.grad(loss, [p for _, p in model.named_parameters()], allow_unused=True)
    worst = 0.0
    for n, a, b in zip(names, g_gpu, g_cpu):
        assert (a is None) == (b is None), n
        if a is None:
            continue
        e2, em = l2_err(a, b), rel_err(a, b)
        worst = max(worst, e2, em)
        assert e2 < TOL32 and em < TOL32, f"{n}: rel L2 {e2:.2e}, max-norm {em:.2e}"
    # for the record (profiles/): the same model and batch in the default bf16 mode against the same oracle gradients
    import json
    import os
    with ops.precision_mode("bf16"):
        loss16 = model(dbatch)
        g16 = torch.autograd.grad(loss16, [p for _, p in model.named_parameters()], allow_unused=True)
    pairs = [(a, b, c) for a, b, c in zip(g_gpu, g16, g_cpu) if a is not None]
    report = {
        "what": "HotPath (2 graph + 1 snt + 2 inference layers, D=512, weights x2.5, 6 graphs) vs the fp32 oracle",
        "relu_units": kinks.units, "relu_units_in_band": kinks.in_band, "relu_choices_taken_from_gpu": kinks.flipped,
        "fp32_mode": {"loss_rel_err": abs(loss.item() - ref.item()) / abs(ref.item()),
                      "grad_rel_l2_worst": max(l2_err(a, c) for a, _, c in pairs),
                      "grad_max_norm_worst": max(rel_err(a, c) for a, _, c in pairs)},
        "bf16_mode": {"loss_rel_err": abs(loss16.item() - ref.item()) / abs(ref.item()),
                      "grad_rel_l2_worst": max(l2_err(b, c) for _, b, c in pairs),
                      "grad_max_norm_worst": max(rel_err(b, c) for _, b, c in pairs)}}
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "fp32_mode_errors.json"), "w") as f:
            json.dump(report, f, indent=1)
    assert report["bf16_mode"]["grad_rel_l2_worst"] > 10 * report["fp32_mode"]["grad_rel_l2_worst"]
    # training mode (dropout 0.2 everywhere): finite, reproducible from the device seed
    model.train()
    for mod in model.modules():
        if hasattr(mod, "dropout") and isinstance(mod.dropout, float):
            mod.dropout = 0.2
    ops.reseed(11, dev)
    l1 = model(dbatch)
    gr = torch.autograd.grad(l1, [p for _, p in model.named_parameters()], allow_unused=True)
    assert torch.isfinite(l1) and all(torch.isfinite(t).all() for t in gr if t is not None)
    assert abs(l1.item() - loss.item()) > 1e-6
