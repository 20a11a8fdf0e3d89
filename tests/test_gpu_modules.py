"""-m gpu: the drop-in nn.Modules (CUDA path through the C ABI) against the CPU oracle on the same
seeded inputs and weights.  Compute is bf16-operand / fp32-accumulate tensor-core math, so the bar is
the north star's bf16 tolerance: 1e-2 of the tensor's max magnitude, for outputs AND gradients."""
import pytest
import torch

from conftest import rel_err
from oracle import gtos_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-2
SEED = 19940117


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    from gtos_b200 import _lib
    _lib.check(_lib.load().gtos_device_check(), "device_check")
    return torch.device("cuda:0")


def boost(module, factor, gen):
    with torch.no_grad():
        for n, p in module.named_parameters():
            if p.dim() >= 2 and "layer_norm" not in n:
                p.mul_(factor)
            elif "bias" in n:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
            elif "layer_norm.weight" in n:
                p.add_(torch.randn(p.shape, generator=gen) * 0.1)


def pad_mask(lens, n):
    return torch.arange(n).unsqueeze(1) >= torch.tensor(lens).unsqueeze(0)


def oracle_params(module):
    return {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in module.state_dict().items()}


def compare_grads(module, P, loss_gpu, loss_cpu, ins_gpu, ins_cpu, tol=TOL):
    names = [n for n, _ in module.named_parameters()]
    g_gpu = torch.autograd.grad(loss_gpu, ins_gpu + [p for _, p in module.named_parameters()], allow_unused=True)
    g_cpu = torch.autograd.grad(loss_cpu, ins_cpu + [P[n] for n in names], allow_unused=True)
    labels = [f"input{i}" for i in range(len(ins_gpu))] + names
    worst = 0.0
    for lab, a, b in zip(labels, g_gpu, g_cpu):
        assert (a is None) == (b is None), lab
        if a is None:
            continue
        e = rel_err(a, b)
        worst = max(worst, e)
        assert e < tol, f"grad {lab}: rel err {e:.3e}"
    return worst


@pytest.mark.parametrize("N,B,D,H,F,L,wf", [(17, 8, 128, 8, 256, 2, 1.0), (17, 8, 128, 8, 256, 2, 6.0),
                                            (41, 6, 512, 8, 1024, 2, 4.0)])
def test_graph_transformer_vs_oracle(dev, N, B, D, H, F, L, wf):
    from gtos_b200.graph_transformer import GraphTransformer
    gen = torch.Generator().manual_seed(SEED)
    m = GraphTransformer(L, D, F, H, 0.0)
    boost(m, wf, gen)
    x = torch.randn(N, B, D, generator=gen)
    rel = torch.randn(N, N, B, D, generator=gen) * 0.5           # asymmetric on purpose (layout bugs)
    lens = [N] + [int(v) for v in torch.randint(N // 2, N + 1, (B - 1,), generator=gen)]
    mask = pad_mask(lens, N)
    wo = torch.randn(N, B, D, generator=gen)
    P = oracle_params(m)
    xc, rc = x.clone().requires_grad_(), rel.clone().requires_grad_()
    ref = O.graph_transformer(P, "", xc, rc, L, H, self_padding_mask=mask)
    m = m.to(dev)
    xg, rg = x.to(dev).requires_grad_(), rel.to(dev).requires_grad_()
    out = m(xg, rg, self_padding_mask=mask.to(dev))
    assert rel_err(out, ref) < TOL
    compare_grads(m, P, (out * wo.to(dev)).sum(), (ref * wo).sum(), [xg, rg], [xc, rc])
    with torch.no_grad():
        attn = m.get_attn_weights(xg, rg, self_padding_mask=mask.to(dev))
    aref = O.graph_transformer(P, "", xc, rc, L, H, self_padding_mask=mask, return_weights=True)
    assert attn.shape == aref.shape and rel_err(attn, aref) < TOL


def test_rel_mha_weights_grad(dev):
    """RelationMultiheadAttention.forward with need_weights and a gradient flowing into the weights."""
    from gtos_b200.graph_transformer import RelationMultiheadAttention
    gen = torch.Generator().manual_seed(SEED + 1)
    N, B, D, H = 17, 4, 128, 8
    m = RelationMultiheadAttention(D, H, 0.0)
    boost(m, 5.0, gen)
    x = torch.randn(N, B, D, generator=gen)
    rel = torch.randn(N, N, B, D, generator=gen) * 0.5
    mask = pad_mask([17, 9, 12, 15], N)
    wo, ww = torch.randn(N, B, D, generator=gen), torch.randn(N, N, B, H, generator=gen)
    P = oracle_params(m)
    xc, rc = x.clone().requires_grad_(), rel.clone().requires_grad_()
    ref, wref = O.rel_mha(P, "", xc, xc, xc, rc, H, mask, need_weights=True)
    m = m.to(dev)
    xg, rg = x.to(dev).requires_grad_(), rel.to(dev).requires_grad_()
    out, w = m(xg, xg, xg, rg, key_padding_mask=mask.to(dev), need_weights=True)
    assert w.shape == wref.shape
    assert rel_err(out, ref) < TOL and rel_err(w, wref) < TOL
    compare_grads(m, P, (out * wo.to(dev)).sum() + (w * ww.to(dev)).sum(), (ref * wo).sum() + (wref * ww).sum(),
                  [xg, rg], [xc, rc])


@pytest.mark.parametrize("T,S,B,D,H,F,L", [(6, 8, 3, 32, 4, 64, 2), (30, 40, 16, 512, 8, 1024, 1)])
def test_transformer_external_vs_oracle(dev, T, S, B, D, H, F, L):
    from gtos_b200.transformer import Transformer
    gen = torch.Generator().manual_seed(SEED + 2)
    m = Transformer(L, D, F, H, 0.0, with_external=True)
    boost(m, 5.0 if D < 100 else 3.0, gen)
    x, kv = torch.randn(T, B, D, generator=gen), torch.randn(T, B, D, generator=gen)
    mem = torch.randn(S, B, D, generator=gen)
    tl = [T] + [int(v) for v in torch.randint(T // 2, T + 1, (B - 1,), generator=gen)]
    sl = [S] + [int(v) for v in torch.randint(S // 2, S + 1, (B - 1,), generator=gen)]
    tmask, smask = pad_mask(tl, T), pad_mask(sl, S)
    cm = O.causal_mask(T)
    wo = torch.randn(T, B, D, generator=gen)
    P = oracle_params(m)
    xc, kc, mc = (t.clone().requires_grad_() for t in (x, kv, mem))
    ref = O.transformer(P, "", xc, L, H, kv=kc, self_padding_mask=tmask, self_attn_mask=cm, external_memories=mc,
                        external_padding_mask=smask, with_external=True)
    m = m.to(dev)
    xg, kg, mg = (t.to(dev).requires_grad_() for t in (x, kv, mem))
    out = m(xg, kv=kg, self_padding_mask=tmask.to(dev), self_attn_mask=cm.to(dev), external_memories=mg,
            external_padding_mask=smask.to(dev))
    assert rel_err(out, ref) < TOL
    compare_grads(m, P, (out * wo.to(dev)).sum(), (ref * wo).sum(), [xg, kg, mg], [xc, kc, mc])
    # self-attention path (kv=None), causal
    ref2 = O.transformer(P, "", xc, L, H, self_padding_mask=tmask, self_attn_mask=cm, external_memories=mc,
                         external_padding_mask=smask, with_external=True)
    out2 = m(xg, self_padding_mask=tmask.to(dev), self_attn_mask=cm.to(dev), external_memories=mg,
             external_padding_mask=smask.to(dev))
    assert rel_err(out2, ref2) < TOL
    compare_grads(m, P, (out2 * wo.to(dev)).sum(), (ref2 * wo).sum(), [xg, mg], [xc, mc])


def test_mha_golden_shapes_and_weights(dev, golden):
    """MultiheadAttention against the committed golden vectors of the reference module (D=32, hd=8)."""
    from gtos_b200.transformer import MultiheadAttention
    g = golden["mha_cross"]
    c = g["cfg"]
    m = MultiheadAttention(c["D"], c["H"], 0.0)
    m.load_state_dict(g["state"])
    m = m.to(dev)
    q, mem = g["q"].to(dev).requires_grad_(), g["mem"].to(dev).requires_grad_()
    out, w = m(q, mem, mem, key_padding_mask=g["smask"].to(dev), need_weights=True)
    assert rel_err(out, g["out"]) < TOL and rel_err(w, g["w"]) < TOL
    gq, gmem = torch.autograd.grad((out * g["wo"].to(dev)).sum() + (w * g["ww"].to(dev)).sum(), [q, mem])
    assert rel_err(gq, g["gq"]) < 2 * TOL and rel_err(gmem, g["gmem"]) < 2 * TOL
    g = golden["mha_self"]
    m.load_state_dict(g["state"])
    q = g["q"].to(dev).requires_grad_()
    out, w = m(q, q, q, key_padding_mask=g["tmask"].to(dev), attn_mask=g["cm"].to(dev), need_weights=True)
    assert rel_err(out, g["out"]) < TOL and rel_err(w, g["w"]) < TOL
    (gq,) = torch.autograd.grad((out * g["wo"].to(dev)).sum(), [q])
    assert rel_err(gq, g["gq"]) < 2 * TOL
