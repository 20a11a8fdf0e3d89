"""-m gpu: reference-supported branches that gtos itself never takes still COMPUTE - checked against
the reference's own modules (oracle/_ref, CPU fp32) at the bf16 tolerance:
  * RelationMultiheadAttention / GraphTransformerLayer with key = value = kv != query (graph_transformer.py:52-55,108-116)
  * RelationMultiheadAttention(weights_dropout=False) (graph_transformer.py:160-161)
  * MultiheadAttention with key is not value (transformer.py:113-118)
  * evaluation multi-path relation mean on the factorised path (generator.py:83-88) through ops.BankedRelation"""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from conftest import l2_err, rel_err             # noqa: E402
from oracle import ref_loader as RL              # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not RL.have_ref("generator"), reason="oracle/_ref not staged")]
SEED = 19940117


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    from gtos_b200 import _lib
    _lib.check(_lib.load().gtos_device_check(), "device_check")
    return torch.device("cuda:0")


def _boost(m, f=4.0):
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() >= 2 and "layer_norm" not in n:
                p.mul_(f)
            elif n.endswith("bias") and "layer_norm" not in n:
                p.normal_(0, 0.05)


def _inputs(N, B, D, seed=SEED):
    g = torch.Generator().manual_seed(seed)
    x = torch.nn.functional.layer_norm(torch.randn(N, B, D, generator=g), (D,))
    kv = torch.nn.functional.layer_norm(torch.randn(N, B, D, generator=g), (D,))
    rel = torch.randn(N, N, B, D, generator=g) * 0.5
    lens = torch.randint(N // 2, N + 1, (B,), generator=g)
    lens[0] = N
    mask = torch.arange(N).unsqueeze(1) >= lens.unsqueeze(0)
    return x, kv, rel, mask


def test_relation_attention_with_separate_kv_matches_reference(dev):
    from gtos_b200 import graph_transformer as GT
    ref_gt = RL.load("generator").graph_transformer
    N, B, D, H, F = 9, 3, 128, 8, 256
    torch.manual_seed(SEED)
    ref = ref_gt.GraphTransformerLayer(D, F, H, 0.0)
    _boost(ref)
    ours = GT.GraphTransformerLayer(D, F, H, 0.0)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(dev)
    x, kv, rel, mask = _inputs(N, B, D)
    xr, kvr, relr = x.clone().requires_grad_(), kv.clone().requires_grad_(), rel.clone().requires_grad_()
    y_ref, w_ref = ref(xr, relr, kv=kvr, self_padding_mask=mask, need_weights=True)
    (y_ref * y_ref).sum().backward()
    xo, kvo, relo = (t.to(dev).requires_grad_() for t in (x, kv, rel))
    y, w = ours(xo, relo, kv=kvo, self_padding_mask=mask.to(dev), need_weights=True)
    (y * y).sum().backward()
    assert rel_err(y, y_ref) < 1e-2 and rel_err(w, w_ref) < 1e-2
    # gradients: relative L2 (weights inflated x4, bf16 operands vs the fp32 reference: single elements near ReLU kinks move more)
    assert l2_err(kvo.grad, kvr.grad) < 5e-2 and l2_err(relo.grad, relr.grad) < 5e-2 and l2_err(xo.grad, xr.grad) < 5e-2
    g_ref = dict(ref.named_parameters())
    for n, p in ours.named_parameters():
        assert l2_err(p.grad, g_ref[n].grad) < 5e-2, n
    # the stack-level entry point takes the same branch
    enc_ref = ref_gt.GraphTransformer(2, D, F, H, 0.0)
    _boost(enc_ref)
    enc = GT.GraphTransformer(2, D, F, H, 0.0)
    enc.load_state_dict(enc_ref.state_dict())
    enc = enc.to(dev)
    with torch.no_grad():
        assert rel_err(enc(x.to(dev), rel.to(dev), kv=kv.to(dev), self_padding_mask=mask.to(dev)),
                       enc_ref(x, rel, kv=kv, self_padding_mask=mask)) < 1e-2


def test_relation_attention_output_dropout_variant(dev):
    """weights_dropout=False: dropout moves from the attention weights to the attention output (:160-161).  p = 0 is the
    exact path (vs the reference); p > 0 is checked statistically (keep rate, 1/(1-p) scaling through out_proj = identity)."""
    from gtos_b200 import graph_transformer as GT
    ref_gt = RL.load("generator").graph_transformer
    N, B, D, H = 11, 4, 128, 8
    torch.manual_seed(SEED + 1)
    ref = ref_gt.RelationMultiheadAttention(D, H, 0.0, weights_dropout=False)
    _boost(ref)
    ours = GT.RelationMultiheadAttention(D, H, 0.0, weights_dropout=False)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(dev)
    x, _, rel, mask = _inputs(N, B, D, SEED + 1)
    y_ref, _ = ref(x, x, x, rel, key_padding_mask=mask)
    xo = x.to(dev).requires_grad_()
    y, _ = ours(xo, xo, xo, rel.to(dev), key_padding_mask=mask.to(dev))
    assert rel_err(y, y_ref) < 1e-2
    with torch.no_grad():                       # out_proj = identity, zero bias: the output IS the dropped attention
        ours.out_proj.weight.copy_(torch.eye(D))
        ours.out_proj.bias.zero_()
    base, _ = ours(xo, xo, xo, rel.to(dev), key_padding_mask=mask.to(dev))
    ours.dropout = 0.25
    ours.train()
    got, _ = ours(xo, xo, xo, rel.to(dev), key_padding_mask=mask.to(dev))
    got.sum().backward()
    assert torch.isfinite(xo.grad).all()
    kept = got.ne(0) & base.abs().gt(1e-3)
    live = base.abs().gt(1e-3)
    rate = kept.sum().item() / live.sum().item()
    assert abs(rate - 0.75) < 0.03, rate
    ratio = (got[kept] / base[kept]).float()
    assert (ratio - 1 / 0.75).abs().max().item() < 2e-2          # bf16 rounding of the out_proj operand


def test_multihead_attention_with_separate_key_and_value(dev):
    from gtos_b200 import transformer as TF
    ref_tf = RL.load("generator").transformer
    T, S, B, D, H = 7, 10, 3, 128, 8
    torch.manual_seed(SEED + 2)
    ref = ref_tf.MultiheadAttention(D, H, 0.0)
    _boost(ref)
    ours = TF.MultiheadAttention(D, H, 0.0)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(dev)
    g = torch.Generator().manual_seed(SEED)
    q, k, v = (torch.randn(n, B, D, generator=g) for n in (T, S, S))
    mask = torch.arange(S).unsqueeze(1) >= torch.tensor([10, 4, 7]).unsqueeze(0)
    y_ref, w_ref = ref(q, k, v, key_padding_mask=mask, need_weights=True)
    qo, ko, vo = (t.to(dev).requires_grad_() for t in (q, k, v))
    y, w = ours(qo, ko, vo, key_padding_mask=mask.to(dev), need_weights=True)
    y.sum().backward()
    assert rel_err(y, y_ref) < 1e-2 and rel_err(w, w_ref) < 1e-2
    assert torch.isfinite(ko.grad).all() and torch.isfinite(vo.grad).all()


def test_banked_relation_evaluation_multipath_mean(dev):
    """generator.py:83-88 through the factorised path: BankedRelation(bank, idx[N,N,B,K]) == the dense mean tensor"""
    from gtos_b200 import graph_transformer as GT, ops
    ref_gt = RL.load("generator").graph_transformer
    N, B, D, H, F, R, K = 10, 3, 128, 8, 256, 40, 3
    torch.manual_seed(SEED + 3)
    enc_ref = ref_gt.GraphTransformer(2, D, F, H, 0.0)
    _boost(enc_ref)
    enc = GT.GraphTransformer(2, D, F, H, 0.0)
    enc.load_state_dict(enc_ref.state_dict())
    enc = enc.to(dev).eval()
    enc_ref.eval()
    g = torch.Generator().manual_seed(SEED + 3)
    x, _, _, mask = _inputs(N, B, D, SEED + 3)
    bank = torch.randn(R, D, generator=g) * 0.5
    idx = torch.randint(1, R, (N, N, B, K), generator=g)
    idx[..., 1:] = idx[..., 1:].masked_fill(torch.rand(N, N, B, K - 1, generator=g) < 0.5, 0)
    idx[0, :, :, :] = 0                                            # pairs with no path at all: count clamps to 1
    # the reference's own lines (generator.py:83-88)
    relation = bank.clone()
    relation[0, :] = 0.
    relation = relation[idx]
    dense = relation.sum(dim=3) / idx.ne(0).sum(dim=3).clamp_(min=1).unsqueeze(-1).type_as(relation)
    with torch.no_grad():
        y_ref = enc_ref(x, dense, self_padding_mask=mask)
        br = ops.BankedRelation(bank.to(dev), idx.to(dev))
        assert rel_err(br.relb.float(), dense) < 1e-2 and rel_err(br.dense(), dense) < 1e-6
        y = enc(x.to(dev), br, self_padding_mask=mask.to(dev))
        w = enc.get_attn_weights(x.to(dev), br, self_padding_mask=mask.to(dev))
        w_ref = enc_ref.get_attn_weights(x, dense, self_padding_mask=mask)
    assert rel_err(y, y_ref) < 1e-2 and rel_err(w, w_ref) < 1e-2
    # with gradients the multi-path bank takes the dense autograd route
    bank_g = bank.to(dev).requires_grad_()
    enc.train()
    enc(x.to(dev), ops.BankedRelation(bank_g, idx.to(dev)), self_padding_mask=mask.to(dev)).sum().backward()
    assert bank_g.grad is not None and torch.isfinite(bank_g.grad).all() and float(bank_g.grad[0].abs().max()) == 0.0


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_unidirectional_relation_encoder_matches_nn_gru(dev, mode):
    """RelationEncoder(bidirectional=False) (encoder.py:67,83,112; gtos itself never builds it): runs on the general
    split-operand Function in both modes, so the bound is the fp32 one.  Checker: torch's own packed nn.GRU on the CPU with
    the same parameters, exactly the reference's forward (encoder.py:90-119)."""
    import copy
    from gtos_b200 import ops
    from gtos_b200.encoder import RelationEncoder

    class V:
        size, padding_idx, unk_idx = 31, 0, 1

    torch.manual_seed(SEED + 5)
    R, Lmax, E, Hh, D = 57, 5, 20, 32, 64
    m = RelationEncoder(V(), E, D, Hh, 2, 0.0, bidirectional=False)
    with torch.no_grad():
        m.rel_embed.weight.mul_(20.0)
    lengths = torch.randint(1, Lmax + 1, (R,))
    lengths[0] = Lmax
    tokens = torch.randint(2, 31, (Lmax, R)).masked_fill(torch.arange(Lmax).unsqueeze(1) >= lengths.unsqueeze(0), 0)
    wo = torch.randn(R, D)
    ref = copy.deepcopy(m)
    ls, order = torch.sort(lengths, descending=True)
    x = ref.rel_embed(tokens.index_select(1, order))
    _, h = ref.rnn(torch.nn.utils.rnn.pack_padded_sequence(x, ls.tolist()))
    out_ref = ref.out_proj(h.view(2, 1, R, Hh)[-1, 0].index_select(0, torch.sort(order)[1]))
    (out_ref * wo).sum().backward()
    m = m.to(dev)
    with ops.precision_mode(mode):
        out = m(tokens.to(dev), lengths.to(dev))
        (out * wo.to(dev)).sum().backward()
    assert out.shape == out_ref.shape and rel_err(out, out_ref) < 1e-3
    gr = dict(ref.named_parameters())
    for n, p in m.named_parameters():
        assert l2_err(p.grad, gr[n].grad) < 1e-3, (n, l2_err(p.grad, gr[n].grad))
