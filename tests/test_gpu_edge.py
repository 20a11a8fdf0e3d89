"""-m gpu: edge shapes of the hot path and the BASELINE.json config-3 shard shape.

The reference has no tests; these cover what its data layer can produce at the extremes
(generator/data.py:126-176): a batch of one graph, a graph of a single node (+ <CLS>), a one-token target, a relation
bank with a single path, graphs whose padding covers almost the whole batch - and the translator-sized shard
(61 nodes incl. <CLS>, 16 graphs per GPU), each against the CPU oracle."""
import pytest
import torch

from conftest import rel_err
from oracle import gtos_oracle as O

pytestmark = pytest.mark.gpu
SEED = 19940117
TOL = 1e-2


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    from gtos_b200 import _lib
    _lib.check(_lib.load().gtos_device_check(), "device_check")
    return torch.device("cuda:0")


class _V:
    def __init__(self, size):
        self.size, self.padding_idx, self.unk_idx = size, 0, 1


def _params(m):
    return {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()}


@pytest.mark.parametrize("N,B,lens", [(2, 1, [2]), (3, 4, [3, 1, 2, 1]), (17, 2, [17, 1])])
def test_graph_transformer_tiny_and_mostly_padded(dev, N, B, lens):
    """one graph / single-node graphs (only <CLS> + 1 node valid, or only <CLS>): outputs and gradients of valid rows"""
    from gtos_b200.graph_transformer import GraphTransformer
    D, H, F, L = 128, 8, 256, 2
    gen = torch.Generator().manual_seed(SEED + N)
    torch.manual_seed(SEED)
    m = GraphTransformer(L, D, F, H, 0.0)
    P = _params(m)
    m = m.to(dev)
    x = torch.randn(N, B, D, generator=gen)
    rel = torch.randn(N, N, B, D, generator=gen) * 0.5
    mask = torch.arange(N).unsqueeze(1) >= torch.tensor(lens).unsqueeze(0)
    xg, rg = x.to(dev).requires_grad_(), rel.to(dev).requires_grad_()
    out = m(xg, rg, self_padding_mask=mask.to(dev))
    xc, rc = x.clone().requires_grad_(), rel.clone().requires_grad_()
    ref = O.graph_transformer(P, "", xc, rc, L, H, self_padding_mask=mask)
    valid = ~mask
    assert rel_err(out[valid.to(dev)], ref[valid]) < TOL
    w = torch.randn(N, B, D, generator=gen) * valid.unsqueeze(-1)
    (out * w.to(dev)).sum().backward()
    (ref * w).sum().backward()
    assert rel_err(xg.grad, xc.grad) < 5e-2 and rel_err(rg.grad, rc.grad) < 5e-2


def test_transformer_single_token_and_single_memory_slot(dev):
    from gtos_b200.transformer import Transformer
    D, H, F = 64, 4, 128
    torch.manual_seed(SEED + 1)
    m = Transformer(2, D, F, H, 0.0, with_external=True)
    P = _params(m)
    m = m.to(dev)
    gen = torch.Generator().manual_seed(SEED + 2)
    for T, S, B in [(1, 1, 1), (1, 5, 3), (4, 1, 2)]:
        x, mem = torch.randn(T, B, D, generator=gen), torch.randn(S, B, D, generator=gen)
        cm = O.causal_mask(T)
        with torch.no_grad():
            out = m(x.to(dev), self_attn_mask=cm.to(dev), external_memories=mem.to(dev))
            ref = O.transformer(P, "", x, 2, H, self_attn_mask=cm, external_memories=mem, with_external=True)
        assert rel_err(out, ref) < TOL, (T, S, B)


def test_relation_encoder_single_path_and_length_one(dev):
    from gtos_b200.encoder import RelationEncoder
    torch.manual_seed(SEED + 3)
    m = RelationEncoder(_V(30), 12, 64, 32, 2, 0.0)
    P = _params(m)
    m = m.to(dev)
    for tokens, lengths in [(torch.tensor([[7]]), torch.tensor([1])),
                            (torch.tensor([[2, 3, 4], [0, 5, 6], [0, 0, 9]]), torch.tensor([1, 2, 3]))]:
        with torch.no_grad():
            out = m(tokens.to(dev), lengths.to(dev))
            ref = O.relation_encoder(P, "", tokens, lengths)
        assert out.shape == ref.shape and rel_err(out, ref) < TOL


def test_decode_layer_single_token_target(dev):
    from gtos_b200.decoder import DecodeLayer
    D, H, F, V = 64, 4, 128, 50
    torch.manual_seed(SEED + 4)
    m = DecodeLayer({"predictable_token": _V(V)}, 1, D, F, H, 24, 6, 0.0)
    P = _params(m)
    m = m.to(dev)
    gen = torch.Generator().manual_seed(SEED + 5)
    T, S, B = 1, 2, 2
    probe, graph, snt = (torch.randn(T, B, D, generator=gen), torch.randn(S, B, D, generator=gen),
                         torch.randn(T, B, D, generator=gen))
    gmask = torch.tensor([[False, False], [False, True]])
    smask = torch.zeros(T, B, dtype=torch.bool)
    copy_seq = torch.tensor([[5, V + 1], [7, 0]])
    target = torch.tensor([[5, V + 1]])
    cm = O.causal_mask(T)
    loss = m(probe.to(dev), graph.to(dev), snt.to(dev), gmask.to(dev), smask.to(dev), cm.to(dev), copy_seq.to(dev),
             target=target.to(dev))
    ref = O.decode_layer(P, "", probe, graph, snt, gmask, smask, cm, copy_seq, 1, H, 0, target=target)
    assert abs(loss.item() - ref.item()) < TOL * max(1.0, abs(ref.item()))


def test_config3_shard_shape_vs_oracle(dev):
    """BASELINE.json config 3 per-GPU shard: 16 dependency graphs of <= 60 nodes (+ <CLS>), D=512, 4 layers: two graphs
    of the shard against the oracle run on each alone; permutation of the shard is bit-exact."""
    from gtos_b200 import ops, synthetic
    from gtos_b200.graph_transformer import GraphTransformer
    D, H, B = 512, 8, 16
    g = synthetic.make_batch(B, 60, D, seed=SEED + 60)
    N = g["N"]
    torch.manual_seed(SEED + 6)
    cpu = GraphTransformer(4, D, 1024, H, 0.0)
    m = GraphTransformer(4, D, 1024, H, 0.0).to(dev)
    m.load_state_dict(cpu.state_dict())
    gen = torch.Generator().manual_seed(SEED + 61)
    bank = torch.randn(g["relation_bank"].shape[1], D, generator=gen) * 0.5
    x, mask, idx = g["x"].to(dev), g["node_mask"].to(dev), g["relation"].to(dev)
    with torch.no_grad():
        out = m(x, ops.bank_gather(bank.to(dev), idx), self_padding_mask=mask)
        outb = m(x, ops.BankedRelation(bank.to(dev), idx), self_padding_mask=mask)
        perm = torch.randperm(B, generator=gen).to(dev)
        outp = m(x[:, perm], ops.bank_gather(bank.to(dev), idx[:, :, perm].contiguous()), self_padding_mask=mask[:, perm])
    assert torch.equal(out[:, perm], outp)
    # factorised relation: the fused gather kernel rounds ra / rb once more (projected bank in bf16); with
    # GTOS_BANKED_FWD=0 the dense bf16 operand feeds the same tcgen05 kernel and the outputs are bit-identical
    assert torch.equal(out, outb) if not ops.banked_fwd_supported(ops.BankedRelation(bank.to(dev), idx), N, B, D, H) \
        else rel_err(outb, out) < 5e-3
    P = {k: v.clone() for k, v in cpu.state_dict().items()}
    for b in (0, 11):
        relb = bank.index_select(0, g["relation"][:, :, b].reshape(-1)).view(N, N, 1, D)
        with torch.no_grad():
            ref = O.graph_transformer(P, "", g["x"][:, b:b + 1], relb, 4, H, self_padding_mask=g["node_mask"][:, b:b + 1])
        assert rel_err(out[:, b:b + 1], ref) < TOL


def test_hot_path_streams_and_planned_weight_copies_match_the_serial_schedule(dev, monkeypatch):
    """The step runs the GRU directions, the bank-side relation backward, weight-gradient GEMMs, LayerNorm parameter
    gradients and the weight operand copies on extra streams (ops.fork, ops.WeightPrepPlan).  Same kernels, same inputs:
    the loss must be bit-identical to the single-stream schedule and every gradient equal up to the summation order of
    the split-K / column-sum atomics; the second forward of a model (copies made ahead of time) must equal the first
    (copies made where they are used); an optimizer-style in-place weight update between two forwards must be seen."""
    from gtos_b200 import hotpath, ops, synthetic
    from conftest import l2_err
    cfg = hotpath.HotPathConfig(embed_dim=128, ff_embed_dim=256, num_heads=8, graph_layers=2, snt_layers=1,
                                inference_layers=1, rnn_hidden_size=64, dropout=0.0, vocab_size=500)
    torch.manual_seed(SEED)
    model = hotpath.HotPath(cfg).to(dev)
    g = synthetic.make_batch(8, 16, 128, T_max=12, T_min=6, V=500, seed=SEED)
    batch = {k: v.to(dev) for k, v in hotpath.batch_tensors(g).items()}
    params = list(model.parameters())

    def run():
        for p in params:
            p.grad = None
        loss = model(batch)
        loss.backward()
        torch.cuda.synchronize()
        return loss.detach().clone(), [None if p.grad is None else p.grad.clone() for p in params]

    with monkeypatch.context() as mp:
        for flag in ("_side_enabled", "_gru_streams", "_rel_streams", "_prep_ahead"):
            mp.setattr(ops, flag, False)
        loss_serial, grads_serial = run()
    model._prep_plan = ops.WeightPrepPlan()
    loss_first, grads_first = run()                 # records the plan, copies made in place
    assert model._prep_plan.items, "the first forward did not record the weight copies it used"
    loss_ahead, grads_ahead = run()                 # copies issued ahead on the third stream
    assert torch.equal(loss_first, loss_serial) and torch.equal(loss_ahead, loss_serial)
    for a, b, c in zip(grads_serial, grads_first, grads_ahead):
        assert (a is None) == (b is None) == (c is None)
        if a is not None:
            assert l2_err(b, a) < 1e-3 and l2_err(c, a) < 1e-3       # a race would be a gross error, not 1e-3
    with torch.no_grad():
        for p in params:
            p.mul_(1.25)                            # what an optimizer step does: same storage, new values
    loss_new, _ = run()
    with monkeypatch.context() as mp:
        for flag in ("_side_enabled", "_gru_streams", "_rel_streams", "_prep_ahead"):
            mp.setattr(ops, flag, False)
        loss_new_serial, _ = run()
    assert torch.equal(loss_new, loss_new_serial) and not torch.equal(loss_new, loss_serial)
