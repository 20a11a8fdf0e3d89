"""-m gpu: the widened rows of SURVEY.md §8(f) -- incremental beam decode (f-1), work=True log-prob table (f-2) and the
flat Adam step (f-4) -- through the C ABI, against the oracles (which are pinned to golden runs of the reference's own
Generator.decode_step / search.py / adam.py, tests/golden/make_golden_{decode,beam,optim}.py).

Tolerances: the decode path computes with bf16 operands / fp32 accumulation like the rest of the path -> 1e-2 on
probabilities (north star's bf16 tolerance); the optimizer is plain fp32 -> 1e-5.
"""
import os
import types

import pytest
import torch

from conftest import rel_err
from oracle import gtos_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-2
SEED = 19940117
GDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


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


def _modules(c, dev, state=None, seed=SEED):
    from gtos_b200.decoder import DecodeLayer
    from gtos_b200.transformer import Transformer
    torch.manual_seed(seed)
    vocabs = {"predictable_token": _V(c["V"])}
    snt = Transformer(c["snt_layers"], c["D"], c["F"], c["H"], 0.2, with_external=True)
    dec = DecodeLayer(vocabs, c["inference_layers"], c["D"], c["F"], c["H"], c["tok_dim"], 6, 0.2)
    if state is not None:
        snt.load_state_dict({k[len("snt_encoder."):]: v for k, v in state.items() if k.startswith("snt_encoder.")})
        dec.load_state_dict({k[len("decoder."):]: v for k, v in state.items() if k.startswith("decoder.")})
    return snt.to(dev).eval(), dec.to(dev).eval()


# ---------------------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Hyp,L,H,hd,Bc,mode", [(7, 5, 4, 8, 3, "cross"), (64, 40, 8, 64, 9, "cross"), (33, 17, 1, 512, 5, "cross"),
                                                (50, 23, 8, 64, 50, "self"), (2048, 40, 8, 64, 256, "cross"),
                                                (16, 300, 2, 16, 16, "self"), (40, 33, 8, 16, 7, "cross"), (20, 9, 4, 64, 20, "self"),
                                                (24, 50, 2, 128, 6, "cross"), (9, 600, 8, 64, 9, "self")])
def test_attn_decode_kernel(dev, Hyp, L, H, hd, Bc, mode):
    from gtos_b200 import decode
    gen = torch.Generator().manual_seed(SEED + Hyp + L)
    D = H * hd
    q = torch.randn(Hyp, D, generator=gen).to(dev)
    cache = (torch.randn(L, Bc, 2 * D, generator=gen) * 0.7).to(torch.bfloat16).to(dev)
    if mode == "cross":
        slot = torch.randint(0, Bc, (Hyp,), generator=gen).to(torch.int32).to(dev)
        slot_ld = 0
        lens = torch.randint(1, L + 1, (Bc,), generator=gen)
        pad = (torch.arange(L).unsqueeze(1) >= lens.unsqueeze(0)).to(torch.uint8).to(dev)
        rows = slot.long().unsqueeze(0).expand(L, Hyp)
    else:
        slot = torch.randint(0, Bc, (L, Hyp), generator=gen).to(torch.int32).to(dev)
        slot_ld = Hyp
        pad = None
        rows = slot.long()
    out, probs = decode.attn_decode(q, (cache, 0), 2 * D, D, Bc, L, H, hd, slot, slot_ld, pad, Bc, hd ** -0.5, want_probs=True)
    torch.cuda.synchronize()
    kv = cache.float()[torch.arange(L, device=dev).unsqueeze(1), rows]              # [L, Hyp, 2D]
    k, v = kv[..., :D].view(L, Hyp, H, hd), kv[..., D:].view(L, Hyp, H, hd)
    s = torch.einsum("nhd,lnhd->nhl", q.view(Hyp, H, hd) * hd ** -0.5, k)
    if pad is not None:
        s = s.masked_fill(pad.bool()[torch.arange(L, device=dev).unsqueeze(1), rows].t().unsqueeze(1), float("-inf"))
    w = torch.softmax(s, -1)
    ref = torch.einsum("nhl,lnhd->nhd", w, v).reshape(Hyp, D)
    assert rel_err(probs, w) < 1e-4
    assert rel_err(out.float(), ref) < 6e-3                                        # output is rounded to bf16


@pytest.mark.parametrize("rows,V,S,Bsrc,ext", [(5, 19, 6, 3, 4), (64, 1000, 40, 8, 16), (12, 50, 9, 12, 0)])
def test_token_logprob_kernel(dev, rows, V, S, Bsrc, ext):
    from gtos_b200 import ops
    gen = torch.Generator().manual_seed(SEED + rows)
    logits = (torch.randn(rows, V, generator=gen) * 3).to(dev)
    gate = torch.randn(rows, 2, generator=gen).to(dev)
    align = torch.softmax(torch.randn(rows, S, generator=gen), -1).to(dev)
    copy_seq = torch.randint(2, V + max(ext, 1), (S, Bsrc), generator=gen).to(dev)
    copy_seq[0] = copy_seq[1]                                                      # duplicate slots must accumulate
    src = torch.randint(0, Bsrc, (rows,), generator=gen).to(torch.int32).to(dev)
    W = max(V, int(copy_seq.max()) + 1)
    for src_index in (src, None):
        table = ops.token_logprob(logits, gate, align, copy_seq, src_index, W, B=Bsrc)
        b = src.long() if src_index is not None else torch.arange(rows, device=dev) % Bsrc
        g = torch.softmax(gate, -1)
        p = torch.zeros(rows, W, device=dev)
        p[:, :V] = torch.softmax(logits, -1) * g[:, :1]
        p.scatter_add_(1, copy_seq.t()[b], align * g[:, 1:])
        ref = (p + 1e-12).log()
        assert (table - ref).abs().max() < 1e-4
        assert rel_err(table.exp(), p) < 1e-5


@pytest.mark.parametrize("rows,V,S,Bsrc,ext,K", [(5, 19, 6, 3, 4, 3), (64, 1000, 40, 8, 16, 8), (33, 10000, 40, 9, 16, 8),
                                                  (12, 50, 9, 12, 0, 1), (7, 40, 5, 2, 3, 16)])
def test_token_topk_kernel(dev, rows, V, S, Bsrc, ext, K):
    """fused log-prob row + top-k (generator.py:157) vs the table kernel + torch.topk"""
    from gtos_b200 import ops
    gen = torch.Generator().manual_seed(SEED + rows + K)
    logits = (torch.randn(rows, V, generator=gen) * 3).to(dev)
    gate = torch.randn(rows, 2, generator=gen).to(dev)
    align = torch.softmax(torch.randn(rows, S, generator=gen) * 2, -1).to(dev)
    copy_seq = torch.randint(2, V + max(ext, 1), (S, Bsrc), generator=gen).to(dev)
    src = torch.randint(0, Bsrc, (rows,), generator=gen).to(torch.int32).to(dev)
    W = max(V, int(copy_seq.max()) + 1)
    ref = ops.token_logprob(logits, gate, align, copy_seq, src, W)
    val, idx, table = ops.token_topk(logits, gate, align, copy_seq, src, W, K, want_table=True)
    assert (table - ref).abs().max() < 1e-5
    rv, ri = torch.topk(ref, K, dim=1)
    assert (val - rv).abs().max() < 1e-5                                   # same k best values, best first
    assert (ref.gather(1, idx.long()) - val).abs().max() < 1e-5            # and the ids point at them
    assert all(len(set(r)) == K for r in idx.tolist())                     # no token twice
    val2, idx2 = ops.token_topk(logits, gate, align, copy_seq, src, W, K)
    assert torch.equal(val2, val) and torch.equal(idx2, idx)


@pytest.mark.parametrize("B,K,W,Tmin,Tmax", [(7, 4, 30, 2, 10), (5, 8, 80, 1, 12), (9, 1, 12, 1, 9), (3, 16, 300, 3, 8)])
def test_beam_update_kernel_matches_beam_state(dev, B, K, W, Tmin, Tmax):
    """gtos_beam_update (one kernel) vs BeamState.update (the torch implementation pinned to the reference's Beam class by
    tests/test_beam_cpu.py): identical state after every step, driven by the same top-k lists"""
    from gtos_b200.decode import BeamState, BeamStateFused
    END, UNK = 3, 1
    gen = torch.Generator().manual_seed(SEED + B * K)
    a = BeamState(B, K, Tmax, Tmin, END, UNK, dev)
    f = BeamStateFused(B, K, Tmax, Tmin, END, UNK, dev)
    parent = torch.zeros(B * K, dtype=torch.int32, device=dev)
    last = torch.zeros(B * K, dtype=torch.int64, device=dev)
    base = (torch.arange(B, device=dev) * K).unsqueeze(1)
    for t in range(Tmax):
        logits = torch.randn(B * K, W, generator=gen) * 2
        logits[:, END] += 0.6 * t
        logits[:, UNK] += 1.5
        table = torch.log_softmax(logits, -1).to(dev)
        tv, ti = torch.topk(table.view(B, K, -1), K, dim=-1)
        # BeamState.update runs the same torch.topk internally
        par, tok = a.update(t, table)
        f.update(t, tv.reshape(B * K, K).contiguous(), ti.reshape(B * K, K).to(torch.int32).contiguous(), parent, last)
        torch.cuda.synchronize()
        assert torch.equal(f.score, a.score), t
        assert torch.equal(f.live.bool(), a.live) and torch.equal(f.n_done.long(), a.n_done) and torch.equal(f.steps.long(), a.steps)
        assert torch.equal(f.tok[t].long(), a.tok[t]) and torch.equal(f.par[t].long(), a.par[t]), t
        assert torch.equal(f.done_score, a.done_score) and torch.equal(f.done_step.long(), a.done_step)
        assert torch.equal(f.done_par.long(), a.done_par)
        assert torch.equal(parent.long().view(B, K), par + base) and torch.equal(last.view(B, K), tok)
    assert int(a.n_done.sum()) > 0 or Tmin >= Tmax
    assert a.k_best(K, 0.6) == f.k_best(K, 0.6)


def test_token_generator_work_mode_vs_oracle(dev):
    """DecodeLayer(work=True) through the drop-in module (decoder.py:76-91) vs the oracle"""
    c = dict(D=64, F=128, H=8, snt_layers=1, inference_layers=2, tok_dim=40, V=300)
    _, dec = _modules(c, dev)
    gen = torch.Generator().manual_seed(SEED + 3)
    T, S, B = 5, 9, 4
    P = {"decoder." + k: v.detach().cpu() for k, v in dec.state_dict().items()}
    probe = torch.randn(T, B, c["D"], generator=gen)
    graph = torch.randn(S, B, c["D"], generator=gen)
    snt = torch.randn(T, B, c["D"], generator=gen)
    gmask = torch.arange(S).unsqueeze(1) >= torch.tensor([9, 4, 7, 2]).unsqueeze(0)
    copy_seq = torch.randint(2, c["V"] + 7, (S, B), generator=gen)
    cm = O.causal_mask(T)
    with torch.no_grad():
        ll = dec(probe.to(dev), graph.to(dev), snt.to(dev), gmask.to(dev), None, cm.to(dev), copy_seq.to(dev), work=True)
        ref = O.decode_layer(P, "decoder.", probe, graph, snt, gmask, None, cm, copy_seq, 2, c["H"], 0, work=True)
    assert ll.shape == ref.shape
    assert rel_err(ll.exp(), ref.exp()) < TOL


# ---------------------------------------------------------------------------------------------------------------
# engine vs the reference's own decode_step (golden) and vs the module path at config-5 size
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["one_snt_layer", "two_snt_layers"])
def test_decode_engine_matches_reference_decode_step(dev, case):
    from gtos_b200.decode import DecodeEngine
    g = torch.load(os.path.join(GDIR, "golden_decode_v1.pt"), map_location="cpu", weights_only=False)[case]
    c = g["cfg"]
    snt, dec = _modules(c, dev, state=g["state"])
    eng = DecodeEngine(snt, dec, max_hyp=8, max_steps=len(g["steps"]))
    m = g["mem"]
    eng.set_memory(m["graph_state"].to(dev), m["graph_padding_mask"].to(dev), m["probe"].to(dev), m["cp_seq"].to(dev))
    for t, st in enumerate(g["steps"]):
        par = st["parent"].to(torch.int32).to(dev) if st["parent"] is not None else None
        ll = eng.step(st["token_repr"][0].to(dev), st["src"].to(torch.int32).to(dev), par, t)
        assert ll.shape == st["ll"].shape
        assert rel_err(ll.exp(), st["ll"].exp()) < TOL, (case, t)
        big = st["ll"] > -7
        assert (ll.cpu()[big] - st["ll"][big]).abs().max() < 0.05, (case, t)


def _embed_fn(table, pos):
    def fn(tok, t):
        return torch.nn.functional.layer_norm(table[tok] + pos[t], (table.shape[1],))
    return fn


def test_decode_engine_matches_module_path_at_config5_size(dev):
    """cfg5: beam 8 x batch 256 = 2048 live hypotheses over ~40-node graphs, D=512.  The engine (cached K/V, ancestry
    table) must reproduce what the unchanged caller gets from the drop-in modules with index_select-ed memory."""
    from gtos_b200.decode import DecodeEngine
    c = dict(D=512, F=1024, H=8, snt_layers=1, inference_layers=3, tok_dim=300, V=2000)
    snt, dec = _modules(c, dev)
    B, K, S, D = 256, 8, 40, 512
    Hyp = B * K
    gen = torch.Generator().manual_seed(SEED + 5)
    graph = torch.randn(S, B, D, generator=gen).to(dev)
    lens = torch.randint(20, S + 1, (B,), generator=gen)
    gmask = (torch.arange(S).unsqueeze(1) >= lens.unsqueeze(0)).to(dev)
    probe = torch.tanh(torch.randn(1, B, D, generator=gen)).to(dev)
    copy_seq = torch.randint(2, c["V"] + 16, (S, B), generator=gen).to(dev)
    eng = DecodeEngine(snt, dec, max_hyp=Hyp, max_steps=4)
    eng.set_memory(graph, gmask, probe, copy_seq)
    src = torch.arange(B, device=dev).repeat_interleave(K)
    state = {}
    with torch.no_grad():
        for t in range(3):
            x = torch.nn.functional.layer_norm(torch.randn(1, Hyp, D, generator=gen), (D,)).to(dev)
            parent = None
            if t > 0:                                           # re-parent inside each beam, as Beam.update does
                parent = (torch.randint(0, K, (Hyp,), generator=gen).to(dev) + src * K)
                state = {k: v.index_select(1, parent) for k, v in state.items()}
            ll = eng.step(x[0], src.to(torch.int32), parent.to(torch.int32) if parent is not None else None, t)
            # module path, wired as generator.py:133-150 with search.py's index_select-ed memory
            kv = torch.cat([state["r0"], x], 0) if "r0" in state else x
            state["r0"] = kv
            g_sel, m_sel = graph.index_select(1, src), gmask.index_select(1, src)
            y, _, _ = snt.layers[0](x, kv=kv, external_memories=g_sel, external_padding_mask=m_sel)
            ts = torch.cat([state["ts"], y], 0) if "ts" in state else y
            state["ts"] = ts
            ref = dec(probe.index_select(1, src), g_sel, ts, m_sel, None, None, copy_seq.index_select(1, src), work=True)[0]
            assert ll.shape == ref.shape
            assert rel_err(ll.exp(), ref.exp()) < TOL, t


def test_beam_search_device_scores_are_consistent_and_graphs_are_exact(dev):
    """Run the device beam search end to end; every returned hypothesis' score must equal the sum of the log-probs of
    its tokens when the sequence is re-scored through the module path (catches any cache / ancestry / re-parenting
    slip), and replaying the steps as CUDA graphs must give bit-identical results."""
    from gtos_b200.decode import BeamSearchDevice, DecodeEngine
    c = dict(D=64, F=128, H=8, snt_layers=1, inference_layers=2, tok_dim=40, V=60)
    snt, dec = _modules(c, dev, seed=SEED + 11)
    with torch.no_grad():
        for mod in (snt, dec):
            for n, p in mod.named_parameters():
                if p.dim() >= 2 and "layer_norm" not in n:
                    p.mul_(4.0)
    B, K, S, D, Tmax = 6, 4, 7, 64, 9
    END, UNK, START = 3, 1, 2
    gen = torch.Generator().manual_seed(SEED + 13)
    graph = torch.randn(S, B, D, generator=gen).to(dev)
    gmask = (torch.arange(S).unsqueeze(1) >= torch.tensor([7, 3, 5, 7, 2, 6]).unsqueeze(0)).to(dev)
    probe = torch.tanh(torch.randn(1, B, D, generator=gen)).to(dev)
    copy_seq = torch.randint(4, c["V"] + 5, (S, B), generator=gen).to(dev)
    W = max(c["V"], int(copy_seq.max()) + 1)
    emb = torch.randn(W, D, generator=gen).to(dev)
    pos = torch.randn(Tmax, D, generator=gen).to(dev)
    with torch.no_grad():
        dec.token_generator.generator.bias[END] += 3.0         # make <END> reachable within Tmax steps
    eng = DecodeEngine(snt, dec, max_hyp=B * K, max_steps=Tmax)
    eng.set_memory(graph, gmask, probe, copy_seq, table_width=W)
    bs = BeamSearchDevice(eng, K, Tmax, 1, END, UNK, START, _embed_fn(emb, pos))
    best = bs.run().k_best(K, 0.6)
    n_done = bs.state.n_done.clone()
    assert int(n_done.sum()) > 0                               # some hypotheses completed with <END>
    # re-score through the module path
    with torch.no_grad():
        for b in range(B):
            for seq, score in best[b]:
                if score == float("-inf"):
                    continue
                toks = [START] + seq[:-1] if seq[-1] == END else [START] + seq
                toks = toks[:len(seq)]
                T = len(toks)
                x = torch.stack([_embed_fn(emb, pos)(torch.tensor([tk], device=dev), t)[0] for t, tk in enumerate(toks)]
                                ).view(T, 1, D)
                cm = O.causal_mask(T).to(dev)
                g1, m1 = graph[:, b:b + 1], gmask[:, b:b + 1]
                y = snt(x, self_attn_mask=cm, external_memories=g1, external_padding_mask=m1)
                ll = dec(probe[:, b:b + 1].expand(T, 1, D), g1, y, m1, None, cm, copy_seq[:, b:b + 1], work=True)
                tot = sum(float(ll[t, 0, seq[t]]) for t in range(T))
                assert abs(tot - score) < 0.05 * max(1.0, abs(score)), (b, seq, tot, score)
    # CUDA-graph replay of the same search
    bs2 = BeamSearchDevice(eng, K, Tmax, 1, END, UNK, START, _embed_fn(emb, pos), use_graphs=True)
    bs2.capture()
    best2 = bs2.run().k_best(K, 0.6)
    assert best2 == best
    # a new batch with the same B but another memory length S: set_memory replaces the engine's buffers, so the graphs
    # captured above (old pointers, old S) must be dropped, not replayed
    S2 = S + 3
    graph2 = torch.randn(S2, B, D, generator=gen).to(dev)
    gmask2 = (torch.arange(S2).unsqueeze(1) >= torch.tensor([10, 3, 5, 8, 2, 6]).unsqueeze(0)).to(dev)
    copy2 = torch.randint(4, c["V"] + 5, (S2, B), generator=gen).to(dev)
    eng.set_memory(graph2, gmask2, probe, copy2, table_width=W)
    got = bs2.run().k_best(K, 0.6)                             # would replay stale graphs without the epoch check
    assert not bs2._graphs
    want = BeamSearchDevice(eng, K, Tmax, 1, END, UNK, START, _embed_fn(emb, pos)).run().k_best(K, 0.6)
    assert got == want
    bs2.capture()
    assert bs2.run().k_best(K, 0.6) == want
    eng.set_memory(graph, gmask, probe, copy_seq, table_width=W)
    # the unfused bookkeeping (full table + torch.topk + BeamState) finds the same hypotheses
    bs3 = BeamSearchDevice(eng, K, Tmax, 1, END, UNK, START, _embed_fn(emb, pos), fused=False)
    best3 = bs3.run().k_best(K, 0.6)
    assert [[h[0] for h in b] for b in best3] == [[h[0] for h in b] for b in best]
    for b3, b1 in zip(best3, best):
        for h3, h1 in zip(b3, b1):
            assert h3[1] == pytest.approx(h1[1], abs=1e-4) or h3[1] == h1[1]


def test_flat_adam_matches_reference_optimizer(dev):
    """FlatAdam (gtos_grad_sumsq + gtos_adam_step) vs the reference's clip_grad_norm_ + AdamWeightDecayOptimizer golden"""
    from gtos_b200.optim import FlatAdam, noam_lr
    g = torch.load(os.path.join(GDIR, "golden_optim_v1.pt"), map_location="cpu", weights_only=False)
    params = {n: torch.nn.Parameter(v.clone().to(dev)) for n, v in g["init"].items()}
    for bind in (True, False):
        for n, v in g["init"].items():
            params[n].data = v.clone().to(dev)
            params[n].grad = None
        opt = FlatAdam(params.items(), lr=1e-3, eps=1e-6, weight_decay=1e-4, max_norm=1.0, bind_grads=bind)
        for k, st in enumerate(g["steps"], start=1):
            opt.zero_grad()
            for n, p in params.items():
                if bind:
                    p.grad.copy_(st["grads"][n].to(dev))
                else:
                    p.grad = st["grads"][n].to(dev)
            lr = noam_lr(g["embed_size"], k, g["warmup"])
            assert lr == pytest.approx(st["lr"], rel=1e-12)
            opt.set_lr(lr)
            opt.step()
            assert float(opt.grad_norm()) == pytest.approx(st["total_norm"], rel=1e-5)
            for n, p in params.items():
                assert rel_err(p, st["params"][n]) < 1e-5, (bind, k, n)
        off = 0
        for p in opt.params:                                    # parameters really live in the flat buffer
            assert p.data_ptr() == opt.flat.data_ptr() + 4 * off
            off += p.numel()


def test_flat_adam_picks_up_gradients_written_outside_the_flat_views(dev):
    """after model.zero_grad() (set_to_none=True is torch's default) backward writes FRESH .grad tensors; the
    step must use them, not the stale zeroed flat buffer."""
    from gtos_b200.optim import FlatAdam
    gen = torch.Generator().manual_seed(SEED + 5)
    lin_a, lin_b = torch.nn.Linear(16, 8).to(dev), torch.nn.Linear(16, 8).to(dev)
    lin_b.load_state_dict(lin_a.state_dict())
    x = torch.randn(4, 16, generator=gen).to(dev)
    opt_a = FlatAdam(list(lin_a.named_parameters()), lr=0.01, max_norm=None)
    opt_b = FlatAdam(list(lin_b.named_parameters()), lr=0.01, max_norm=None)
    opt_a.zero_grad()
    lin_a(x).pow(2).sum().backward()                      # accumulates into the views
    lin_b.zero_grad()                                     # torch default: grads become None
    lin_b(x).pow(2).sum().backward()                      # fresh .grad tensors
    assert lin_b.weight.grad.data_ptr() != opt_b.bucket.flat.data_ptr()
    opt_a.step()
    opt_b.step()
    assert torch.equal(lin_a.weight, lin_b.weight) and torch.equal(lin_a.bias, lin_b.bias)
    assert lin_b.weight.grad.data_ptr() == opt_b.bucket.flat.data_ptr()      # re-attached


def test_flat_adam_large_buffer_property(dev):
    """38.6 M parameters (the reference model's size): one step equals the elementwise formula on a random sample"""
    from gtos_b200.optim import FlatAdam
    n = 38_600_003
    gen = torch.Generator(device=dev).manual_seed(SEED)
    w = torch.nn.Parameter(torch.randn(n - 1000, device=dev, generator=gen))
    b = torch.nn.Parameter(torch.randn(1000, device=dev, generator=gen))
    w0, b0 = w.detach().clone(), b.detach().clone()
    opt = FlatAdam([("w.weight", w), ("w.bias", b)], lr=0.01, max_norm=1.0)
    w.grad.copy_(torch.randn(n - 1000, device=dev, generator=gen) * 1e-3)
    b.grad.copy_(torch.randn(1000, device=dev, generator=gen) * 1e-3)
    gw, gb = w.grad.clone(), b.grad.clone()
    opt.step()
    total = torch.sqrt(gw.double().pow(2).sum() + gb.double().pow(2).sum()).float()
    assert float(opt.grad_norm()) == pytest.approx(float(total), rel=1e-5)
    coef = torch.clamp(1.0 / (total + 1e-6), max=1.0)
    for p0, g0, p1, wd in ((w0, gw, w, 1e-4), (b0, gb, b, 0.0)):
        gg = g0 * coef
        m, v = 0.1 * gg, 0.001 * gg * gg
        ref = p0 - 0.01 * (m / (v.sqrt() + 1e-6) + wd * p0)
        assert rel_err(p1, ref) < 1e-5
