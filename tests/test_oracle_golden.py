"""Pin the CPU oracle (oracle/gtos_oracle.py) against golden vectors produced by the real
reference modules (tests/golden/make_golden.py).  CPU only, fp32, tolerance 2e-5 relative."""
import torch

from conftest import rel_err
from oracle import gtos_oracle as O

TOL = 2e-5


def _params(state):
    return {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in state.items()}


def _check_grads(P, gp, loss_inputs, grads_in, loss):
    names = list(gp.keys())
    gs = torch.autograd.grad(loss, loss_inputs + [P[n] for n in names], allow_unused=True)
    for g, ref in zip(gs[:len(loss_inputs)], grads_in):
        assert rel_err(g, ref) < TOL
    for n, g in zip(names, gs[len(loss_inputs):]):
        assert g is not None, n
        assert rel_err(g, gp[n]) < TOL, n


def test_rel_mha(golden):
    g = golden["rel_mha"]
    P = _params(g["state"])
    x, rel = g["x"].clone().requires_grad_(), g["rel"].clone().requires_grad_()
    out, w = O.rel_mha(P, "", x, x, x, rel, g["cfg"]["H"], g["mask"], need_weights=True)
    assert rel_err(out, g["out"]) < TOL and rel_err(w, g["w"]) < TOL
    _check_grads(P, g["gp"], [x, rel], [g["gx"], g["grel"]], (out * g["wo"]).sum() + (w * g["ww"]).sum())


def test_graph_transformer(golden):
    g = golden["graph_transformer"]
    c = g["cfg"]
    P = _params(g["state"])
    x, rel = g["x"].clone().requires_grad_(), g["rel"].clone().requires_grad_()
    out = O.graph_transformer(P, "", x, rel, c["L"], c["H"], self_padding_mask=g["mask"])
    assert rel_err(out, g["out"]) < TOL
    _check_grads(P, g["gp"], [x, rel], [g["gx"], g["grel"]], (out * g["wo"]).sum())
    attn = O.graph_transformer(P, "", x, rel, c["L"], c["H"], self_padding_mask=g["mask"], return_weights=True)
    assert rel_err(attn, g["attn"]) < TOL


def test_mha_self_and_cross(golden):
    g = golden["mha_self"]
    P = _params(g["state"])
    q = g["q"].clone().requires_grad_()
    out, w = O.mha(P, "", q, q, q, g["cfg"]["H"], g["tmask"], g["cm"], need_weights=True)
    assert rel_err(out, g["out"]) < TOL and rel_err(w, g["w"]) < TOL
    _check_grads(P, g["gp"], [q], [g["gq"]], (out * g["wo"]).sum())
    g = golden["mha_cross"]
    P = _params(g["state"])
    q, mem = g["q"].clone().requires_grad_(), g["mem"].clone().requires_grad_()
    out, w = O.mha(P, "", q, mem, mem, g["cfg"]["H"], g["smask"], None, need_weights=True)
    assert rel_err(out, g["out"]) < TOL and rel_err(w, g["w"]) < TOL
    _check_grads(P, g["gp"], [q, mem], [g["gq"], g["gmem"]], (out * g["wo"]).sum() + (w * g["ww"]).sum())


def test_transformer_external(golden):
    g = golden["transformer_ext"]
    c = g["cfg"]
    P = _params(g["state"])
    x, kv, mem = (g[k].clone().requires_grad_() for k in ("x", "kv", "mem"))
    out = O.transformer(P, "", x, c["L"], c["H"], kv=kv, self_padding_mask=g["tmask"], self_attn_mask=g["cm"],
                        external_memories=mem, external_padding_mask=g["smask"], with_external=True)
    assert rel_err(out, g["out"]) < TOL
    _check_grads(P, g["gp"], [x, kv, mem], [g["gx"], g["gkv"], g["gmem"]], (out * g["wo"]).sum())


def test_relation_encoder(golden):
    g = golden["relation_encoder"]
    P = _params(g["state"])
    out = O.relation_encoder(P, "", g["tokens"], g["lengths"], num_layers=2)
    assert rel_err(out, g["out"]) < TOL
    _check_grads(P, g["gp"], [], [], (out * g["wo"]).sum())


def test_decode_layer(golden):
    g = golden["decode_layer"]
    c = g["cfg"]
    P = _params(g["state"])
    probe, graph, snt = (g[k].clone().requires_grad_() for k in ("probe", "graph", "snt"))
    loss = O.decode_layer(P, "", probe, graph, snt, g["smask"], g["tmask"], g["cm"], g["copy_seq"], c["L"], c["H"],
                          0, target=g["target"])
    assert rel_err(loss, g["loss"]) < TOL
    _check_grads(P, g["gp"], [probe, graph, snt], [g["gprobe"], g["ggraph"], g["gsnt"]], loss)
    ll = O.decode_layer(P, "", probe, graph, snt, g["smask"], g["tmask"], g["cm"], g["copy_seq"], c["L"], c["H"],
                        0, work=True)
    assert rel_err(ll, g["ll"]) < TOL


# ---- optimizer step (f-4) and one-token decode step (f-1) ----------------------------------
import os                                           # noqa: E402
import types                                        # noqa: E402

import pytest                                       # noqa: E402

GDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_optim_oracle_matches_reference_adam_and_clip():
    """oracle/optim_oracle.py vs the reference's AdamWeightDecayOptimizer + clip_grad_norm_ + update_lr, 4 steps"""
    from oracle import optim_oracle as OO
    g = torch.load(os.path.join(GDIR, "golden_optim_v1.pt"), map_location="cpu", weights_only=False)
    params = {n: v.clone() for n, v in g["init"].items()}
    m = {n: torch.zeros_like(v) for n, v in params.items()}
    v2 = {n: torch.zeros_like(v) for n, v in params.items()}
    wd = {n: 0.0 if OO.no_decay(n) else 1e-4 for n in params}
    for k, st in enumerate(g["steps"], start=1):
        lr = OO.update_lr(g["embed_size"], k, g["warmup"])
        assert lr == pytest.approx(st["lr"], rel=1e-12)
        total, _ = OO.clip_coef(list(st["grads"].values()), 1.0)
        assert float(total) == pytest.approx(st["total_norm"], rel=1e-6)
        OO.adam_step(params, st["grads"], m, v2, lr, wd, eps=1e-6, max_norm=1.0)
        for n in params:
            assert rel_err(params[n], st["params"][n]) < 1e-6, (k, n)
            assert rel_err(m[n], st["exp_avg"][n]) < 1e-6 and rel_err(v2[n], st["exp_avg_sq"][n]) < 1e-6, (k, n)


@pytest.mark.parametrize("case", ["one_snt_layer", "two_snt_layers"])
def test_decode_step_oracle_matches_reference_generator(case):
    """oracle/decode_oracle.py vs Generator.decode_step driven with search.py's re-parenting, every step's full table"""
    from oracle import decode_oracle as DO
    g = torch.load(os.path.join(GDIR, "golden_decode_v1.pt"), map_location="cpu", weights_only=False)[case]
    c = g["cfg"]
    cfg = types.SimpleNamespace(snt_layers=c["snt_layers"], inference_layers=c["inference_layers"], num_heads=c["H"])
    mem = dict(g["mem"], copy_seq=g["mem"]["cp_seq"])
    state = {}
    with torch.no_grad():
        for st in g["steps"]:
            ll, state = DO.decode_step(g["state"], cfg, mem, st["token_repr"], state, st["src"], st["parent"])
            assert ll.shape == st["ll"].shape
            assert (ll - st["ll"]).abs().max() < 2e-4          # log-probs down to log(1e-12) = -27.6
            assert rel_err(ll.exp(), st["ll"].exp()) < TOL
