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
