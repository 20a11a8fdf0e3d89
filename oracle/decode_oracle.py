"""CPU oracle for one beam-search decode step exactly as the reference computes it: no caches --
every step re-runs the sentence layer(s) on the new token with kv = the whole prefix of that layer's inputs, and the
DecodeLayer over the whole prefix of token states, for every hypothesis (generator/generator.py:120-167).

TEST INFRASTRUCTURE ONLY: nothing under ``gtos_b200/`` imports this.

State per hypothesis set, as the reference keeps it (generator.py:133-149): `token_repr_{l}` = inputs of sentence layer l
for positions 0..t ([t+1, Hyp, D]) and `token_state` = outputs of the last sentence layer ([t+1, Hyp, D]); a beam
re-parenting is `index_select(1, parent)` on every state (search.py:72-76).  The arithmetic is oracle/gtos_oracle.py's
(pinned by tests/golden/golden_v1.pt); this file only restates the wiring, and tests/golden/make_golden_decode.py pins
that wiring against the reference's own Generator.decode_step driven by the reference's search_by_batch.
"""
import torch

from . import gtos_oracle as O


def decode_step(P, cfg, mem, token_repr, state, src_index, parent):
    """P: parameter dict with `snt_encoder.` / `decoder.` prefixes (HotPath / Generator state_dict names).
    mem: dict(graph_state [S,B,D], graph_padding_mask [S,B], probe [1,B,D], copy_seq [S,B]).
    token_repr [1,Hyp,D]; state: dict from the previous call (or {}); src_index, parent: int64 [Hyp] (parent None at t=0).
    Returns (log-prob table [Hyp, W], new state)."""
    if parent is not None:
        state = {k: v.index_select(1, parent) for k, v in state.items()}            # search.py:72-76
    graph = mem["graph_state"].index_select(1, src_index)                           # search.py:139-143
    gmask = mem["graph_padding_mask"].index_select(1, src_index)
    probe = mem["probe"].index_select(1, src_index)
    copy_seq = mem["copy_seq"].index_select(1, src_index)
    new_state = {}
    x = token_repr
    for l in range(cfg.snt_layers):                                                 # generator.py:133-142
        name = f"token_repr_{l}"
        kv = torch.cat([state[name], x], 0) if name in state else x
        new_state[name] = kv
        x, _, _ = O.transformer_layer(P, f"snt_encoder.layers.{l}.", x, cfg.num_heads, kv=kv, external_memories=graph,
                                      external_padding_mask=gmask, with_external=True)
    ts = torch.cat([state["token_state"], x], 0) if "token_state" in state else x   # generator.py:143-149
    new_state["token_state"] = ts
    ll = O.decode_layer(P, "decoder.", probe, graph, ts, gmask, None, None, copy_seq, cfg.inference_layers, cfg.num_heads,
                        0, work=True)                                               # generator.py:150
    return ll.squeeze(0), new_state
