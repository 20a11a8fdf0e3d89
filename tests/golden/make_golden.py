"""Generate golden vectors by running the UNMODIFIED reference modules (CPU, fp32).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/golden_v1.pt.  The reference has no golden vectors of its own
(SURVEY.md §4), so these pin the oracle (tests/test_oracle_golden.py) and, through it, the
CUDA path.  Seed 19940117 is the reference's own (generator/train.py:98).

Compatibility shims for torch 2.x, neither changes numerics:
  * causal mask passed as bool (reference builds uint8, transformer.py:212)
  * MultiheadAttention.in_proj_qkv returns clones (in-place `q *= scaling` on a chunk view,
    transformer.py:120)
"""
import os
import sys

import torch

REF = "/root/reference/generator"
sys.path.insert(0, REF)
import graph_transformer as ref_gt   # noqa: E402
import transformer as ref_tf         # noqa: E402
import encoder as ref_enc            # noqa: E402
import decoder as ref_dec            # noqa: E402

_orig_qkv = ref_tf.MultiheadAttention.in_proj_qkv
ref_tf.MultiheadAttention.in_proj_qkv = lambda self, q: tuple(t.clone() for t in _orig_qkv(self, q))

SEED = 19940117


class FakeVocab:
    def __init__(self, size):
        self.size, self.padding_idx, self.unk_idx = size, 0, 1

    def idx2token(self, i):
        return f"tok{i}"


def boost(module, factor):
    """Inflate std-0.02 weights so softmaxes are peaky and errors cannot hide."""
    with torch.no_grad():
        for n, p in module.named_parameters():
            if p.dim() >= 2 and "layer_norm" not in n:
                p.mul_(factor)
            elif "bias" in n:
                p.normal_(0, 0.1)
            elif "layer_norm.weight" in n:
                p.add_(torch.randn_like(p) * 0.1)


def grads_of(module, inputs, loss):
    names = [n for n, _ in module.named_parameters()]
    params = [p for _, p in module.named_parameters()]
    gs = torch.autograd.grad(loss, list(inputs) + params, allow_unused=True)
    gi = [g.detach().clone() if g is not None else None for g in gs[:len(inputs)]]
    gp = {n: g.detach().clone() for n, g in zip(names, gs[len(inputs):]) if g is not None}
    return gi, gp


def pad_mask(lens, n):
    return torch.arange(n).unsqueeze(1) >= torch.tensor(lens).unsqueeze(0)    # [n, B] True = pad


def main():
    torch.manual_seed(SEED)
    torch.set_num_threads(4)
    G = {}

    # ---- 1. RelationMultiheadAttention -------------------------------------------------
    N, B, D, H = 7, 3, 32, 4
    m = ref_gt.RelationMultiheadAttention(D, H, dropout=0.0)
    boost(m, 8.0)
    x = torch.randn(N, B, D, requires_grad=True)
    rel = (torch.randn(N, N, B, D) * 0.5).requires_grad_()
    mask = pad_mask([7, 5, 6], N)
    out, w = m(x, x, x, rel, key_padding_mask=mask, need_weights=True)
    wo, ww = torch.randn_like(out), torch.randn_like(w)
    gi, gp = grads_of(m, [x, rel], (out * wo).sum() + (w * ww).sum())
    G["rel_mha"] = dict(cfg=dict(N=N, B=B, D=D, H=H), state=m.state_dict(), x=x.detach(), rel=rel.detach(),
                        mask=mask, out=out.detach(), w=w.detach(), wo=wo, ww=ww, gx=gi[0], grel=gi[1], gp=gp)

    # ---- 2. GraphTransformer (2 layers), and the stacked attention weights -----------
    N, B, D, H, Fd, L = 9, 4, 32, 4, 64, 2
    m = ref_gt.GraphTransformer(L, D, Fd, H, 0.0)
    boost(m, 5.0)
    x = torch.randn(N, B, D, requires_grad=True)
    rel = (torch.randn(N, N, B, D) * 0.5).requires_grad_()
    mask = pad_mask([9, 4, 7, 6], N)
    out = m(x, rel, self_padding_mask=mask)
    wo = torch.randn_like(out)
    gi, gp = grads_of(m, [x, rel], (out * wo).sum())
    with torch.no_grad():
        attn = m.get_attn_weights(x, rel, self_padding_mask=mask)
    G["graph_transformer"] = dict(cfg=dict(N=N, B=B, D=D, H=H, F=Fd, L=L), state=m.state_dict(), x=x.detach(),
                                  rel=rel.detach(), mask=mask, out=out.detach(), wo=wo, gx=gi[0], grel=gi[1],
                                  gp=gp, attn=attn)

    # ---- 3. MultiheadAttention: causal self-attn and cross-attn with weights ----------
    T, S, B, D, H = 6, 8, 3, 32, 4
    m = ref_tf.MultiheadAttention(D, H, dropout=0.0)
    boost(m, 8.0)
    q = torch.randn(T, B, D, requires_grad=True)
    tmask = pad_mask([6, 4, 5], T)
    cm = torch.ones(T, T, dtype=torch.bool).triu_(1)
    out, w = m(q, q, q, key_padding_mask=tmask, attn_mask=cm, need_weights=True)
    wo = torch.randn_like(out)
    gi, gp = grads_of(m, [q], (out * wo).sum())
    G["mha_self"] = dict(cfg=dict(T=T, B=B, D=D, H=H), state=m.state_dict(), q=q.detach(), tmask=tmask, cm=cm,
                         out=out.detach(), w=w.detach(), wo=wo, gq=gi[0], gp=gp)
    mem = torch.randn(S, B, D, requires_grad=True)
    smask = pad_mask([8, 3, 6], S)
    out, w = m(q, mem, mem, key_padding_mask=smask, need_weights=True)
    ww = torch.randn_like(w)
    gi, gp = grads_of(m, [q, mem], (out * wo).sum() + (w * ww).sum())
    G["mha_cross"] = dict(cfg=dict(T=T, S=S, B=B, D=D, H=H), state=m.state_dict(), q=q.detach(), mem=mem.detach(),
                          smask=smask, out=out.detach(), w=w.detach(), wo=wo, ww=ww, gq=gi[0], gmem=gi[1], gp=gp)

    # ---- 4. Transformer with external memory (2 layers), kv != x ----------------------
    T, S, B, D, H, Fd, L = 6, 8, 3, 32, 4, 64, 2
    m = ref_tf.Transformer(L, D, Fd, H, 0.0, with_external=True)
    boost(m, 5.0)
    x = torch.randn(T, B, D, requires_grad=True)
    kv = torch.randn(T, B, D, requires_grad=True)
    mem = torch.randn(S, B, D, requires_grad=True)
    out = m(x, kv=kv, self_padding_mask=tmask, self_attn_mask=cm, external_memories=mem, external_padding_mask=smask)
    wo = torch.randn_like(out)
    gi, gp = grads_of(m, [x, kv, mem], (out * wo).sum())
    G["transformer_ext"] = dict(cfg=dict(T=T, S=S, B=B, D=D, H=H, F=Fd, L=L), state=m.state_dict(), x=x.detach(),
                                kv=kv.detach(), mem=mem.detach(), tmask=tmask, cm=cm, smask=smask, out=out.detach(),
                                wo=wo, gx=gi[0], gkv=gi[1], gmem=gi[2], gp=gp)

    # ---- 5. RelationEncoder (packed 2-layer bi-GRU) -----------------------------------
    R, Lmax, rel_dim, hid, D, V = 23, 4, 12, 16, 32, 19
    m = ref_enc.RelationEncoder(FakeVocab(V), rel_dim, D, hid, 2, 0.0)
    boost(m.rel_embed, 20.0)
    lengths = torch.randint(1, Lmax + 1, (R,))
    lengths[0] = Lmax
    tokens = torch.randint(2, V, (Lmax, R))
    tokens = tokens.masked_fill(torch.arange(Lmax).unsqueeze(1) >= lengths.unsqueeze(0), 0)
    out = m(tokens, lengths)
    wo = torch.randn_like(out)
    _, gp = grads_of(m, [], (out * wo).sum())
    G["relation_encoder"] = dict(cfg=dict(R=R, Lmax=Lmax, rel_dim=rel_dim, hid=hid, D=D, V=V), state=m.state_dict(),
                                 tokens=tokens, lengths=lengths, out=out.detach(), wo=wo, gp=gp)

    # ---- 6. DecodeLayer: training loss and work=True log-probs ------------------------
    T, S, B, D, H, Fd, L, V, tok_dim = 6, 8, 3, 32, 4, 64, 2, 41, 24
    vocabs = {"predictable_token": FakeVocab(V)}
    m = ref_dec.DecodeLayer(vocabs, L, D, Fd, H, tok_dim, 0, 0.0)
    boost(m, 5.0)
    probe = torch.randn(1, B, D).expand(T, B, D).clone().requires_grad_()
    graph = torch.randn(S, B, D, requires_grad=True)
    snt = torch.randn(T, B, D, requires_grad=True)
    copy_seq = torch.randint(2, V + 5, (S, B))
    target = torch.randint(2, V + 5, (T, B)).masked_fill(tmask, 0)
    loss = m(probe, graph, snt, smask, tmask, cm, copy_seq, target=target)
    gi, gp = grads_of(m, [probe, graph, snt], loss)
    with torch.no_grad():
        ll = m(probe, graph, snt, smask, tmask, cm, copy_seq, work=True)
    G["decode_layer"] = dict(cfg=dict(T=T, S=S, B=B, D=D, H=H, F=Fd, L=L, V=V, tok_dim=tok_dim), state=m.state_dict(),
                             probe=probe.detach(), graph=graph.detach(), snt=snt.detach(), smask=smask, tmask=tmask,
                             cm=cm, copy_seq=copy_seq, target=target, loss=loss.detach(), ll=ll, gprobe=gi[0],
                             ggraph=gi[1], gsnt=gi[2], gp=gp)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.pt")
    torch.save(G, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
