"""Golden vectors for one-token-at-a-time decoding from the reference's OWN Generator.decode_step
(generator/generator.py:120-167), with hypothesis re-parenting done exactly as search.py:72-76,139-143 does it
(index_select of every state / memory tensor).  Build container only (needs /root/reference):
    python tests/golden/make_golden_decode.py      -> tests/golden/golden_decode_v1.pt

The token embedding front-end (TokenEncoder + position + LayerNorm, generator.py:131-132) is outside the hot path
(SURVEY.md §2.1): its output `token_repr` is recorded as an input.  The full log-prob table is recovered by asking
decode_step for top-k with k = table width.  Shims: the two of make_golden.py (no numeric effect).
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/generator")
warnings.filterwarnings("ignore")
import transformer as ref_tf                      # noqa: E402
import generator as ref_gen                       # noqa: E402

_orig_qkv = ref_tf.MultiheadAttention.in_proj_qkv
ref_tf.MultiheadAttention.in_proj_qkv = lambda self, q: tuple(t.clone() for t in _orig_qkv(self, q))
SEED = 19940117


class FakeVocab:
    def __init__(self, size):
        self.size, self.padding_idx, self.unk_idx = size, 0, 1

    def idx2token(self, i):
        return f"tok{i}"

    def token2idx(self, s):
        return int(s[3:])


def boost(module, factor):
    with torch.no_grad():
        for n, p in module.named_parameters():
            if p.dim() >= 2 and "layer_norm" not in n:
                p.mul_(factor)
            elif "bias" in n:
                p.normal_(0, 0.1)
            elif "layer_norm.weight" in n:
                p.add_(torch.randn_like(p) * 0.1)


def run_case(name, D, F, H, snt_layers, inf_layers, tok_dim, V, S, B, plan):
    torch.manual_seed(SEED + len(name))
    vocabs = {k: FakeVocab(n) for k, n in dict(concept=23, token=31, predictable_token=V, token_char=17, concept_char=17,
                                               relation=12).items()}
    m = ref_gen.Generator(vocabs, 8, 12, 8, tok_dim, [(3, 10)], 14, 14, 6, 8, 1, D, F, H, 0.2, snt_layers, 1, inf_layers,
                          None, "cpu")
    boost(m.snt_encoder, 3.0)
    boost(m.decoder, 3.0)
    m.eval()
    gen = torch.Generator().manual_seed(SEED + 7)
    graph_state = torch.randn(S, B, D, generator=gen)
    lens = torch.randint(2, S + 1, (B,), generator=gen)
    lens[0] = S
    gmask = torch.arange(S).unsqueeze(1) >= lens.unsqueeze(0)
    probe = torch.tanh(torch.randn(1, B, D, generator=gen))
    cp_seq = torch.randint(2, V + 5, (S, B), generator=gen)
    W = max(V, int(cp_seq.max()) + 1)
    mem = dict(graph_state=graph_state, graph_padding_mask=gmask, probe=probe, cp_seq=cp_seq)
    state, steps = {}, []
    src = torch.arange(B)
    with torch.no_grad():
        for t, parent in enumerate(plan):
            if parent is not None:
                parent = torch.tensor(parent)
                state = {k: v.index_select(1, parent) for k, v in state.items()}             # search.py:72-76
                src = src.index_select(0, parent)
            hyp = src.numel()
            cur = {k: v.index_select(1, src) for k, v in mem.items()}                        # search.py:139-143
            cur["local_idx2token"] = [dict() for _ in range(hyp)]
            tok = torch.randint(2, vocabs["token"].size, (1, hyp), generator=gen)
            tok_char = torch.randint(2, vocabs["token_char"].size, (1, hyp, 5), generator=gen)
            token_repr = m.embed_scale * m.token_encoder(tok, tok_char) + m.token_position(tok, t)   # generator.py:131
            token_repr = m.token_embed_layer_norm(token_repr)
            state, results = m.decode_step((tok, tok_char), state, cur, t, W)
            ll = torch.empty(hyp, W)
            for h, res in enumerate(results):
                for s, sc in res:
                    ll[h, int(s[3:])] = sc
            steps.append(dict(token_repr=token_repr.clone(), src=src.clone(), parent=parent, ll=ll))
    sd = {k: v.clone() for k, v in m.state_dict().items() if k.startswith(("snt_encoder.", "decoder."))}
    return dict(cfg=dict(D=D, F=F, H=H, snt_layers=snt_layers, inference_layers=inf_layers, tok_dim=tok_dim, V=V, S=S, B=B, W=W),
                state=sd, mem=mem, steps=steps)


def main():
    out = {
        "one_snt_layer": run_case("a", 32, 64, 4, 1, 2, 24, 19, 6, 3,
                                  [None, [0, 0, 1, 2, 2], [4, 0, 0, 3, 1, 1], [5, 2], [1, 1, 0, 0]]),
        "two_snt_layers": run_case("bb", 64, 128, 8, 2, 3, 40, 50, 9, 4,
                                   [None, [0, 1, 1, 2, 3, 3, 3], [6, 5, 4, 3, 2, 1, 0], [0, 0, 0, 6]]),
    }
    torch.save(out, os.path.join(HERE, "golden_decode_v1.pt"))
    for k, v in out.items():
        print(k, [tuple(s["ll"].shape) for s in v["steps"]], float(v["steps"][-1]["ll"].exp().sum(1)[0]))


if __name__ == "__main__":
    main()
