"""Synthetic inputs for the reference's OWN `Generator` (generator/generator.py): fake vocabularies with the methods the
model and search.py call, and the batch dictionary `data.batchify` would emit (data.py:252-266), built from
gtos_b200.synthetic graphs.  Used by tests that run the unmodified caller (Generator.forward / Generator.work /
search_by_batch) over (a) the reference modules on the CPU and (b) the gtos_b200 drop-in modules on the GPU."""
import torch

from gtos_b200 import synthetic

PAD, UNK, STR, END = "<PAD>", "<UNK>", "<STR>", "<END>"


class Vocab:
    """data.Vocab's public surface (data.py:12-57) over a fixed token list"""

    def __init__(self, prefix, size):
        self._idx2token = [PAD, UNK, STR, END] + [f"{prefix}{i}" for i in range(4, size)]
        self._token2idx = {t: i for i, t in enumerate(self._idx2token)}

    @property
    def size(self):
        return len(self._idx2token)

    @property
    def unk_idx(self):
        return 1

    @property
    def padding_idx(self):
        return 0

    def idx2token(self, x):
        if isinstance(x, list):
            return [self.idx2token(i) for i in x]
        return self._idx2token[x]

    def token2idx(self, x):
        if isinstance(x, list):
            return [self.token2idx(i) for i in x]
        return self._token2idx.get(x, self.unk_idx)


def make_vocabs(V_pred=45):
    rel = synthetic.RelVocab(100)
    rel.token2idx = lambda x: 0
    return {"concept": Vocab("c", 50), "concept_char": Vocab("", 30), "relation": rel, "token": Vocab("w", V_pred + 15),
            "token_char": Vocab("", 30), "predictable_token": Vocab("w", V_pred)}


GEN_ARGS = dict(word_char_dim=16, word_dim=300, concept_char_dim=16, concept_dim=300, cnn_filters=[(3, 64)],
                char2word_dim=64, char2concept_dim=64, rel_dim=100, rnn_hidden_size=64, rnn_num_layers=2,
                embed_dim=128, ff_embed_dim=256, num_heads=8, snt_layers=1, graph_layers=2, inference_layers=2)


def build_generator(gen_module, vocabs, dropout, device):
    a = GEN_ARGS
    return gen_module.Generator(vocabs, a["word_char_dim"], a["word_dim"], a["concept_char_dim"], a["concept_dim"],
                                a["cnn_filters"], a["char2word_dim"], a["char2concept_dim"], a["rel_dim"],
                                a["rnn_hidden_size"], a["rnn_num_layers"], a["embed_dim"], a["ff_embed_dim"], a["num_heads"],
                                dropout, a["snt_layers"], a["graph_layers"], a["inference_layers"], None, device)


def boost(module, factor, seed=1):
    """inflate the std-0.02 weights of the hot-path modules so softmaxes are peaky and errors cannot hide"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in module.named_parameters():
            if not n.startswith(("graph_encoder.", "snt_encoder.", "decoder.", "relation_encoder.out_proj", "probe_generator")):
                continue
            if p.dim() >= 2 and "layer_norm" not in n:
                p.mul_(factor)
            elif n.endswith("bias") and "layer_norm" not in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)


def make_data(vocabs, B=6, n_max=12, T=9, seed=19940117, chars=7, eval_paths=0):
    """the dictionary of data.py:252-266.  eval_paths = K > 0: the evaluation layout (relation [N,N,B,K], bank row 0 =
    <PAD>, data.py:176-225) with up to K paths per pair."""
    g = synthetic.make_graph_batch(B, n_max, max_path_len=4, seed=seed)
    gen = torch.Generator().manual_seed(seed)
    N = g["N"]
    counts = g["node_counts"]
    node_pad = torch.arange(N).unsqueeze(1) >= (counts + 1).unsqueeze(0)                     # [N,B]
    Vc, Vp, Vt = vocabs["concept"].size, vocabs["predictable_token"].size, vocabs["token"].size
    concept = torch.randint(4, Vc, (N, B), generator=gen).masked_fill(node_pad, 0)
    concept[0] = 2                                                                           # <CLS> slot
    concept_char = torch.randint(4, vocabs["concept_char"].size, (N, B, chars), generator=gen)
    concept_char = concept_char.masked_fill(node_pad.unsqueeze(-1), 0)
    depth = torch.randint(0, 8, (N, B), generator=gen).masked_fill(node_pad, 0)
    t_len = torch.randint(T // 2, T + 1, (B,), generator=gen)
    t_len[0] = T
    tok_pad = torch.arange(T).unsqueeze(1) >= t_len.unsqueeze(0)
    token_in = torch.randint(4, Vt, (T, B), generator=gen).masked_fill(tok_pad, 0)
    token_in[0] = 2                                                                          # <STR>
    token_char_in = torch.randint(4, vocabs["token_char"].size, (T, B, chars), generator=gen)
    token_char_in = token_char_in.masked_fill(tok_pad.unsqueeze(-1), 0)
    n_ext = 4
    cp_seq = torch.randint(2, Vp + n_ext, (N - 1, B), generator=gen).masked_fill(node_pad[1:], 0)
    token_out = torch.randint(2, Vp + n_ext, (T, B), generator=gen)
    token_out = torch.minimum(token_out, cp_seq.max()).masked_fill(tok_pad, 0)        # targets inside the extended vocabulary
    local_idx2token = [{Vp + k: f"cp{b}_{k}" for k in range(n_ext)} for b in range(B)]
    local_token2idx = [{v: k for k, v in d.items()} for d in local_idx2token]
    rel, bank, length = g["relation"], g["relation_bank"], g["relation_length"]
    if eval_paths:
        K = eval_paths
        R = bank.shape[1]
        bank = torch.cat([torch.zeros(bank.shape[0], 1, dtype=bank.dtype), bank], dim=1)    # row 0 = <PAD> = (pad_idx,)
        length = torch.cat([torch.ones(1, dtype=length.dtype), length])
        first = rel + 1
        inside = ~(node_pad.unsqueeze(0) | node_pad.unsqueeze(1))                            # [N,N,B] both nodes real
        first = first.masked_fill(~inside, 0)
        more = torch.randint(4, R + 1, (N, N, B, K - 1), generator=gen)
        keep = (torch.rand(N, N, B, K - 1, generator=gen) < 0.4) & inside.unsqueeze(-1)
        keep[0] = False
        keep[:, 0] = False                                                                   # <CLS> pairs have one path
        rel = torch.cat([first.unsqueeze(-1), more.masked_fill(~keep, 0)], dim=-1)
    return {"concept": concept, "concept_char": concept_char, "concept_depth": depth, "relation": rel,
            "relation_bank": bank, "relation_length": length, "local_idx2token": local_idx2token,
            "local_token2idx": local_token2idx, "token_in": token_in, "token_char_in": token_char_in,
            "token_out": token_out, "cp_seq": cp_seq}


def to_device(data, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
