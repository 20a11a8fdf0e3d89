"""The hot path of ``Generator.forward`` assembled from the drop-in modules, exactly as the reference wires
them (generator/generator.py:71-94 encode_step, :169-182 forward), minus the char-CNN embedding front-end
(out of scope): inputs are the post-LayerNorm node features and token features.

This is the public API bench.py and the end-to-end tests drive; a reference user gets the same thing by
putting ``gtos_b200/dropin`` on sys.path in front of generator/ (INTEGRATION.md).
"""
import math

import torch
from torch import nn

from . import ops
from .decoder import DecodeLayer
from .encoder import RelationEncoder
from .graph_transformer import GraphTransformer
from .transformer import Transformer


class HotPathConfig:
    """generator/train.sh:2-40 defaults (BASELINE.json config 2)."""

    def __init__(self, embed_dim=512, ff_embed_dim=1024, num_heads=8, graph_layers=4, snt_layers=1,
                 inference_layers=3, rel_dim=100, rnn_hidden_size=256, rnn_num_layers=2, concept_dim=300,
                 dropout=0.2, vocab_size=10000, rel_vocab_size=206):
        self.__dict__.update(locals())
        del self.__dict__["self"]


class _V:
    def __init__(self, size):
        self.size, self.padding_idx, self.unk_idx = size, 0, 1


class HotPath(nn.Module):
    def __init__(self, cfg: HotPathConfig):
        super().__init__()
        c = self.cfg = cfg
        vocabs = {"relation": _V(c.rel_vocab_size), "predictable_token": _V(c.vocab_size)}
        self.vocabs = vocabs
        self.relation_encoder = RelationEncoder(vocabs["relation"], c.rel_dim, c.embed_dim, c.rnn_hidden_size,
                                                c.rnn_num_layers, c.dropout)
        self.graph_encoder = GraphTransformer(c.graph_layers, c.embed_dim, c.ff_embed_dim, c.num_heads, c.dropout)
        self.snt_encoder = Transformer(c.snt_layers, c.embed_dim, c.ff_embed_dim, c.num_heads, c.dropout,
                                       with_external=True)
        self.decoder = DecodeLayer(vocabs, c.inference_layers, c.embed_dim, c.ff_embed_dim, c.num_heads,
                                   c.concept_dim, c.rel_dim, c.dropout)
        self._prep_plan = ops.WeightPrepPlan()
        self.banked_relation = True      # False: dense `relation_bank[idx]` exactly as generator.py:79 builds it
        self.probe_generator = nn.Linear(c.embed_dim, c.embed_dim)
        nn.init.normal_(self.probe_generator.weight, std=0.02)
        nn.init.constant_(self.probe_generator.bias, 0.)

    def encode(self, batch):
        """generator.py:76-94: relation bank -> dense relation -> graph encoder -> probe / node states."""
        bank = self.relation_encoder(batch["relation_bank"], batch["relation_length"])
        idx = batch["relation"]
        if self.banked_relation:
            # §8 f-0: keep relation = bank[idx] factorised (the 2-line caller change): no fp32
            # [N,N,B,D] tensor, bank-row GEMMs in the backward
            relation = ops.BankedRelation(bank, idx)
        else:
            relation = ops.bank_gather(bank, idx)                                   # generator.py:79 (+ bf16 copy)
        h = self.graph_encoder(batch["x"], relation, self_padding_mask=batch["node_mask"])
        probe = torch.tanh(self.probe_generator(h[:1]))
        return h[1:], batch["node_mask"][1:], probe

    def forward(self, batch):
        """generator.py:169-182 -> scalar loss."""
        with self._prep_plan.step():             # weight operand copies issued ahead, beside the RelationEncoder
            return self._forward(batch)

    def _forward(self, batch):
        concept_repr, concept_mask, probe = self.encode(batch)
        token_repr = batch["token_repr"]
        attn_mask = batch["causal_mask"]
        token_repr = self.snt_encoder(token_repr, self_padding_mask=batch["token_mask"], self_attn_mask=attn_mask,
                                      external_memories=concept_repr, external_padding_mask=concept_mask)
        probe = probe.expand_as(token_repr)
        return self.decoder(probe, concept_repr, token_repr, concept_mask, batch["token_mask"], attn_mask,
                            batch["copy_seq"], target=batch["target"])


BATCH_KEYS = ("relation_bank", "relation_length", "relation", "x", "node_mask", "token_repr", "token_mask",
              "copy_seq", "target", "causal_mask")


def batch_tensors(g):
    """synthetic.make_batch output -> the dict HotPath consumes (host tensors)."""
    T = g["T"]
    out = {k: g[k] for k in BATCH_KEYS if k in g}
    out["causal_mask"] = torch.ones(T, T, dtype=torch.bool).triu_(1)
    return out
