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
    """`modules`: where the four hot-path classes come from - None = this package (the B200 implementation); tests and
    bench.py's CPU arm pass the reference's own modules (same constructor calls as generator.py:27-42, same wiring), so
    both arms run literally the same caller code.

    `relation_mode`: how `relation = bank[idx]` reaches the graph encoder -
      "index_select"  the caller's own line, generator.py:79: a dense fp32 [N,N,B,D] tensor built by torch (the
                      unchanged-caller contract; the only mode the reference modules understand)
      "gather"        ops.bank_gather: same dense fp32 tensor + its bf16 operand copy in one pass (1-line caller change)
      "banked"        ops.BankedRelation: kept factorised (SURVEY.md 8 f-0, 2-line caller change)"""

    def __init__(self, cfg: HotPathConfig, modules=None, relation_mode="banked"):
        super().__init__()
        c = self.cfg = cfg
        vocabs = {"relation": _V(c.rel_vocab_size), "predictable_token": _V(c.vocab_size)}
        self.vocabs = vocabs
        RelEnc = modules.RelationEncoder if modules is not None else RelationEncoder
        GraphTf = modules.GraphTransformer if modules is not None else GraphTransformer
        Tf = modules.Transformer if modules is not None else Transformer
        DecL = modules.DecodeLayer if modules is not None else DecodeLayer
        self.native = modules is None
        self.relation_encoder = RelEnc(vocabs["relation"], c.rel_dim, c.embed_dim, c.rnn_hidden_size,
                                       c.rnn_num_layers, c.dropout)
        self.graph_encoder = GraphTf(c.graph_layers, c.embed_dim, c.ff_embed_dim, c.num_heads, c.dropout)
        self.snt_encoder = Tf(c.snt_layers, c.embed_dim, c.ff_embed_dim, c.num_heads, c.dropout, with_external=True)
        self.decoder = DecL(vocabs, c.inference_layers, c.embed_dim, c.ff_embed_dim, c.num_heads,
                            c.concept_dim, c.rel_dim, c.dropout)
        self._prep_plan = ops.WeightPrepPlan()
        if not self.native:
            relation_mode = "index_select"
        assert relation_mode in ("index_select", "gather", "banked")
        self.relation_mode = relation_mode
        self.probe_generator = nn.Linear(c.embed_dim, c.embed_dim)
        nn.init.normal_(self.probe_generator.weight, std=0.02)
        nn.init.constant_(self.probe_generator.bias, 0.)

    def encode(self, batch):
        """generator.py:76-94: relation bank -> dense relation -> graph encoder -> probe / node states."""
        if self.native and batch.get("relation_row_counts") is not None:
            self.relation_encoder.row_counts = batch["relation_row_counts"]       # host-known: no device read-back
        bank = self.relation_encoder(batch["relation_bank"], batch["relation_length"])
        idx = batch["relation"]
        if self.relation_mode == "banked":
            # §8 f-0: keep relation = bank[idx] factorised (the 2-line caller change): no fp32
            # [N,N,B,D] tensor, bank-row GEMMs in the backward
            relation = ops.BankedRelation(bank, idx)
        elif self.relation_mode == "gather":
            relation = ops.bank_gather(bank, idx)                                   # generator.py:79 (+ bf16 copy)
        else:
            relation = bank.index_select(0, idx.view(-1)).view(*idx.size(), -1)     # generator.py:79, verbatim
        h = self.graph_encoder(batch["x"], relation, self_padding_mask=batch["node_mask"])
        probe = torch.tanh(self.probe_generator(h[:1]))
        return h[1:], batch["node_mask"][1:], probe

    def forward(self, batch):
        """generator.py:169-182 -> scalar loss."""
        if not self.native:
            return self._forward(batch)
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


def backward(loss, fresh_grads=True):
    """`loss.backward()` of the training loop (train.py:148).  fresh_grads=True (every .grad is None or will be
    overwritten, i.e. not a gradient accumulation step): the weight / bias / LayerNorm gradient kernels that run on the side
    stream are joined once at the end of the pass instead of after every sub-layer (ops.deferred_param_grads)."""
    if fresh_grads:
        with ops.deferred_param_grads():
            loss.backward()
    else:
        loss.backward()


BATCH_KEYS = ("relation_bank", "relation_length", "relation", "x", "node_mask", "token_repr", "token_mask",
              "copy_seq", "target", "causal_mask")


def batch_tensors(g):
    """synthetic.make_batch output -> the dict HotPath consumes (host tensors)."""
    T = g["T"]
    out = {k: g[k] for k in BATCH_KEYS if k in g}
    out["causal_mask"] = torch.ones(T, T, dtype=torch.bool).triu_(1)
    return out


def relation_row_counts(relation_length, Lmax):
    """[number of paths longer than t for t in range(Lmax)] from the HOST copy of relation_length (data.py:170-176 builds
    it there) - put it into the batch dictionary as "relation_row_counts" (a plain list, it does not travel to the GPU)"""
    ln = relation_length.cpu()
    return [int((ln > t).sum()) for t in range(int(Lmax))]
