"""CPU oracle of the assembled hot path (gtos_b200/hotpath.py == generator/generator.py:71-94,169-182).
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/gtos_oracle.py header)."""
import torch

from . import gtos_oracle as O


def hotpath_loss(P, batch, cfg, dropout=0.0, training=False):
    """P: HotPath.state_dict()-keyed parameter dict (CPU fp32); batch: hotpath.batch_tensors() output."""
    bank = O.relation_encoder(P, "relation_encoder.", batch["relation_bank"], batch["relation_length"],
                              num_layers=cfg.rnn_num_layers, dropout=dropout, training=training)
    relation = O.bank_to_dense(bank, batch["relation"])
    h = O.graph_transformer(P, "graph_encoder.", batch["x"], relation, cfg.graph_layers, cfg.num_heads,
                            self_padding_mask=batch["node_mask"], dropout=dropout, training=training)
    probe = torch.tanh(h[:1] @ P["probe_generator.weight"].t() + P["probe_generator.bias"])
    concept_repr, concept_mask = h[1:], batch["node_mask"][1:]
    tok = O.transformer(P, "snt_encoder.", batch["token_repr"], cfg.snt_layers, cfg.num_heads,
                        self_padding_mask=batch["token_mask"], self_attn_mask=batch["causal_mask"],
                        external_memories=concept_repr, external_padding_mask=concept_mask, with_external=True,
                        dropout=dropout, training=training)
    probe = probe.expand_as(tok)
    return O.decode_layer(P, "decoder.", probe, concept_repr, tok, concept_mask, batch["token_mask"],
                          batch["causal_mask"], batch["copy_seq"], cfg.inference_layers, cfg.num_heads, 0,
                          target=batch["target"], dropout=dropout, training=training)


def encoder_only(P, x, relation, mask, cfg, dropout=0.0, training=False):
    return O.graph_transformer(P, "graph_encoder.", x, relation, cfg.graph_layers, cfg.num_heads,
                               self_padding_mask=mask, dropout=dropout, training=training)
