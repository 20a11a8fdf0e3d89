"""B200-native drop-in for the reference's ``transformer.py``: Transformer, TransformerLayer,
MultiheadAttention (same surface / parameter names), plus the small host-side helpers the callers
import from this module (Embedding, SelfAttentionMask, positional embeddings) kept as PyTorch glue
(SURVEY.md §2.1).  Reference: generator/transformer.py.
"""
import math

import torch
from torch import nn
from torch.nn import Parameter

from . import ops


class Transformer(nn.Module):
    """reference: transformer.py:7-20"""

    def __init__(self, layers, embed_dim, ff_embed_dim, num_heads, dropout, with_external=False, weights_dropout=True):
        super().__init__()
        self.layers = nn.ModuleList()
        for _ in range(layers):
            self.layers.append(TransformerLayer(embed_dim, ff_embed_dim, num_heads, dropout, with_external,
                                                weights_dropout))

    def forward(self, x, kv=None, self_padding_mask=None, self_attn_mask=None, external_memories=None,
                external_padding_mask=None):
        xb = kvb = memb = None
        for layer in self.layers:
            x, xb, kvb, memb, _, _ = layer._forward(x, xb, kv, kvb, self_padding_mask, self_attn_mask,
                                                    external_memories, memb, external_padding_mask, False)
        return x


class TransformerLayer(nn.Module):
    """reference: transformer.py:22-72"""

    def __init__(self, embed_dim, ff_embed_dim, num_heads, dropout, with_external=False, weights_dropout=True):
        super().__init__()
        self.self_attn = MultiheadAttention(embed_dim, num_heads, dropout, weights_dropout)
        self.fc1 = nn.Linear(embed_dim, ff_embed_dim)
        self.fc2 = nn.Linear(ff_embed_dim, embed_dim)
        self.attn_layer_norm = nn.LayerNorm(embed_dim)
        self.ff_layer_norm = nn.LayerNorm(embed_dim)
        self.with_external = with_external
        self.dropout = dropout
        if self.with_external:
            self.external_attn = MultiheadAttention(embed_dim, num_heads, dropout, weights_dropout)
            self.external_layer_norm = nn.LayerNorm(embed_dim)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.fc1.weight, std=0.02)
        nn.init.normal_(self.fc2.weight, std=0.02)
        nn.init.constant_(self.fc1.bias, 0.)
        nn.init.constant_(self.fc2.bias, 0.)

    def _forward(self, x, xb, kv, kvb, self_padding_mask, self_attn_mask, external_memories, memb,
                 external_padding_mask, need_weights):
        p = self.dropout if self.training else 0.0
        if kv is None:
            a, sw = self.self_attn._forward(x, xb, x, xb, True, self_padding_mask, self_attn_mask, need_weights)
        else:
            if kvb is None and not ops.fp32_mode():
                kvb = ops.cast_bf16(kv.contiguous().view(-1, kv.shape[-1])).view(kv.shape)
            a, sw = self.self_attn._forward(x, xb, kv, kvb, False, self_padding_mask, self_attn_mask, need_weights)
        x, xb = ops.add_layer_norm(a, x, self.attn_layer_norm.weight, self.attn_layer_norm.bias, p)
        ew = None
        if self.with_external:
            if memb is None and not ops.fp32_mode():
                memb = ops.cast_bf16(external_memories.contiguous().view(-1, external_memories.shape[-1])).view(
                    external_memories.shape)
            a, ew = self.external_attn._forward(x, xb, external_memories, memb, False, external_padding_mask, None,
                                                need_weights)
            x, xb = ops.add_layer_norm(a, x, self.external_layer_norm.weight, self.external_layer_norm.bias, p)
        h = ops.ffn(x, xb, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, p)
        x, xb = ops.add_layer_norm(h, x, self.ff_layer_norm.weight, self.ff_layer_norm.bias, p)
        return x, xb, kvb, memb, sw, ew

    def forward(self, x, kv=None, self_padding_mask=None, self_attn_mask=None, external_memories=None,
                external_padding_mask=None, need_weights=False):
        x, _, _, _, sw, ew = self._forward(x, None, kv, None, self_padding_mask, self_attn_mask, external_memories,
                                           None, external_padding_mask, need_weights)
        return x, sw, ew


class MultiheadAttention(nn.Module):
    """reference: transformer.py:74-196"""

    def __init__(self, embed_dim, num_heads, dropout=0., weights_dropout=True):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        self.scaling = self.head_dim ** -0.5
        self.in_proj_weight = Parameter(torch.Tensor(3 * embed_dim, embed_dim))
        self.in_proj_bias = Parameter(torch.Tensor(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.weights_dropout = weights_dropout
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.in_proj_weight, std=0.02)
        nn.init.normal_(self.out_proj.weight, std=0.02)
        nn.init.constant_(self.in_proj_bias, 0.)
        nn.init.constant_(self.out_proj.bias, 0.)

    def _forward(self, query, qb, key, kb, self_attn, key_padding_mask, attn_mask, need_weights):
        p = self.dropout if self.training else 0.0
        if ops.fp32_mode():
            from . import ops32
            out, w = ops32.MHA32Fn.apply(query, key, bool(self_attn), ops.as_u8(key_padding_mask), ops.as_u8(attn_mask),
                                         self.in_proj_weight, self.in_proj_bias, self.out_proj.weight, self.out_proj.bias,
                                         self.num_heads, float(p), bool(self.weights_dropout), bool(need_weights))
            if w is not None:
                w = w.max(dim=1)[0].transpose(0, 1)
            return out, w
        out, w = ops.MHAFn.apply(query, qb, key, kb, bool(self_attn), ops.as_u8(key_padding_mask),
                                 ops.as_u8(attn_mask), self.in_proj_weight, self.in_proj_bias, self.out_proj.weight,
                                 self.out_proj.bias, self.num_heads, float(p), bool(self.weights_dropout),
                                 bool(need_weights))
        if w is not None:
            # maximum attention weight over heads, [tgt, bsz, src]  (:164-169)
            w = w.max(dim=1)[0].transpose(0, 1)
        return out, w

    def forward(self, query, key, value, key_padding_mask=None, attn_mask=None, need_weights=False):
        """Input shape: Time x Batch x Channel; key_padding_mask: Time x batch; attn_mask: tgt_len x src_len"""
        qkv_same = query.data_ptr() == key.data_ptr() == value.data_ptr() and query.shape == key.shape
        kv_same = key.data_ptr() == value.data_ptr() and key.shape == value.shape
        if not kv_same:
            # transformer.py:113-118 (separate k and v inputs): never used by gtos - composed path
            p = self.dropout if self.training else 0.0
            return ops.mha_composed(query, key, value, key_padding_mask, attn_mask, self.in_proj_weight, self.in_proj_bias,
                                    self.out_proj.weight, self.out_proj.bias, self.num_heads, float(p),
                                    bool(self.weights_dropout), bool(need_weights))
        return self._forward(query, None, key, None, qkv_same, key_padding_mask, attn_mask, need_weights)

    def _in_proj(self, input, start=0, end=None):
        end = 3 * self.embed_dim if end is None else end
        W, b = self.in_proj_weight[start:end].contiguous(), self.in_proj_bias[start:end].contiguous()
        shp = input.shape
        if ops.fp32_mode():
            from . import ops32
            return ops32.mm3(input.reshape(-1, shp[-1]), W, b).view(*shp[:-1], end - start)
        Wb, _ = ops.weight_prep(W, want_t=False)
        y, _ = ops.gemm_tn(ops.cast_bf16(input.reshape(-1, shp[-1])), Wb, end - start, bias=b)
        return y.view(*shp[:-1], end - start)

    def in_proj_qkv(self, query):
        return self._in_proj(query).chunk(3, dim=-1)

    def in_proj_kv(self, key):
        return self._in_proj(key, start=self.embed_dim).chunk(2, dim=-1)

    def in_proj_q(self, query):
        return self._in_proj(query, end=self.embed_dim)

    def in_proj_k(self, key):
        return self._in_proj(key, start=self.embed_dim, end=2 * self.embed_dim)

    def in_proj_v(self, value):
        return self._in_proj(value, start=2 * self.embed_dim)


# ---- host-side glue the reference callers import from this module (plain PyTorch, not on the hot path) ----
from .host_glue import Embedding, SelfAttentionMask, SinusoidalPositionalEmbedding  # noqa: E402,F401
