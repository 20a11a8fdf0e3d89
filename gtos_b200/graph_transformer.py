"""B200-native drop-in for the reference's ``graph_transformer.py`` (generator/ and translator/).

Same classes, constructor arguments, ``forward`` signatures, attributes and parameter names
(``state_dict`` compatible with reference checkpoints, generator/work.py:109); the math runs in
libgtos_b200.so: one fused tcgen05 kernel per layer projects every (i, j) relation vector and
folds it into the attention scores (reference: generator/graph_transformer.py:93-174).
"""
import torch
from torch import nn
from torch.nn import Parameter

from . import ops


class GraphTransformer(nn.Module):
    """reference: graph_transformer.py:6-27"""

    def __init__(self, layers, embed_dim, ff_embed_dim, num_heads, dropout, weights_dropout=True):
        super().__init__()
        self.layers = nn.ModuleList()
        for _ in range(layers):
            self.layers.append(GraphTransformerLayer(embed_dim, ff_embed_dim, num_heads, dropout, weights_dropout))

    def forward(self, x, relation, kv=None, self_padding_mask=None, self_attn_mask=None):
        # every layer reads the SAME relation tensor (:16-17): stage its bf16 copy once (or reuse the one
        # ops.bank_gather made while building the tensor)
        acc, token = None, None
        if kv is not None and kv is not x:
            # never used by gtos (generator.py:90); computed on the composed path layer by layer
            if isinstance(relation, ops.BankedRelation):
                relation = relation.dense()
            for layer in self.layers:
                x, _ = layer(x, relation, kv, self_padding_mask, self_attn_mask)
            return x
        if ops.fp32_mode():
            return self._forward_fp32(x, relation, self_padding_mask, self_attn_mask, False)
        if not isinstance(relation, ops.BankedRelation):
            src = ops.factorised_source(relation)       # opt-in: a dense tensor that still knows it is bank[idx]
            if src is not None:
                relation = src
        if isinstance(relation, ops.BankedRelation) and relation.multi and torch.is_grad_enabled() and relation.requires_grad:
            relation = relation.dense()            # gradients through the evaluation multi-path mean: dense autograd path
        if isinstance(relation, ops.BankedRelation):
            # bank-factorised relation: dense bf16 operand for the fused kernels, bank-row GEMMs
            # in the backward
            banked, relation, relb = relation, None, None
            if torch.is_grad_enabled() and banked.requires_grad:
                banked.prepare(self.layers[0].self_attn.num_heads)
                acc = ops.RelGradAcc(banked, len(self.layers))
                token = ops.BankTokenFn.apply(banked.bank, acc)
        else:
            banked = None
            relb = _staged_bf16(relation)
            # one shared gradient buffer for the relation tensor instead of L per-layer tensors summed by autograd
            if torch.is_grad_enabled() and relation.requires_grad:
                acc = ops.RelGradAcc()
                token = ops.RelTokenFn.apply(relation, acc)
        xb = None
        for layer in self.layers:
            x, xb, _ = layer._forward(x, xb, relation, relb, kv, self_padding_mask, self_attn_mask, False,
                                      rel_token=token, rel_acc=acc, banked=banked)
        return x

    def _forward_fp32(self, x, relation, self_padding_mask, self_attn_mask, need_weights):
        """fp32 mode (ops.set_precision("fp32"), 1e-3 against the fp32 reference): the dense relation tensor, staged ONCE as
        the split-bf16 operand all layers' K-tripled projection GEMMs read; a factorised relation is gathered first."""
        from . import ops32
        if isinstance(relation, ops.BankedRelation):
            relation = relation.dense() if relation.multi else ops.bank_gather(relation.bank, relation.idx)
        N1, N2, B, D = relation.shape
        rel3 = ops32.split3(relation.detach().contiguous().view(N1 * N2 * B, D), 0)
        acc = token = None
        if torch.is_grad_enabled() and relation.requires_grad:
            acc = ops.RelGradAcc()
            token = ops.RelTokenFn.apply(relation, acc)
        attns = []
        for layer in self.layers:
            x, _, w = layer._forward(x, None, relation, rel3, None, self_padding_mask, self_attn_mask, need_weights,
                                     rel_token=token, rel_acc=acc)
            attns.append(w)
        return torch.stack(attns) if need_weights else x

    def get_attn_weights(self, x, relation, kv=None, self_padding_mask=None, self_attn_mask=None):
        if kv is not None and kv is not x:
            if isinstance(relation, ops.BankedRelation):
                relation = relation.dense()
            attns = []
            for layer in self.layers:
                x, attn = layer(x, relation, kv, self_padding_mask, self_attn_mask, need_weights=True)
                attns.append(attn)
            return torch.stack(attns)
        if ops.fp32_mode():
            return self._forward_fp32(x, relation, self_padding_mask, self_attn_mask, True)
        banked = None
        if isinstance(relation, ops.BankedRelation):
            banked, relation, relb = relation, None, None
        else:
            relb = _staged_bf16(relation)
        attns, xb = [], None
        for layer in self.layers:
            x, xb, attn = layer._forward(x, xb, relation, relb, kv, self_padding_mask, self_attn_mask, True, banked=banked)
            attns.append(attn)
        return torch.stack(attns)


def _staged_bf16(relation):
    relb = ops.staged_relation_bf16(relation)
    if relb is not None:
        return relb
    return ops.relation_to_bf16(relation.detach().contiguous())


class GraphTransformerLayer(nn.Module):
    """reference: graph_transformer.py:29-66 (post-LN residual blocks)"""

    def __init__(self, embed_dim, ff_embed_dim, num_heads, dropout, weights_dropout=True):
        super().__init__()
        self.self_attn = RelationMultiheadAttention(embed_dim, num_heads, dropout, weights_dropout)
        self.fc1 = nn.Linear(embed_dim, ff_embed_dim)
        self.fc2 = nn.Linear(ff_embed_dim, embed_dim)
        self.attn_layer_norm = nn.LayerNorm(embed_dim)
        self.ff_layer_norm = nn.LayerNorm(embed_dim)
        self.dropout = dropout
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.fc1.weight, std=0.02)
        nn.init.normal_(self.fc2.weight, std=0.02)
        nn.init.constant_(self.fc1.bias, 0.)
        nn.init.constant_(self.fc2.bias, 0.)

    def _forward(self, x, xb, relation, relb, kv, self_padding_mask, self_attn_mask, need_weights, rel_token=None,
                 rel_acc=None, banked=None):
        p = self.dropout if self.training else 0.0
        if kv is not None and kv is not x:
            # graph_transformer.py:52-55 `self_attn(query=x, key=kv, value=kv, ...)`: never used by gtos
            # (generator.py:90); the general entry point computes it on the composed path
            a, w = self.self_attn(x, kv, kv, relation, key_padding_mask=self_padding_mask, attn_mask=self_attn_mask,
                                  need_weights=need_weights)
        else:
            a, w = self.self_attn._forward(x, xb, relation, relb, self_padding_mask, self_attn_mask, need_weights,
                                           rel_token, rel_acc, banked)
        x, xb = ops.add_layer_norm(a, x, self.attn_layer_norm.weight, self.attn_layer_norm.bias, p)
        h = ops.ffn(x, xb, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, p)
        x, xb = ops.add_layer_norm(h, x, self.ff_layer_norm.weight, self.ff_layer_norm.bias, p)
        return x, xb, w

    def forward(self, x, relation, kv=None, self_padding_mask=None, self_attn_mask=None, need_weights=False):
        if ops.fp32_mode() and isinstance(relation, ops.BankedRelation):
            relation = relation.dense() if relation.multi else ops.bank_gather(relation.bank, relation.idx)
        if isinstance(relation, ops.BankedRelation) and (kv is None or kv is x):
            x, _, w = self._forward(x, None, None, None, kv, self_padding_mask, self_attn_mask, need_weights, banked=relation)
        else:
            x, _, w = self._forward(x, None, relation, None, kv, self_padding_mask, self_attn_mask, need_weights)
        return x, w


class RelationMultiheadAttention(nn.Module):
    """reference: graph_transformer.py:68-197"""

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
        self.relation_in_proj = nn.Linear(embed_dim, 2 * embed_dim, bias=False)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.weights_dropout = weights_dropout
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.in_proj_weight, std=0.02)
        nn.init.normal_(self.out_proj.weight, std=0.02)
        nn.init.normal_(self.relation_in_proj.weight, std=0.02)
        nn.init.constant_(self.in_proj_bias, 0.)
        nn.init.constant_(self.out_proj.bias, 0.)

    def _forward(self, x, xb, relation, relb, key_padding_mask, attn_mask, need_weights, rel_token=None, rel_acc=None,
                 banked=None):
        p = self.dropout if self.training else 0.0
        if ops.fp32_mode():
            # fp32 mode: `relb` carries the staged split operand of the relation (or None), never a bf16 copy
            from . import ops32
            if banked is not None:
                relation = banked.dense() if banked.multi else ops.bank_gather(banked.bank, banked.idx)
                relb = None
            if relb is not None and relb.dim() != 2:
                relb = None
            out, w = ops32.RelAttn32Fn.apply(x, relation, relb, ops.as_u8(key_padding_mask), ops.as_u8(attn_mask),
                                             self.in_proj_weight, self.in_proj_bias, self.relation_in_proj.weight,
                                             self.out_proj.weight, self.out_proj.bias, self.num_heads, float(p),
                                             bool(need_weights), rel_token, rel_acc, bool(self.weights_dropout))
            if w is not None:
                w = w.permute(2, 3, 0, 1)
            return out, w
        out, w = ops.RelAttnFn.apply(x, xb, relation, relb, ops.as_u8(key_padding_mask), ops.as_u8(attn_mask),
                                     self.in_proj_weight, self.in_proj_bias, self.relation_in_proj.weight,
                                     self.out_proj.weight, self.out_proj.bias, self.num_heads, float(p),
                                     bool(need_weights), rel_token, rel_acc, bool(self.weights_dropout), banked)
        if w is not None:
            w = w.permute(2, 3, 0, 1)           # [B,H,T,S] -> [tgt, src, bsz, heads]   (:168-170)
        return out, w

    def forward(self, query, key, value, relation, key_padding_mask=None, attn_mask=None, need_weights=False):
        """Input shape: Time x Batch x Channel; relation: tgt_len x src_len x bsz x dim (:94-98)."""
        if not (key is query and value is query) and not (
                key.data_ptr() == query.data_ptr() == value.data_ptr() and key.shape == query.shape):
            # the `kv_same` / general branches of graph_transformer.py:108-116: gtos always passes query = key = value
            # (generator.py:90), so these run on the composed path (projections on the tcgen05 GEMM) instead of the
            # fused kernel
            if isinstance(relation, ops.BankedRelation):
                relation = relation.dense()
            p = self.dropout if self.training else 0.0
            return ops.rel_attention_composed(query, key, value, relation, key_padding_mask, attn_mask,
                                              self.in_proj_weight, self.in_proj_bias, self.relation_in_proj.weight,
                                              self.out_proj.weight, self.out_proj.bias, self.num_heads, float(p),
                                              bool(self.weights_dropout), bool(need_weights))
        if isinstance(relation, ops.BankedRelation):
            return self._forward(query, None, None, None, key_padding_mask, attn_mask, need_weights, banked=relation)
        return self._forward(query, None, relation, None, key_padding_mask, attn_mask, need_weights)

    # projection helpers kept for API parity (:176-197); they run the tcgen05 GEMM
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
