"""CPU oracle for the gtos graph-transformer encode/decode hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gtos_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and there only as the checker / the timed CPU baseline.

This is a *restatement* (not a copy) of the reference algorithms as pure functions over a
flat parameter dict that uses the reference's ``state_dict`` key names, so weights move
between the reference modules, this oracle and the B200 modules with ``state_dict()`` /
``load_state_dict()`` and nothing else.  All math is torch CPU fp32; gradients come from
torch autograd over these functions.

Parity pin: ``tests/golden/make_golden.py`` imports the real reference modules from
``/root/reference/generator`` (seed 19940117, reference ``train.py:98``), records inputs,
weights, outputs and gradients, and ``tests/test_oracle_golden.py`` checks every function
here against those vectors.  The reference ships no golden vectors of its own.

Reference lines followed (all relative to /root/reference/generator):
  rel_mha              graph_transformer.py:93-174
  graph_layer          graph_transformer.py:47-66
  graph_transformer    graph_transformer.py:14-27
  mha                  transformer.py:98-173
  transformer_layer    transformer.py:44-72
  transformer          transformer.py:15-20
  relation_encoder     encoder.py:90-119   (nn.GRU semantics: gates r,z,n; packed sequences)
  token_generator      decoder.py:30-65
  decode_layer         decoder.py:76-94
"""
import math

import torch
import torch.nn.functional as F

NEG_INF = float("-inf")


# Matmul arithmetic of the oracle.  "fp32" is the reference's arithmetic.  "bf16" restates what the
# sm_100a tensor cores compute: both operands of every Linear rounded to bfloat16, products accumulated in
# fp32, and in backward the incoming gradient rounded to bfloat16 before the two gradient GEMMs.  It exists
# so gradient parity can be asserted tightly against a same-rounding reference (ReLU-kink flips make a bf16
# run differ from the fp32 one by several % in max-norm whatever the implementation).
_PRECISION = "fp32"


def set_matmul_precision(mode):
    global _PRECISION
    assert mode in ("fp32", "bf16")
    _PRECISION = mode


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


class _LinBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        xb, wb = _bf(x), _bf(w)
        ctx.save_for_backward(xb, wb)
        return xb @ wb.t()

    @staticmethod
    def backward(ctx, g):
        xb, wb = ctx.saved_tensors
        gb = _bf(g)
        dx = gb @ wb
        dw = gb.reshape(-1, gb.shape[-1]).t() @ xb.reshape(-1, xb.shape[-1])
        return dx, dw


class _RoundSTE(torch.autograd.Function):
    """bf16 rounding with a straight-through gradient (operands the kernels stage as bf16)"""

    @staticmethod
    def forward(ctx, x):
        return _bf(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _MatmulBf16(torch.autograd.Function):
    """a @ b with both operands (and, in backward, the incoming gradient) rounded to bf16, fp32 accumulation:
    what the attention-core kernels do for Q K^T, P V and their gradients (gtos_b200/csrc/attention.cu)."""

    @staticmethod
    def forward(ctx, a, b):
        ab, bb = _bf(a), _bf(b)
        ctx.save_for_backward(ab, bb)
        return ab @ bb

    @staticmethod
    def backward(ctx, g):
        ab, bb = ctx.saved_tensors
        gb = _bf(g)
        return gb @ bb.transpose(-1, -2), ab.transpose(-1, -2) @ gb


def _mm(a, b):
    return _MatmulBf16.apply(a, b) if _PRECISION == "bf16" else a @ b


def _lin32(x, w, b):
    """always-fp32 Linear: the 2-way copy/generate gate (decoder.py:42), which the B200 path keeps in fp32"""
    return x @ w.t() + b


def _lin(x, w, b=None):
    y = _LinBf16.apply(x, w) if _PRECISION == "bf16" else x @ w.t()
    return y if b is None else y + b


# ReLU sub-gradient protocol of the fp32-mode parity tests (tests/test_gpu_fp32_mode.py).  The gradient of relu is
# discontinuous at 0: two correct fp32 implementations whose pre-activations differ in the last bits disagree about a unit
# whose pre-activation is ~1e-6 of the layer's scale, and ONE such unit moves a row of the input gradient by percents.  A
# test may install a hook that decides the activation pattern (it takes it from the implementation under test for units
# within a narrow band around zero and checks that there is no disagreement outside the band).  None = plain relu.
_RELU_HOOK = None


def _relu(pre):
    return torch.relu(pre) if _RELU_HOOK is None else _RELU_HOOK(pre)


def _drop(x, p, training):
    return F.dropout(x, p=p, training=training) if (training and p > 0) else x


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


# ----------------------------------------------------------------------------------------
# Relation-aware multi-head attention (graph_transformer.py:93-174)
# ----------------------------------------------------------------------------------------
def rel_mha(P, pre, query, key, value, relation, num_heads, key_padding_mask=None,
            attn_mask=None, need_weights=False, dropout=0.0, training=False,
            weights_dropout=True):
    """query [T,B,D], key/value [S,B,D], relation [T,S,B,D] with relation[x][y] = path y->x.

    score[i,j] = hd^-1/2 * < q_i + Wa r[j,i] , k_j + Wb r[j,i] >   (graph_transformer.py:122-133);
    the reference's transpose at :123-124 means query i / key j reads relation[j][i].
    """
    T, B, D = query.shape
    S = key.shape[0]
    H = num_heads
    hd = D // H
    Win, bin_ = P[pre + "in_proj_weight"], P[pre + "in_proj_bias"]
    q = _lin(query, Win[:D], bin_[:D]).view(T, B, H, hd)
    k = _lin(key, Win[D:2 * D], bin_[D:2 * D]).view(S, B, H, hd)
    v = _lin(value, Win[2 * D:], bin_[2 * D:]).view(S, B, H, hd)
    if _PRECISION == "bf16":                                      # the fused kernel stages q, k as bf16
        q, k = _RoundSTE.apply(q), _RoundSTE.apply(k)
    Wr = P[pre + "relation_in_proj.weight"]                       # [2D, D], no bias (:80)
    ra = _lin(relation, Wr[:D]).view(relation.shape[0], relation.shape[1], B, H, hd)
    rb = _lin(relation, Wr[D:]).view(relation.shape[0], relation.shape[1], B, H, hd)
    # index as [j, i]: relation[j][i] feeds query i / key j
    qq = (q.unsqueeze(0) + ra) * (hd ** -0.5)                     # [j(S), i(T), B, H, hd]
    kk = k.unsqueeze(1) + rb                                      # [j, i, B, H, hd]
    score = (qq * kk).sum(-1)                                     # [j, i, B, H]
    score = score.permute(1, 0, 2, 3)                             # [i, j, B, H]
    if attn_mask is not None:                                     # [T,S] True = blocked (:136-140)
        score = score.masked_fill(attn_mask.bool()[:, :, None, None], NEG_INF)
    if key_padding_mask is not None:                              # [S,B] True = pad (:142-149)
        score = score.masked_fill(key_padding_mask.bool()[None, :, :, None], NEG_INF)
    w = torch.softmax(score, dim=1)                               # over keys j (:152)
    if weights_dropout:
        w = _drop(w, dropout, training)
    out = _mm(w.permute(2, 3, 0, 1), v.permute(1, 2, 0, 3)).permute(2, 0, 1, 3)   # [b,h,i,j]@[b,h,j,d] (:159)
    if not weights_dropout:
        out = _drop(out, dropout, training)
    out = _lin(out.reshape(T, B, D), P[pre + "out_proj.weight"], P[pre + "out_proj.bias"])
    return out, (w if need_weights else None)                     # weights [T,S,B,H] (:168-170)


def graph_layer(P, pre, x, relation, num_heads, kv=None, self_padding_mask=None,
                self_attn_mask=None, need_weights=False, dropout=0.0, training=False):
    """graph_transformer.py:47-66 (post-LN residual blocks)."""
    src = x if kv is None else kv
    a, w = rel_mha(P, pre + "self_attn.", x, src, src, relation, num_heads,
                   self_padding_mask, self_attn_mask, need_weights, dropout, training)
    x = _layer_norm(x + _drop(a, dropout, training),
                    P[pre + "attn_layer_norm.weight"], P[pre + "attn_layer_norm.bias"])
    h = _relu(_lin(x, P[pre + "fc1.weight"], P[pre + "fc1.bias"]))
    h = _drop(h, dropout, training)
    h = _drop(_lin(h, P[pre + "fc2.weight"], P[pre + "fc2.bias"]), dropout, training)
    x = _layer_norm(x + h, P[pre + "ff_layer_norm.weight"], P[pre + "ff_layer_norm.bias"])
    return x, w


def graph_transformer(P, pre, x, relation, num_layers, num_heads, kv=None,
                      self_padding_mask=None, self_attn_mask=None, dropout=0.0,
                      training=False, return_weights=False):
    """graph_transformer.py:14-27; every layer reads the SAME relation tensor."""
    ws = []
    for l in range(num_layers):
        x, w = graph_layer(P, f"{pre}layers.{l}.", x, relation, num_heads, kv,
                           self_padding_mask, self_attn_mask, return_weights, dropout, training)
        ws.append(w)
    return torch.stack(ws) if return_weights else x


# ----------------------------------------------------------------------------------------
# Vanilla multi-head attention / transformer layers (transformer.py)
# ----------------------------------------------------------------------------------------
def mha(P, pre, query, key, value, num_heads, key_padding_mask=None, attn_mask=None,
        need_weights=False, dropout=0.0, training=False, weights_dropout=True):
    """transformer.py:98-173.  Returned weights are max-over-heads, [T,B,S] (:164-169)."""
    T, B, D = query.shape
    S = key.shape[0]
    H = num_heads
    hd = D // H
    Win, bin_ = P[pre + "in_proj_weight"], P[pre + "in_proj_bias"]
    q = _lin(query, Win[:D], bin_[:D]).view(T, B, H, hd) * (hd ** -0.5)
    k = _lin(key, Win[D:2 * D], bin_[D:2 * D]).view(S, B, H, hd)
    v = _lin(value, Win[2 * D:], bin_[2 * D:]).view(S, B, H, hd)
    score = _mm(q.permute(1, 2, 0, 3), k.permute(1, 2, 3, 0))     # [b,h,t,d] @ [b,h,d,s]
    if attn_mask is not None:                                     # [T,S]
        score = score.masked_fill(attn_mask.bool()[None, None], NEG_INF)
    if key_padding_mask is not None:                              # [S,B]
        score = score.masked_fill(key_padding_mask.bool().t()[:, None, None, :], NEG_INF)
    w = torch.softmax(score, dim=-1)
    if weights_dropout:
        w = _drop(w, dropout, training)
    out = _mm(w, v.permute(1, 2, 0, 3)).permute(2, 0, 1, 3)         # [b,h,t,s] @ [b,h,s,d] -> [t,b,h,d]
    if not weights_dropout:
        out = _drop(out, dropout, training)
    out = _lin(out.reshape(T, B, D), P[pre + "out_proj.weight"], P[pre + "out_proj.bias"])
    wmax = w.max(dim=1)[0].transpose(0, 1) if need_weights else None   # [T,B,S]
    return out, wmax


def transformer_layer(P, pre, x, num_heads, kv=None, self_padding_mask=None,
                      self_attn_mask=None, external_memories=None, external_padding_mask=None,
                      with_external=False, need_weights=False, dropout=0.0, training=False):
    """transformer.py:44-72."""
    src = x if kv is None else kv
    a, sw = mha(P, pre + "self_attn.", x, src, src, num_heads, self_padding_mask,
                self_attn_mask, need_weights, dropout, training)
    x = _layer_norm(x + _drop(a, dropout, training),
                    P[pre + "attn_layer_norm.weight"], P[pre + "attn_layer_norm.bias"])
    ew = None
    if with_external:
        a, ew = mha(P, pre + "external_attn.", x, external_memories, external_memories,
                    num_heads, external_padding_mask, None, need_weights, dropout, training)
        x = _layer_norm(x + _drop(a, dropout, training),
                        P[pre + "external_layer_norm.weight"], P[pre + "external_layer_norm.bias"])
    h = _relu(_lin(x, P[pre + "fc1.weight"], P[pre + "fc1.bias"]))
    h = _drop(h, dropout, training)
    h = _drop(_lin(h, P[pre + "fc2.weight"], P[pre + "fc2.bias"]), dropout, training)
    x = _layer_norm(x + h, P[pre + "ff_layer_norm.weight"], P[pre + "ff_layer_norm.bias"])
    return x, sw, ew


def transformer(P, pre, x, num_layers, num_heads, kv=None, self_padding_mask=None,
                self_attn_mask=None, external_memories=None, external_padding_mask=None,
                with_external=False, dropout=0.0, training=False):
    """transformer.py:15-20."""
    for l in range(num_layers):
        x, _, _ = transformer_layer(P, f"{pre}layers.{l}.", x, num_heads, kv, self_padding_mask,
                                    self_attn_mask, external_memories, external_padding_mask,
                                    with_external, False, dropout, training)
    return x


# ----------------------------------------------------------------------------------------
# RelationEncoder: embedding -> 2-layer bidirectional GRU over packed sequences -> Linear
# (encoder.py:90-119).  nn.GRU cell: r = s(Wir x + bir + Whr h + bhr), z likewise,
# n = tanh(Win x + bin + r*(Whn h + bhn)), h' = (1-z)*n + z*h.  Packed semantics: a sequence
# of length L only sees steps 0..L-1; the reverse direction starts at step L-1.
# ----------------------------------------------------------------------------------------
def _gru_dir(x, lengths, w_ih, w_hh, b_ih, b_hh, reverse):
    Lmax, R, _ = x.shape
    Hh = w_hh.shape[1]
    h = x.new_zeros(R, Hh)
    outs = [None] * Lmax
    order = range(Lmax - 1, -1, -1) if reverse else range(Lmax)
    for t in order:
        gi = _lin(x[t], w_ih, b_ih)
        gh = _lin(h, w_hh, b_hh)
        i_r, i_z, i_n = gi.chunk(3, -1)
        h_r, h_z, h_n = gh.chunk(3, -1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h_new = (1 - z) * n + z * h
        live = (lengths > t).unsqueeze(-1)
        h = torch.where(live, h_new, h)
        outs[t] = torch.where(live, h, torch.zeros_like(h))
    return torch.stack(outs), h


def relation_encoder(P, pre, src_tokens, src_lengths, num_layers=2, dropout=0.0, training=False):
    """src_tokens [Lmax,R] int64 (0 = pad), src_lengths [R] -> [R, embed_dim]."""
    x = P[pre + "rel_embed.weight"][src_tokens]                  # encoder.py:96
    x = _drop(x, dropout, training)
    finals = None
    for l in range(num_layers):
        f_out, f_h = _gru_dir(x, src_lengths, P[f"{pre}rnn.weight_ih_l{l}"], P[f"{pre}rnn.weight_hh_l{l}"],
                              P[f"{pre}rnn.bias_ih_l{l}"], P[f"{pre}rnn.bias_hh_l{l}"], False)
        b_out, b_h = _gru_dir(x, src_lengths, P[f"{pre}rnn.weight_ih_l{l}_reverse"],
                              P[f"{pre}rnn.weight_hh_l{l}_reverse"], P[f"{pre}rnn.bias_ih_l{l}_reverse"],
                              P[f"{pre}rnn.bias_hh_l{l}_reverse"], True)
        x = torch.cat([f_out, b_out], -1)
        if l < num_layers - 1:
            x = _drop(x, dropout, training)                       # nn.GRU inter-layer dropout
        finals = torch.cat([f_h, b_h], -1)                        # encoder.py:108-111
    return _lin(finals, P[pre + "out_proj.weight"], P[pre + "out_proj.bias"])   # :117


# ----------------------------------------------------------------------------------------
# Decoder (decoder.py)
# ----------------------------------------------------------------------------------------
def token_generator(P, pre, outs, graph_state, graph_padding_mask, copy_seq, pad_idx,
                    target=None, work=False, dropout=0.0, training=False):
    """decoder.py:30-65: copy / generate mixture over the (batch-extended) vocabulary."""
    a, align = mha(P, pre + "alignment_layer.", outs, graph_state, graph_state, 1,
                   graph_padding_mask, None, True, dropout, training, weights_dropout=False)
    outs = _layer_norm(outs + _drop(a, dropout, training),
                       P[pre + "alignment_layer_norm.weight"], P[pre + "alignment_layer_norm.bias"])
    T, B, _ = outs.shape
    tok = torch.tanh(_lin(outs, P[pre + "transfer.weight"], P[pre + "transfer.bias"]))
    tok = _drop(tok, dropout, training)
    gate = torch.softmax(_lin32(tok, P[pre + "diverter.weight"], P[pre + "diverter.bias"]), -1)
    gen_gate, copy_gate = gate[..., :1], gate[..., 1:]
    probs = gen_gate * torch.softmax(_lin(tok, P[pre + "generator.weight"], P[pre + "generator.bias"]), -1)
    V = probs.shape[-1]
    tot = 1 + int(copy_seq.max())
    if tot > V:
        probs = torch.cat([probs, probs.new_zeros(T, B, tot - V)], -1)
    index = copy_seq.t().reshape(1, B, -1).expand(T, -1, -1)      # copy_seq [S,B]
    probs = probs.scatter_add(-1, index, copy_gate * align)
    ll = torch.log(probs + 1e-12)
    if work:
        return ll
    loss = -ll.gather(-1, target.unsqueeze(-1)).squeeze(-1)
    loss = loss.masked_fill(target.eq(pad_idx), 0.).sum(0)
    return loss


def decode_layer(P, pre, probe, graph_state, snt_state, graph_padding_mask, snt_padding_mask,
                 attn_mask, copy_seq, num_layers, num_heads, pad_idx, target=None, work=False,
                 dropout=0.0, training=False):
    """decoder.py:76-94."""
    outs = _drop(probe, dropout, training)
    outs = transformer(P, pre + "inference_core.", outs, num_layers, num_heads, kv=snt_state,
                       self_padding_mask=snt_padding_mask, self_attn_mask=attn_mask,
                       external_memories=graph_state, external_padding_mask=graph_padding_mask,
                       with_external=True, dropout=dropout, training=training)
    if work:
        return token_generator(P, pre + "token_generator.", outs, graph_state, graph_padding_mask,
                               copy_seq, pad_idx, work=True, dropout=dropout, training=training)
    loss = token_generator(P, pre + "token_generator.", outs, graph_state, graph_padding_mask,
                           copy_seq, pad_idx, target=target, dropout=dropout, training=training)
    tot = snt_padding_mask.shape[0] - snt_padding_mask.float().sum(0)
    return (loss / tot).mean()


def causal_mask(T):
    """transformer.py:204-219 (bool instead of the torch-1.1 uint8)."""
    return torch.ones(T, T, dtype=torch.bool).triu_(1)


def bank_to_dense(bank, idx):
    """generator.py:79 (training path): relation[x][y][b] = bank[idx[x][y][b]]."""
    return bank.index_select(0, idx.reshape(-1)).view(*idx.shape, -1)
