"""B200-native drop-in for the reference's ``decoder.py``: DecodeLayer (+ TokenGenerator).

DecodeLayer's three attention layers (masked self-attn over the token states + cross-attn over the
graph, generator/decoder.py:76-94) and the TokenGenerator's 1-head alignment attention run through the
sm_100a kernels.  The vocabulary tail of TokenGenerator (two softmaxes, copy scatter, log / NLL; decoder.py:42-64,
SURVEY.md §8 row f-2) is one fused kernel for the training loss and one for the work=True log-prob table.
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .transformer import MultiheadAttention, Transformer


class TokenGenerator(nn.Module):
    """reference: decoder.py:10-65"""

    def __init__(self, vocabs, embed_dim, token_size, dropout):
        super().__init__()
        self.alignment_layer = MultiheadAttention(embed_dim, 1, dropout, weights_dropout=False)
        self.alignment_layer_norm = nn.LayerNorm(embed_dim)
        self.transfer = nn.Linear(embed_dim, token_size)
        self.generator = nn.Linear(token_size, vocabs['predictable_token'].size)
        self.diverter = nn.Linear(token_size, 2)
        self.vocabs = vocabs
        self.dropout = dropout
        # size of the batch-extended vocabulary; the reference reads it back from the device every call
        # (copy_seq.max().item(), decoder.py:46).  A caller that knows it on the host (the data loader
        # does) can set it to avoid the sync and make the step CUDA-graph capturable.
        self.static_tot_ext = None
        self.reset_parameters()

    def reset_parameters(self):
        # N(0, 0.02) weights, zero biases (decoder.py:21-27)
        for lin in (self.transfer, self.diverter, self.generator):
            nn.init.normal_(lin.weight, std=0.02)
            nn.init.zeros_(lin.bias)

    def forward(self, outs, graph_state, graph_padding_mask, copy_seq, target=None, work=False):
        p = self.dropout if self.training else 0.0
        x, alignment_weight = self.alignment_layer(outs, graph_state, graph_state,
                                                   key_padding_mask=graph_padding_mask, need_weights=True)
        outs, _ = ops.add_layer_norm(x, outs, self.alignment_layer_norm.weight, self.alignment_layer_norm.bias, p)
        seq_len, bsz, _ = outs.size()
        # ---- vocabulary tail: both wide projections on the tcgen05 GEMM; the training NLL and the
        #      work=True (beam search) log-prob table are one fused kernel each ----
        outs_token = torch.tanh(ops.linear(outs, self.transfer.weight, self.transfer.bias))
        outs_token = F.dropout(outs_token, p=self.dropout, training=self.training)
        gate_logits = self.diverter(outs_token)
        logits = ops.linear(outs_token, self.generator.weight, self.generator.bias)
        if not work:
            # training: NLL of the target under the copy/generate mixture, fused over the vocabulary
            # (decoder.py:42-64; copy ids >= vocab size only ever receive copy mass, as in the reference)
            token_loss = ops.token_nll(logits, gate_logits, alignment_weight, copy_seq, target,
                                       self.vocabs['predictable_token'].padding_idx)
            return token_loss.sum(0)
        # work=True (beam search): full log-probability table over the batch-extended vocabulary (decoder.py:44-59)
        return self._log_prob_table(logits, gate_logits, alignment_weight, copy_seq)

    def _log_prob_table(self, logits, gate_logits, align, copy_seq):
        """decoder.py:44-59 as one fused kernel (inference only: the table carries no autograd graph, like the
        reference's use of work=True under torch.no_grad(), generator.py:97-98)."""
        T, B, V = logits.shape
        S = copy_seq.size(0)
        width = self.static_tot_ext if self.static_tot_ext is not None else int(copy_seq.max()) + 1
        table = ops.token_logprob(logits.detach().reshape(T * B, V), gate_logits.detach().reshape(T * B, 2),
                                  align.detach().reshape(T * B, S), copy_seq, None, max(width, V), B=B)
        return table.view(T, B, -1)


class DecodeLayer(nn.Module):
    """reference: decoder.py:67-94"""

    def __init__(self, vocabs, inference_layers, embed_dim, ff_embed_dim, num_heads, token_size, rel_size, dropout):
        super().__init__()
        self.inference_core = Transformer(inference_layers, embed_dim, ff_embed_dim, num_heads, dropout,
                                          with_external=True)
        self.token_generator = TokenGenerator(vocabs, embed_dim, token_size, dropout)
        self.dropout = dropout
        self.vocabs = vocabs

    def forward(self, probe, graph_state, snt_state, graph_padding_mask, snt_padding_mask, attn_mask, copy_seq,
                target=None, work=False):
        # probe: tgt_len x bsz x embed_dim ; snt_state, graph_state: seq_len x bsz x embed_dim
        outs = probe
        if self.training and self.dropout > 0:                                     # decoder.py:82
            outs = _DropoutFn.apply(probe, float(self.dropout))
        outs = self.inference_core(outs, kv=snt_state, self_padding_mask=snt_padding_mask, self_attn_mask=attn_mask,
                                   external_memories=graph_state, external_padding_mask=graph_padding_mask)
        if work:
            return self.token_generator(outs, graph_state, graph_padding_mask, copy_seq, work=True)
        token_loss = self.token_generator(outs, graph_state, graph_padding_mask, copy_seq, target=target, work=False)
        token_tot = snt_padding_mask.size(0) - snt_padding_mask.float().sum(0)
        token_loss = token_loss / token_tot
        return token_loss.mean()


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p):
        seed, off = ops.rng_state(x.device), ops.new_seed_off()
        ctx.meta = (p, seed, off)
        return ops.dropout_f32(x.contiguous(), p, seed, off)

    @staticmethod
    def backward(ctx, dy):
        p, seed, off = ctx.meta
        return ops.dropout_f32(dy.contiguous(), p, seed, off), None
