"""B200-native drop-in for ``RelationEncoder`` of the reference's ``encoder.py`` (generator/encoder.py:66-119):
embeds every distinct shortest-path label sequence of the batch, runs a 2-layer bidirectional GRU over the
packed sequences and projects the final states to the relation bank [R, embed_dim].

The TokenEncoder / CNN / Highway embedding front-end of the same reference file is out of the hot path
 and is not re-implemented here.
"""
import torch
from torch import nn

import os

from . import ops
from .transformer import Embedding

# GTOS_GRU_LEN_SORT=0: every path at every time step (masked in the kernels) instead of the length-sorted prefix schedule
_LEN_SORT = os.environ.get("GTOS_GRU_LEN_SORT", "1") == "1"


def AMREmbedding(vocab, embedding_dim, pretrained_file=None, amr=False, dump_file=None):
    """reference: encoder.py:9-64 (host-side construction; the pretrained branch reads a text file of vectors)."""
    if pretrained_file is not None:
        from .host_glue import pretrained_embedding
        return pretrained_embedding(vocab, embedding_dim, pretrained_file, amr=amr, dump_file=dump_file)
    return Embedding(vocab.size, embedding_dim, vocab.padding_idx)


class RelationEncoder(nn.Module):
    def __init__(self, vocab, rel_dim, embed_dim, hidden_size, num_layers, dropout, bidirectional=True):
        super().__init__()
        self.vocab = vocab
        self.embed_dim = embed_dim
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.dropout = dropout
        self.bidirectional = bidirectional
        self.rel_embed = AMREmbedding(vocab, rel_dim)
        # nn.GRU is only the parameter container (reference names rnn.weight_ih_l0[_reverse] ...); the
        # recurrence itself runs on the tcgen05 GEMM + gate kernels of libgtos_b200.so.
        self.rnn = nn.GRU(input_size=rel_dim, hidden_size=hidden_size, num_layers=num_layers,
                          dropout=self.dropout if num_layers > 1 else 0., bidirectional=bidirectional)
        tot_dim = 2 * hidden_size if bidirectional else hidden_size
        self.out_proj = nn.Linear(tot_dim, embed_dim)
        # [number of paths longer than t for t in range(Lmax)] of the NEXT forward call, if the caller knows it on the host
        # (the data loader builds relation_length there): lets the packed-sequence schedule skip finished paths without
        # reading the lengths back from the device.  Consumed (reset to None) by forward.
        self.row_counts = None

    def reset_parameters(self):
        nn.init.normal_(self.out_proj.weight, std=0.02)
        nn.init.constant_(self.out_proj.bias, 0.)

    def _gru_weights(self):
        ws = []
        for l in range(self.num_layers):
            for sfx in (("", "_reverse") if self.bidirectional else ("",)):
                ws += [getattr(self.rnn, f"weight_ih_l{l}{sfx}"), getattr(self.rnn, f"weight_hh_l{l}{sfx}"),
                       getattr(self.rnn, f"bias_ih_l{l}{sfx}"), getattr(self.rnn, f"bias_hh_l{l}{sfx}")]
        return ws

    def forward(self, src_tokens, src_lengths):
        """src_tokens [Lmax, R] int64, src_lengths [R] int64 -> [R, embed_dim]  (encoder.py:90-119).
        No host sync: lengths stay on the device (the reference calls .tolist(), encoder.py:99)."""
        p = self.dropout if self.training else 0.0
        if ops.fp32_mode() or not self.bidirectional:
            # fp32 mode; also the unidirectional encoder (encoder.py:67,83 - gtos itself always builds the bidirectional one,
            # generator.py:27): it runs on the general (any number of directions) fp32-mode Function in both modes
            from . import ops32
            self.row_counts = None
            bank = ops32.GRUBank32Fn.apply(src_tokens, src_lengths, self.rel_embed.weight, self.out_proj.weight,
                                           self.out_proj.bias, self.num_layers, self.hidden_size, float(p),
                                           *self._gru_weights())
            return ops.as_bank_tensor(bank)
        counts, self.row_counts = self.row_counts, None
        if not _LEN_SORT:
            counts = None
        elif counts is None and src_tokens.is_cuda and not torch.cuda.is_current_stream_capturing():
            # like the reference (encoder.py:99 reads every length back for pack_padded_sequence), but Lmax integers
            counts = ops.gru_row_counts(src_lengths, src_tokens.shape[0])
        bank = ops.GRUBankFn.apply(src_tokens, src_lengths, self.rel_embed.weight, self.out_proj.weight,
                                   self.out_proj.bias, self.num_layers, self.hidden_size, float(p), counts,
                                   *self._gru_weights())
        # a tensor whose dim-0 index_select (the caller's own gather, generator.py:79) runs on this library's kernels
        return ops.as_bank_tensor(bank)
