"""B200-native drop-in for ``RelationEncoder`` of the reference's ``encoder.py`` (generator/encoder.py:66-119):
embeds every distinct shortest-path label sequence of the batch, runs a 2-layer bidirectional GRU over the
packed sequences and projects the final states to the relation bank [R, embed_dim].

The TokenEncoder / CNN / Highway embedding front-end of the same reference file is out of the hot path
 and is not re-implemented here.
"""
import torch
from torch import nn

from . import ops
from .transformer import Embedding


def AMREmbedding(vocab, embedding_dim, pretrained_file=None, amr=False, dump_file=None):
    """reference: encoder.py:9-64; only the randomly initialised branch is part of the hot path."""
    if pretrained_file is not None:
        raise NotImplementedError("pretrained embedding files belong to the reference's data pipeline")
    return Embedding(vocab.size, embedding_dim, vocab.padding_idx)


class RelationEncoder(nn.Module):
    def __init__(self, vocab, rel_dim, embed_dim, hidden_size, num_layers, dropout, bidirectional=True):
        super().__init__()
        if not bidirectional:
            raise NotImplementedError("gtos always builds the bidirectional RelationEncoder (generator.py:27)")
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
        tot_dim = 2 * hidden_size
        self.out_proj = nn.Linear(tot_dim, embed_dim)

    def reset_parameters(self):
        nn.init.normal_(self.out_proj.weight, std=0.02)
        nn.init.constant_(self.out_proj.bias, 0.)

    def _gru_weights(self):
        ws = []
        for l in range(self.num_layers):
            for sfx in ("", "_reverse"):
                ws += [getattr(self.rnn, f"weight_ih_l{l}{sfx}"), getattr(self.rnn, f"weight_hh_l{l}{sfx}"),
                       getattr(self.rnn, f"bias_ih_l{l}{sfx}"), getattr(self.rnn, f"bias_hh_l{l}{sfx}")]
        return ws

    def forward(self, src_tokens, src_lengths):
        """src_tokens [Lmax, R] int64, src_lengths [R] int64 -> [R, embed_dim]  (encoder.py:90-119).
        No host sync: lengths stay on the device (the reference calls .tolist(), encoder.py:99)."""
        p = self.dropout if self.training else 0.0
        bank = ops.GRUBankFn.apply(src_tokens, src_lengths, self.rel_embed.weight, self.out_proj.weight,
                                   self.out_proj.bias, self.num_layers, self.hidden_size, float(p),
                                   *self._gru_weights())
        # a tensor whose dim-0 index_select (the caller's own gather, generator.py:79) runs on this library's kernels
        return ops.as_bank_tensor(bank)
