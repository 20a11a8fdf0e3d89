"""Small host-side helpers that the reference's callers import from ``transformer`` (generator/generator.py:8,
encoder.py:6): token embedding factory, causal mask cache, sinusoidal position table.  They produce index /
constant tensors once per step and are not part of the GPU hot path, so they stay PyTorch.
Behaviour follows generator/transformer.py:198-281."""
import math

import torch
from torch import nn


def Embedding(num_embeddings, embedding_dim, padding_idx):
    """N(0, 0.02) embedding table with a zero padding row (transformer.py:198-202)."""
    table = nn.Embedding(num_embeddings, embedding_dim, padding_idx=padding_idx)
    with torch.no_grad():
        table.weight.normal_(mean=0.0, std=0.02)
        table.weight[padding_idx].zero_()
    return table


def pretrained_embedding(vocab, embedding_dim, pretrained_file, amr=False, dump_file=None):
    """Embedding table initialised from a text file of `token v1 .. vD` lines (encoder.py:9-64): rows of vocabulary tokens
    found in the file take the file's vector (with amr=True a sense suffix `-NN` is ignored when matching), the other rows
    are drawn from N(mean, std) of the vectors that were found, the padding row is zero; the table stays trainable.
    dump_file: the matching lines are copied there.  Host-side, once per model construction."""
    import re

    def norm(tok):
        return re.sub(r"-\d\d$", "", tok) if amr else tok

    wanted = {norm(vocab.idx2token(i)) for i in range(vocab.size)}
    found = {}
    dump = open(dump_file, "w", encoding="utf8") if dump_file is not None else None
    try:
        with open(pretrained_file, encoding="utf8") as fh:
            for line in fh:
                parts = line.rstrip().split(" ")
                if len(parts) - 1 != embedding_dim or parts[0] not in wanted:
                    continue
                if dump is not None:
                    dump.write(line)
                found[parts[0]] = torch.tensor([float(v) for v in parts[1:]], dtype=torch.float32)
    finally:
        if dump is not None:
            dump.close()
    if not found:
        raise ValueError(f"{pretrained_file}: no {embedding_dim}-dimensional vector of a vocabulary token")
    stacked = torch.stack(list(found.values()))
    mean, std = float(stacked.mean()), float(stacked.std(unbiased=False))
    table = torch.empty(vocab.size, embedding_dim).normal_(mean, std)
    for i in range(vocab.size):
        tok = vocab.idx2token(i)
        vec = found.get(tok)
        if vec is None and amr:
            vec = found.get(norm(tok))
        if vec is not None:
            table[i] = vec
    table[vocab.padding_idx].zero_()
    return nn.Embedding.from_pretrained(table, freeze=False)


class SelfAttentionMask(nn.Module):
    """Cached strictly-upper-triangular mask, True = may not attend (transformer.py:204-219; bool, not uint8)."""

    def __init__(self, device, init_size=100):
        super().__init__()
        self.device = device
        self.weights = self.get_mask(init_size)

    @staticmethod
    def get_mask(size):
        idx = torch.arange(size)
        return idx.unsqueeze(0) > idx.unsqueeze(1)

    def forward(self, size):
        if self.weights is None or self.weights.size(0) < size:
            self.weights = self.get_mask(size)
        return self.weights[:size, :size].to(self.device)


class SinusoidalPositionalEmbedding(nn.Module):
    """tensor2tensor-style table [sin | cos] grown on demand (transformer.py:240-281)."""

    def __init__(self, embedding_dim, device, init_size=512):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.device = device
        self.weights = self.get_embedding(init_size, embedding_dim)

    @staticmethod
    def get_embedding(num_embeddings, embedding_dim):
        half = embedding_dim // 2
        step = math.log(10000) / (half - 1)
        inv_freq = torch.exp(-step * torch.arange(half, dtype=torch.float))
        angle = torch.outer(torch.arange(num_embeddings, dtype=torch.float), inv_freq)
        table = torch.cat((angle.sin(), angle.cos()), dim=1)
        if embedding_dim % 2:
            table = torch.nn.functional.pad(table, (0, 1))
        return table

    def forward(self, input, offset=0):
        """input: [seq_len, bsz] (only its shape is used) -> [seq_len, bsz, dim] positions offset..offset+seq_len."""
        seq_len, bsz = input.shape
        need = offset + seq_len
        if self.weights is None or self.weights.size(0) < need:
            self.weights = self.get_embedding(need, self.embedding_dim)
        rows = self.weights[offset:need]
        return rows.unsqueeze(1).expand(seq_len, bsz, rows.size(-1)).to(self.device)
