"""Synthetic AMR-shaped batches with the exact tensor dictionary the reference's data layer emits
(generator/data.py:126-267), so the hot path can be driven without the licensed corpora.

Per graph: a random rooted tree plus ~10 % re-entrancy edges, every edge labelled from a relation
vocabulary with `_reverse_` twins (generator/AMRGraph.py:79-80), nodes in BFS order (AMRGraph.py:82-98),
all-pairs shortest label paths (AMRGraph.py:100-115) with `[] -> <SELF>` and over-long paths `-> <TL>`
(data.py:151-154), a prepended <CLS> row/column (data.py:138-147), and the distinct paths of the batch
de-duplicated into a bank [Lmax, R] + lengths [R] + index tensor idx[x][y][b] = path y -> x (data.py:164-176).
Seed 19940117 is the reference's own (generator/train.py:98).
"""
import math
from collections import deque

import numpy as np
import torch

SEED = 19940117
PAD, UNK, CLS, RCLS, SELF, TL = 0, 1, 2, 3, 4, 5
N_SPECIAL = 6


class RelVocab:
    """stand-in for data.Vocab with the attributes the modules read (.size/.padding_idx/.unk_idx)."""

    def __init__(self, n_labels=100):
        self.n_labels = n_labels
        self.size = N_SPECIAL + 2 * n_labels
        self.padding_idx, self.unk_idx = PAD, UNK

    def label(self, k, reverse=False):
        return N_SPECIAL + 2 * k + (1 if reverse else 0)

    def idx2token(self, i):
        return f"rel{i}"


class TokenVocab:
    def __init__(self, size):
        self.size, self.padding_idx, self.unk_idx = size, 0, 1

    def idx2token(self, i):
        return f"tok{i}"


def _random_graph(n, rng, vocab):
    """adjacency list of (neighbour, label id) for a BFS-ordered random DAG-ish graph on n nodes."""
    adj = [[] for _ in range(n)]

    def add(u, v):
        k = int(rng.integers(vocab.n_labels))
        adj[u].append((v, vocab.label(k)))
        adj[v].append((u, vocab.label(k, reverse=True)))

    for v in range(1, n):                       # node ids are already a valid BFS order of this tree
        lo = max(0, v - 1 - int(rng.integers(0, 4)))
        add(int(rng.integers(lo // 2, v)), v)
    for _ in range(max(0, n // 10)):            # re-entrancies
        u, v = int(rng.integers(n)), int(rng.integers(n))
        if u != v:
            add(u, v)
    return adj


def _shortest_label_paths(adj, src):
    """BFS from src; returns for every node the label sequence of one shortest path src -> node."""
    n = len(adj)
    path = [None] * n
    path[src] = ()
    dq = deque([src])
    while dq:
        u = dq.popleft()
        for v, lab in adj[u]:
            if path[v] is None:
                path[v] = path[u] + (lab,)
                dq.append(v)
    return path


def make_graphs(B, n_max, n_labels=100, seed=SEED, full=False):
    """the synthetic graphs alone: (adjacency lists, node counts, vocabulary) - same graphs as make_graph_batch draws"""
    rng = np.random.default_rng(seed)
    vocab = RelVocab(n_labels)
    graphs, counts = [], []
    for b in range(B):
        n = n_max if (b == 0 or full) else int(rng.integers(math.ceil(0.5 * n_max), n_max + 1))
        counts.append(n)
        graphs.append(_random_graph(n, rng, vocab))
    return graphs, counts, vocab


def edge_arrays(graphs):
    """adjacency lists -> (graph, src, dst, label) rows, one per directed edge in insertion order (paths.pack_edges)"""
    g_, u_, v_, l_ = [], [], [], []
    for b, adj in enumerate(graphs):
        for u, a in enumerate(adj):
            for v, lab in a:
                g_.append(b); u_.append(u); v_.append(v); l_.append(lab)
    return (np.asarray(g_, dtype=np.int64), np.asarray(u_, dtype=np.int64), np.asarray(v_, dtype=np.int64),
            np.asarray(l_, dtype=np.int64))


def make_graph_batch_device(B, n_max, device, max_path_len=4, n_labels=100, seed=SEED, full=False, seed_off=0):
    """make_graph_batch with the relation tensors built ON THE GPU: the synthetic graphs travel as a
    padded adjacency, gtos_graph_paths draws one shortest label path per ordered pair, assemble_relation_batch
    de-duplicates them into bank / lengths / index.  Returns the same dictionary (tensors on the host)."""
    from . import paths as P
    graphs, counts, vocab = make_graphs(B, n_max, n_labels, seed, full)
    packed = P.pack_edges(counts, *edge_arrays(graphs), n_max=n_max, device=device)
    sl, pl = P.shortest_label_paths(*packed, max_path_len, SELF, TL, seed_off=seed_off)
    out = P.assemble_relation_batch(sl, pl, packed[0], CLS, RCLS, SELF)
    return dict(relation_bank=out["relation_bank"].cpu(), relation_length=out["relation_length"].cpu(),
                relation=out["relation"].cpu(), node_counts=torch.tensor(counts), N=n_max + 1, rel_vocab=vocab,
                packed_adjacency=packed)


def make_graph_batch(B, n_max, max_path_len=4, n_labels=100, seed=SEED, full=False):
    """Returns dict(relation_bank [Lmax,R] int64, relation_length [R] int64, relation [N,N,B] int64,
    node_counts [B], N) with N = n_max + 1 (the <CLS> slot)."""
    rng = np.random.default_rng(seed)
    vocab = RelVocab(n_labels)
    N = n_max + 1
    bank = {(CLS,): 0, (RCLS,): 1, (SELF,): 2}
    idx = np.zeros((B, N, N), dtype=np.int64)          # brs[b][x][y] as built by data.py:139-160
    counts = []
    for b in range(B):
        n = n_max if (b == 0 or full) else int(rng.integers(math.ceil(0.5 * n_max), n_max + 1))
        counts.append(n)
        adj = _random_graph(n, rng, vocab)
        idx[b, 0, 0] = 2
        idx[b, 0, 1:n + 1] = 0                         # <CLS> row (data.py:143)
        for i in range(n):
            paths = _shortest_label_paths(adj, i)
            idx[b, i + 1, 0] = 1                       # <rCLS> column (data.py:146)
            for j in range(n):
                p = paths[j]
                if p is None or len(p) > max_path_len:
                    p = (TL,)
                elif len(p) == 0:
                    p = (SELF,)
                r = bank.get(p)
                if r is None:
                    r = bank[p] = len(bank)
                idx[b, i + 1, j + 1] = r
    R = len(bank)
    Lmax = max(len(k) for k in bank)
    bank_t = np.zeros((Lmax, R), dtype=np.int64)
    lengths = np.zeros(R, dtype=np.int64)
    for k, v in bank.items():
        bank_t[:len(k), v] = k
        lengths[v] = len(k)
    rel = torch.from_numpy(idx).permute(2, 1, 0).contiguous()      # transpose_(0, 2): [y][x][b] (data.py:164)
    return dict(relation_bank=torch.from_numpy(bank_t), relation_length=torch.from_numpy(lengths), relation=rel,
                node_counts=torch.tensor(counts), N=N, rel_vocab=vocab)


def make_batch(B, n_max, D, T_max=60, T_min=20, V=10000, max_path_len=4, seed=SEED, full=False, device_paths=None):
    """Graph batch + the dense inputs of the hot path: node features x [N,B,D] = LayerNorm(randn),
    padding mask [N,B], teacher-forced token states [T,B,D], token padding mask, copy_seq [N-1,B], target [T,B].
    device_paths = a CUDA device: the relation tensors come from the GPU path builder (make_graph_batch_device)."""
    if device_paths is not None:
        g = make_graph_batch_device(B, n_max, device_paths, max_path_len=max_path_len, seed=seed, full=full)
    else:
        g = make_graph_batch(B, n_max, max_path_len=max_path_len, seed=seed, full=full)
    gen = torch.Generator().manual_seed(seed)
    N = g["N"]
    x = torch.nn.functional.layer_norm(torch.randn(N, B, D, generator=gen), (D,))
    node_mask = torch.arange(N).unsqueeze(1) >= (g["node_counts"] + 1).unsqueeze(0)          # [N,B] True = pad
    t_len = torch.randint(T_min, T_max + 1, (B,), generator=gen)
    t_len[0] = T_max
    T = T_max
    tok = torch.nn.functional.layer_norm(torch.randn(T, B, D, generator=gen), (D,))
    tok_mask = torch.arange(T).unsqueeze(1) >= t_len.unsqueeze(0)
    copy_seq = torch.randint(2, V + 16, (N - 1, B), generator=gen)
    target = torch.randint(2, V, (T, B), generator=gen).masked_fill(tok_mask, 0)       # always < tot_ext
    g.update(x=x, node_mask=node_mask, token_repr=tok, token_mask=tok_mask, copy_seq=copy_seq, target=target, T=T,
             t_len=t_len, V=V)
    return g
