"""the relation part of batch construction on the GPU.

The reference builds it on the host: `AMRGraph.collect_concepts_and_relations` enumerates every shortest path between
every ordered pair of nodes with networkx (generator/AMRGraph.py:100-115; offline, stored as JSON by extract.py:173-180),
and `batchify` draws one of them per pair each time a batch is formed, substitutes <SELF> / <TL>, and de-duplicates the
label sequences of the batch into `relation_bank` [Lmax, R], `relation_length` [R] and the index tensor
`relation` [N, N, B] (generator/data.py:126-176; translator/data.py:134-177 does the same from dependency trees).

Here the graphs travel as a padded adjacency (`pack_adjacency`), `shortest_label_paths` runs gtos_graph_paths - one CTA per
(graph, target), uniform draw among equally short paths by path counting - and `assemble_relation_batch` turns the
per-pair label sequences into the three tensors with device-side index arithmetic (sort / unique; one host read for R).
The bank comes out in sorted-key order instead of first-seen order: a permutation of the bank rows with the index tensor
permuted consistently, which no consumer can observe (RelationEncoder encodes rows independently, generator.py:76-79
gathers by index); rows 0 / 1 / 2 stay <CLS> / <rCLS> / <SELF> as data.py:138-140 fixes them.

There is no CPU path: `shortest_label_paths` needs CUDA tensors and libgtos_b200.so.
"""
import torch

from . import _lib
from .ops import _need_cuda, _p, _st, rng_state


def pack_adjacency(graphs, n_max=None, deg_max=None, device=None):
    """graphs: list (one per graph) of adjacency lists adj[v] = [(u, label id), ...] in the node order the batch uses
    (BFS order, AMRGraph.py:82-98).  A repeated neighbour keeps its LAST label (networkx DiGraph.add_edge overwrites,
    AMRGraph.py:79-80).  Checks that the structure is symmetric.  Returns int32 tensors (n_nodes [B], deg [B,n_max],
    nbr [B,n_max,deg_max], lab [B,n_max,deg_max])."""
    clean = []
    for adj in graphs:
        g = []
        for a in adj:
            last = {}
            for u, l in a:
                last[int(u)] = int(l)
            g.append(list(last.items()))
        for v, a in enumerate(g):
            for u, _ in a:
                if not 0 <= u < len(g) or all(w != v for w, _ in g[u]):
                    raise ValueError(f"adjacency is not symmetric at edge {v} -> {u}: every edge needs its reverse twin "
                                     "(AMRGraph._add_edge, AMRGraph.py:76-80)")
        clean.append(g)
    B = len(clean)
    n_max = n_max or max((len(g) for g in clean), default=1)
    deg_max = deg_max or max(1, max((len(a) for g in clean for a in g), default=1))
    n_nodes = torch.tensor([len(g) for g in clean], dtype=torch.int32)
    deg = torch.zeros(B, n_max, dtype=torch.int32)
    nbr = torch.zeros(B, n_max, deg_max, dtype=torch.int32)
    lab = torch.zeros(B, n_max, deg_max, dtype=torch.int32)
    for b, g in enumerate(clean):
        for v, a in enumerate(g):
            deg[b, v] = len(a)
            for k, (u, l) in enumerate(a):
                nbr[b, v, k], lab[b, v, k] = u, l
    out = (n_nodes, deg, nbr, lab)
    return tuple(t.to(device) for t in out) if device is not None else out


def adjacency_from_reference_item(item, token2idx):
    """The reference's preprocessed JSON (generator/extract.py:173-180) stores, per graph, `relation[i][j]` = the list of
    all shortest paths i -> j with their edge labels - not the edges.  The edges are the paths of length one:
    adj[i] = [(j, id of the label of i -> j)].  `item` = one entry of that JSON (keys may be str after the JSON round trip,
    data.py:149), `token2idx` = vocabs['relation'].token2idx (data.py:48-51)."""
    rel = item["relation"]
    n = len(item["concept"])
    adj = [[] for _ in range(n)]
    for i in range(n):
        row = rel[str(i)] if str(i) in rel else rel[i]
        for j in range(n):
            paths = row[str(j)] if str(j) in row else row[j]
            edge = paths[0]["edge"]
            if len(edge) == 1:                      # a direct edge: every shortest path i -> j is that edge
                adj[i].append((j, int(token2idx(edge[0]))))
    return adj


def pack_edges(n_nodes, graph, src, dst, label, n_max=None, deg_max=None, device=None):
    """Vectorised pack_adjacency for large batches.  One row per DIRECTED edge in insertion order: graph[e], src[e], dst[e],
    label[e] (the caller lists both the edge and its reverse twin, as AMRGraph._add_edge does).  A pair (src, dst) listed
    more than once keeps its LAST label.  Returns the same int32 tensors as pack_adjacency; neighbours of a node are
    ordered by destination index."""
    import numpy as np
    n_nodes = np.asarray(n_nodes, dtype=np.int64)
    g, u, v, l = (np.asarray(x, dtype=np.int64) for x in (graph, src, dst, label))
    B = len(n_nodes)
    n_max = int(n_max or (n_nodes.max() if B else 1))
    if len(g):
        if (u < 0).any() or (v < 0).any() or (u >= n_nodes[g]).any() or (v >= n_nodes[g]).any():
            raise ValueError("edge endpoint outside its graph")
        key = (g * n_max + u) * n_max + v
        order = np.lexsort((np.arange(len(key)), key))                  # by (graph, src, dst), insertion order inside
        key, l = key[order], l[order]
        last = np.ones(len(key), dtype=bool)
        last[:-1] = key[1:] != key[:-1]                                 # last occurrence of every (graph, src, dst)
        key, l = key[last], l[last]
        rev = (key // (n_max * n_max)) * n_max * n_max + (key % n_max) * n_max + (key // n_max) % n_max
        if not np.isin(rev, key).all():
            raise ValueError("adjacency is not symmetric: every edge needs its reverse twin (AMRGraph._add_edge, "
                             "AMRGraph.py:76-80)")
        node = key // n_max                                             # graph * n_max + src
        start = np.r_[True, node[1:] != node[:-1]]
        first = np.flatnonzero(start)
        slot = np.arange(len(key)) - np.repeat(first, np.diff(np.r_[first, len(key)]))
        dmax = int(slot.max()) + 1
    else:
        key = l = node = slot = np.zeros(0, dtype=np.int64)
        dmax = 1
    deg_max = int(deg_max or dmax)
    if dmax > deg_max:
        raise ValueError(f"a node has {dmax} neighbours, deg_max is {deg_max}")
    deg = np.zeros(B * n_max, dtype=np.int32)
    np.add.at(deg, node, 1)
    nbr = np.zeros((B * n_max, deg_max), dtype=np.int32)
    lab = np.zeros((B * n_max, deg_max), dtype=np.int32)
    nbr[node, slot] = key % n_max
    lab[node, slot] = l
    out = (torch.from_numpy(n_nodes.astype(np.int32)), torch.from_numpy(deg).view(B, n_max),
           torch.from_numpy(nbr).view(B, n_max, deg_max), torch.from_numpy(lab).view(B, n_max, deg_max))
    return tuple(t.to(device) for t in out) if device is not None else out


def bfs_order(n_nodes, deg, nbr, root):
    """AMRGraph.bfs on the device (AMRGraph.py:82-98): -> (order [B,n_max], depth [B,n_max], pos [B,n_max], reached [B]),
    int32.  order[b][k] = k-th node of the root's BFS queue, depth[b][k] = its depth (`concept_depth`, data.py:129),
    pos = inverse permutation (-1: not reached), reached[b] == n_nodes[b] iff graph b is connected."""
    _need_cuda(n_nodes, deg, nbr, root)
    for t in (n_nodes, deg, nbr, root):
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise ValueError("bfs_order takes contiguous int32 tensors (pack_adjacency)")
    B, n_max, deg_max = nbr.shape
    dev = nbr.device
    order, depth, pos = (torch.empty(B, n_max, dtype=torch.int32, device=dev) for _ in range(3))
    reached = torch.empty(B, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().gtos_graph_bfs(_p(n_nodes), _p(deg), _p(nbr), _p(root), B, n_max, deg_max, _p(order), _p(depth),
                                          _p(pos), _p(reached), _st()), "graph_bfs")
    return order, depth, pos, reached


def relabel_adjacency(deg, nbr, lab, order, pos):
    """the padded adjacency renumbered by a node order (index arithmetic, any device): node order[b][k] becomes node k.
    Graphs must be connected (every node has a position)."""
    B, n_max, deg_max = nbr.shape
    o = order.to(torch.int64).clamp(min=0)                                             # [B, n_max]; -1 rows are padding
    live = (order >= 0)
    deg2 = torch.where(live, torch.gather(deg.to(torch.int64), 1, o), torch.zeros_like(o)).to(torch.int32)
    rows = o.unsqueeze(-1).expand(B, n_max, deg_max)
    nbr_old = torch.gather(nbr.to(torch.int64), 1, rows)                               # neighbours of node order[b][k], old ids
    lab2 = torch.gather(lab.to(torch.int64), 1, rows)
    used = torch.arange(deg_max, device=nbr.device).view(1, 1, deg_max) < deg2.unsqueeze(-1).to(torch.int64)
    nbr2 = torch.gather(pos.to(torch.int64), 1, nbr_old.reshape(B, -1)).reshape(B, n_max, deg_max)
    zero = torch.zeros_like(nbr2)
    return deg2, torch.where(used, nbr2, zero).to(torch.int32), torch.where(used, lab2, zero).to(torch.int32)


def _check_graph(n_nodes, deg, nbr):
    """GTOS_CHECK_GRAPH=1: range check of a padded adjacency that did not come from pack_adjacency / pack_edges (the
    kernels index shared memory with these values and trust them).  Costs a device read-back."""
    import os
    if os.environ.get("GTOS_CHECK_GRAPH") != "1":
        return
    B, n_max, deg_max = nbr.shape
    n = n_nodes.to(torch.int64).view(B, 1)
    if int((n_nodes < 0).any() | (n_nodes > n_max).any()):
        raise ValueError("n_nodes outside [0, n_max]")
    inside = torch.arange(n_max, device=nbr.device).view(1, n_max) < n
    if int(((deg < 0) | (deg > deg_max))[inside].any()):
        raise ValueError("a node degree is outside [0, deg_max]")
    used = torch.arange(deg_max, device=nbr.device).view(1, 1, deg_max) < deg.to(torch.int64).unsqueeze(-1)
    bad = ((nbr < 0) | (nbr.to(torch.int64) >= n.view(B, 1, 1))) & used & inside.unsqueeze(-1)
    if int(bad.any()):
        raise ValueError("a neighbour index is outside its graph")


def shortest_label_paths(n_nodes, deg, nbr, lab, max_len, self_id, tl_id, seed_off=None, seed=None):
    """-> (paths [B,n_max,n_max,max_len] int32, plen [B,n_max,n_max] int32); paths[b,i,j] = labels of the drawn shortest
    path i -> j (AMRGraph.py:107-112 + data.py:150-154).  `seed`: int64 device tensor (default: the library's dropout
    seed, ops.rng_state), `seed_off`: per-call offset - the draw is reproducible from (seed + seed_off, b, i, j); None
    (default) takes a fresh offset per call, like the reference's fresh random.choice per batch (data.py:150)."""
    _need_cuda(n_nodes, deg, nbr, lab)
    _check_graph(n_nodes, deg, nbr)
    if seed_off is None:
        from .ops import new_seed_off
        seed_off = new_seed_off()
    for t in (n_nodes, deg, nbr, lab):
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise ValueError("shortest_label_paths takes contiguous int32 tensors (pack_adjacency)")
    B, n_max, deg_max = nbr.shape
    dev = nbr.device
    seed = rng_state(dev) if seed is None else seed
    paths = torch.empty(B, n_max, n_max, max_len, dtype=torch.int32, device=dev)
    plen = torch.empty(B, n_max, n_max, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().gtos_graph_paths(_p(n_nodes), _p(deg), _p(nbr), _p(lab), B, n_max, deg_max, max_len, self_id, tl_id,
                                            _p(seed), seed_off & ((1 << 64) - 1), _p(paths), _p(plen), _st()), "graph_paths")
    return paths, plen


def _unique_keys(hi, lo):
    """distinct (hi, lo) pairs in lexicographic order + the index of every input pair in that list - what
    torch.unique(stack([hi, lo], 1), dim=0, return_inverse=True) returns, from two stable radix sorts instead of a
    comparator sort over rows (which is the slow part of that call on the GPU at ~1e5 keys)."""
    n = hi.numel()
    if n == 0:
        return torch.stack([hi, lo], dim=1), hi.new_zeros((0,))
    o1 = torch.sort(lo, stable=True).indices
    o2 = torch.sort(hi[o1], stable=True).indices
    order = o1[o2]                                                    # lexicographic by (hi, lo)
    h, l = hi[order], lo[order]
    new = torch.ones(n, dtype=torch.bool, device=hi.device)
    new[1:] = (h[1:] != h[:-1]) | (l[1:] != l[:-1])
    gid = torch.cumsum(new.to(torch.int64), 0) - 1                    # group of every sorted element
    inv = torch.empty(n, dtype=torch.int64, device=hi.device)
    inv[order] = gid
    return torch.stack([h[new], l[new]], dim=1), inv


def assemble_relation_batch(paths, plen, n_nodes, cls_id, rcls_id, self_id):
    """per-pair label sequences -> dict(relation [N,N,B] int64, relation_bank [Lmax,R] int64, relation_length [R] int64),
    N = n_max + 1 (the <CLS> slot), with the layouts of data.py:138-176: relation[j+1][i+1][b] = bank row of the path
    i -> j, relation[0][0][b] = 2 (<SELF>), relation[x][0][b] = 0 (<CLS>) and relation[0][y][b] = 1 (<rCLS>) for nodes
    inside graph b, 0 elsewhere (ArraysToTensor padding).  Index arithmetic only (runs where the tensors live)."""
    B, n_max, _, L = paths.shape
    dev = paths.device
    if L > 8:
        raise ValueError("label sequences of more than 8 labels are replaced by <TL> in the reference (data.py:153)")
    p64 = paths.to(torch.int64)
    if p64.numel() and int(p64.max()) >= (1 << 15):
        raise ValueError("relation label ids must be < 32768")
    # 120-bit key of a sequence: four 15-bit labels per word, first label most significant (0 = padding sorts first)
    pad = torch.zeros(B, n_max, n_max, 8 - L, dtype=torch.int64, device=dev)
    p8 = torch.cat([p64, pad], dim=-1)
    hi = (p8[..., 0] << 45) | (p8[..., 1] << 30) | (p8[..., 2] << 15) | p8[..., 3]
    lo = (p8[..., 4] << 45) | (p8[..., 5] << 30) | (p8[..., 6] << 15) | p8[..., 7]
    valid = plen > 0
    is_self = valid & (plen == 1) & (p64[..., 0] == self_id)
    real = valid & ~is_self
    uniq, inv = _unique_keys(hi[real], lo[real])                                    # sorted keys; dynamic size: host read
    R = 3 + uniq.shape[0]
    ids = torch.zeros(B, n_max, n_max, dtype=torch.int64, device=dev)
    ids[is_self] = 2
    ids[real] = inv + 3
    N = n_max + 1
    rel = torch.zeros(B, N, N, dtype=torch.int64, device=dev)                       # brs[b][x][y], data.py:142-161
    inside = torch.arange(n_max, device=dev).unsqueeze(0) < n_nodes.to(torch.int64).unsqueeze(1)      # [B, n_max]
    rel[:, 0, 0] = 2
    rel[:, 0, 1:] = 0                                                               # <CLS> id 0 (also the padding value)
    rel[:, 1:, 0] = inside.to(torch.int64)                                          # <rCLS> id 1 for real nodes
    rel[:, 1:, 1:] = ids
    relation = rel.permute(2, 1, 0).contiguous()                                    # transpose_(0, 2), data.py:164
    bank = torch.zeros(8, R, dtype=torch.int64, device=dev)
    bank[0, 0], bank[0, 1], bank[0, 2] = cls_id, rcls_id, self_id
    if uniq.shape[0]:
        u_hi, u_lo = uniq[:, 0], uniq[:, 1]
        cols = [(u_hi >> 45) & 0x7FFF, (u_hi >> 30) & 0x7FFF, (u_hi >> 15) & 0x7FFF, u_hi & 0x7FFF,
                (u_lo >> 45) & 0x7FFF, (u_lo >> 30) & 0x7FFF, (u_lo >> 15) & 0x7FFF, u_lo & 0x7FFF]
        bank[:, 3:] = torch.stack(cols, dim=0)
    length = (bank != 0).sum(0)
    Lmax = max(1, int(length.max()))
    return dict(relation=relation, relation_bank=bank[:Lmax].contiguous(), relation_length=length)


def all_shortest_label_paths(n_nodes, deg, nbr, lab, max_len, K, self_id, tl_id):
    """-> (all_paths [B,n_max,n_max,K,max_len] int32, pcount [B,n_max,n_max] int32): every shortest path of every pair
    (evaluation batches, data.py:176-225), at most K per pair; raises if a pair has more (one host read)."""
    _need_cuda(n_nodes, deg, nbr, lab)
    for t in (n_nodes, deg, nbr, lab):
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise ValueError("all_shortest_label_paths takes contiguous int32 tensors (pack_adjacency)")
    B, n_max, deg_max = nbr.shape
    dev = nbr.device
    all_paths = torch.empty(B, n_max, n_max, K, max_len, dtype=torch.int32, device=dev)
    pcount = torch.empty(B, n_max, n_max, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().gtos_graph_all_paths(_p(n_nodes), _p(deg), _p(nbr), _p(lab), B, n_max, deg_max, max_len, K, self_id,
                                                tl_id, _p(all_paths), _p(pcount), _st()), "graph_all_paths")
    worst = int(pcount.max()) if pcount.numel() else 0
    if worst > K:
        raise ValueError(f"a node pair has more than K={K} shortest paths; call again with a larger K")
    return all_paths, pcount


def assemble_eval_relation_batch(all_paths, pcount, n_nodes, pad_id, cls_id, rcls_id, self_id):
    """all paths per pair -> dict(relation [N,N,B,Kb] int64, relation_bank [Lmax,R] int64, relation_length [R] int64) with
    the layouts of the evaluation branch of batchify (data.py:176-225): bank rows 0..3 = <PAD>, <CLS>, <rCLS>, <SELF>;
    relation[j+1][i+1][b][k] = bank row of the k-th shortest path i -> j, 0 (<PAD>) beyond the pair's count; Kb = the
    largest count in the batch.  Bank rows >= 4 come out in sorted-key order (a permutation of the reference's first-seen
    order; generator.py:83-88 only gathers rows by index and averages)."""
    B, n_max, _, K, L = all_paths.shape
    dev = all_paths.device
    if L > 8:
        raise ValueError("label sequences of more than 8 labels are replaced by <TL> in the reference (data.py:203)")
    p64 = all_paths.to(torch.int64)
    if p64.numel() and int(p64.max()) >= (1 << 15):
        raise ValueError("relation label ids must be < 32768")
    pad = torch.zeros(B, n_max, n_max, K, 8 - L, dtype=torch.int64, device=dev)
    p8 = torch.cat([p64, pad], dim=-1)
    hi = (p8[..., 0] << 45) | (p8[..., 1] << 30) | (p8[..., 2] << 15) | p8[..., 3]
    lo = (p8[..., 4] << 45) | (p8[..., 5] << 30) | (p8[..., 6] << 15) | p8[..., 7]
    cnt = pcount.to(torch.int64).clamp(max=K)
    present = torch.arange(K, device=dev).view(1, 1, 1, K) < cnt.unsqueeze(-1)                 # [B,n,n,K]
    is_self = present & (p64[..., 0] == self_id) & (p64[..., 1:].sum(-1) == 0) if L > 1 else present & (p64[..., 0] == self_id)
    real = present & ~is_self
    uniq, inv = _unique_keys(hi[real], lo[real])
    R = 4 + uniq.shape[0]
    ids = torch.zeros(B, n_max, n_max, K, dtype=torch.int64, device=dev)
    ids[is_self] = 3
    ids[real] = inv + 4
    Kb = max(1, int(cnt.max())) if cnt.numel() else 1
    N = n_max + 1
    rel = torch.zeros(B, N, N, Kb, dtype=torch.int64, device=dev)                              # [b][x][y][k], data.py:194-219
    inside = (torch.arange(n_max, device=dev).unsqueeze(0) < n_nodes.to(torch.int64).unsqueeze(1)).to(torch.int64)
    rel[:, 0, 0, 0] = 3                                                                        # <SELF>
    rel[:, 0, 1:, 0] = inside                                                                  # <CLS> id 1 for real nodes
    rel[:, 1:, 0, 0] = 2 * inside                                                              # <rCLS> id 2
    rel[:, 1:, 1:, :] = ids[..., :Kb]
    relation = rel.permute(2, 1, 0, 3).contiguous()                                            # transpose_(0, 2), data.py:221
    bank = torch.zeros(8, R, dtype=torch.int64, device=dev)
    bank[0, 0], bank[0, 1], bank[0, 2], bank[0, 3] = pad_id, cls_id, rcls_id, self_id
    if uniq.shape[0]:
        u_hi, u_lo = uniq[:, 0], uniq[:, 1]
        cols = [(u_hi >> 45) & 0x7FFF, (u_hi >> 30) & 0x7FFF, (u_hi >> 15) & 0x7FFF, u_hi & 0x7FFF,
                (u_lo >> 45) & 0x7FFF, (u_lo >> 30) & 0x7FFF, (u_lo >> 15) & 0x7FFF, u_lo & 0x7FFF]
        bank[:, 4:] = torch.stack(cols, dim=0)
    length = (bank != 0).sum(0)
    length[0] = 1                                                                              # the <PAD> row is (pad_id,), length 1 (data.py:183)
    Lmax = max(1, int(length.max()))
    return dict(relation=relation, relation_bank=bank[:Lmax].contiguous(), relation_length=length)


def relation_batch(graphs, max_len, cls_id, rcls_id, self_id, tl_id, device, seed_off=None):
    """adjacency lists -> the three relation tensors of a training batch (data.py:134-176), paths drawn on the GPU."""
    n_nodes, deg, nbr, lab = pack_adjacency(graphs, device=device)
    paths, plen = shortest_label_paths(n_nodes, deg, nbr, lab, max_len, self_id, tl_id, seed_off=seed_off)
    return assemble_relation_batch(paths, plen, n_nodes, cls_id, rcls_id, self_id)
