"""CPU oracle for SURVEY.md §8 f-3 (batch construction): shortest label paths of a graph batch and the relation
bank / index tensors built from them.  TEST INFRASTRUCTURE ONLY - imported by tests/ (and nothing in gtos_b200/).

Restates, in plain Python / numpy:
  * AMRGraph.collect_concepts_and_relations, generator/AMRGraph.py:100-115 (networkx all_shortest_paths over the
    bidirected labelled graph built by _add_edge, AMRGraph.py:76-80)        -> all_shortest_label_paths()
  * the per-pair choice and the <SELF> / <TL> substitutions of batchify, generator/data.py:148-154
                                                                            -> sample_paths()
  * the first-seen de-duplication into relation_bank / relation_length / relation, data.py:134-176
                                                                            -> assemble_first_seen()
Pinned by tests/golden/golden_paths.json, produced by tests/golden/make_golden_paths.py from the reference's own
AMRGraph and batchify (tests/test_paths_cpu.py).

sample_paths() draws ONE shortest path per ordered pair uniformly among all shortest node paths - the distribution of the
reference's random.choice over the enumerated list - by counting paths in a BFS and walking down the counts.  It follows
the arithmetic of gtos_graph_paths (include/gtos_b200.h) operation by operation in float32, so the CUDA kernel is held
to it bit for bit: same hash, same summation order (adjacency order), same tie rule.
"""
from collections import deque

import numpy as np

M64 = (1 << 64) - 1


def uniform(seed, idx):
    """counter-based uniform in [0,1) of the path sampler (csrc/graph_paths_core.h paths_uniform), as a float32"""
    z = (seed + idx * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z = z ^ (z >> 31)
    return np.float32(z >> 40) * np.float32(1.0 / 16777216.0)


def _bfs_dist(adj, src):
    dist = [-1] * len(adj)
    dist[src] = 0
    dq = deque([src])
    while dq:
        u = dq.popleft()
        for v, _ in adj[u]:
            if dist[v] < 0:
                dist[v] = dist[u] + 1
                dq.append(v)
    return dist


def all_shortest_label_paths(adj, i, j):
    """every shortest NODE path i -> j as its tuple of edge labels (AMRGraph.py:107-112); one entry per node path, so a
    label sequence reached through two different node paths appears twice, as in the reference's list.
    adj[v] = [(neighbour, label of v -> neighbour), ...]"""
    dj = _bfs_dist_to(adj, j)                       # distances TO j: the walk only ever steps one level closer
    if dj[i] < 0:
        return []
    out = []

    def walk(v, labels):
        if v == j:
            out.append(tuple(labels))
            return
        for u, lab in adj[v]:
            if dj[u] == dj[v] - 1:
                walk(u, labels + [lab])

    walk(i, [])
    return out


def _bfs_dist_to(adj, tgt):
    """distance v -> tgt; the structure is symmetric (every edge has a reverse twin), so a BFS from tgt gives it"""
    return _bfs_dist(adj, tgt)


def sample_paths(n_nodes, deg, nbr, lab, max_len, self_id, tl_id, seed):
    """int32 arrays n_nodes [B], deg [B,n_max], nbr / lab [B,n_max,deg_max]  ->  (paths [B,n_max,n_max,max_len] int32,
    plen [B,n_max,n_max] int32), exactly what gtos_graph_paths writes for the combined seed `seed`."""
    B, n_max, deg_max = nbr.shape
    paths = np.zeros((B, n_max, n_max, max_len), dtype=np.int32)
    plen = np.zeros((B, n_max, n_max), dtype=np.int32)
    f32 = np.float32
    for b in range(B):
        n = int(n_nodes[b])
        for j in range(n):
            dist = np.full(n_max, -1, dtype=np.int64)
            sigma = np.zeros(n_max, dtype=np.float32)
            dist[j], sigma[j] = 0, f32(1.0)
            level = 0
            frontier = [j]
            while True:
                # unvisited neighbours of the current level; each one sums sigma over ITS OWN adjacency order
                cand_nodes = sorted({int(nbr[b, u, k]) for u in frontier for k in range(int(deg[b, u]))
                                     if dist[int(nbr[b, u, k])] < 0})
                new = []
                for v in cand_nodes:
                    s = f32(0.0)
                    for k in range(int(deg[b, v])):
                        u = int(nbr[b, v, k])
                        if dist[u] == level:
                            s = f32(s + sigma[u])
                    new.append((v, s))
                if not new:
                    break
                for v, s in new:
                    dist[v] = level + 1
                    sigma[v] = s
                mx = max(s for _, s in new)
                if mx > f32(1.0e30):
                    for v, _ in new:
                        sigma[v] = f32(sigma[v] * f32(f32(1.0) / f32(1.0e30)))
                frontier = [v for v, _ in new]
                level += 1
            for i in range(n):
                d = int(dist[i])
                if d == 0:
                    paths[b, i, j, 0], plen[b, i, j] = self_id, 1
                elif d < 0 or d > max_len:
                    paths[b, i, j, 0], plen[b, i, j] = tl_id, 1
                else:
                    v = i
                    for s in range(d):
                        want = d - s - 1
                        cand = [k for k in range(int(deg[b, v])) if dist[int(nbr[b, v, k])] == want]
                        total = f32(0.0)
                        for k in cand:
                            total = f32(total + sigma[int(nbr[b, v, k])])
                        e = (((b * n_max + i) * n_max + j) * max_len + s) & M64
                        r = f32(uniform(seed & M64, e) * total)
                        cum, pick = f32(0.0), cand[-1]
                        for k in cand:
                            cum = f32(cum + sigma[int(nbr[b, v, k])])
                            if cum > r:
                                pick = k
                                break
                        paths[b, i, j, s] = lab[b, v, pick]
                        v = int(nbr[b, v, pick])
                    plen[b, i, j] = d
    return paths, plen


def pack_adjacency(graphs, n_max=None, deg_max=None):
    """graphs: list of adjacency lists adj[v] = [(u, label id), ...] (one entry per neighbour) -> int32 arrays"""
    B = len(graphs)
    n_max = n_max or max(len(g) for g in graphs)
    deg_max = deg_max or max(1, max((len(a) for g in graphs for a in g), default=1))
    n_nodes = np.array([len(g) for g in graphs], dtype=np.int32)
    deg = np.zeros((B, n_max), dtype=np.int32)
    nbr = np.zeros((B, n_max, deg_max), dtype=np.int32)
    lab = np.zeros((B, n_max, deg_max), dtype=np.int32)
    for b, g in enumerate(graphs):
        for v, a in enumerate(g):
            deg[b, v] = len(a)
            for k, (u, l) in enumerate(a):
                nbr[b, v, k], lab[b, v, k] = u, l
    return n_nodes, deg, nbr, lab


def assemble_first_seen(chosen, cls_id, rcls_id, self_id):
    """data.py:134-176 given the ALREADY CHOSEN and substituted label tuple of every pair: chosen[b][i][j] = tuple of label
    ids of the path i -> j (<SELF> / <TL> substitutions done).  Returns (relation [N,N,B], relation_bank [Lmax,R],
    relation_length [R]) as numpy int64 with the reference's first-seen bank order and layouts."""
    bank = {(cls_id,): 0, (rcls_id,): 1, (self_id,): 2}
    mats = []
    for per_graph in chosen:
        n = len(per_graph)
        brs = [[2] + [0] * n]                                           # data.py:143
        for i in range(n):
            rs = [1]                                                    # data.py:146
            for j in range(n):
                p = tuple(per_graph[i][j])
                r = bank.get(p, len(bank))
                if r == len(bank):
                    bank[p] = r
                rs.append(r)
            brs.append(rs)
        mats.append(np.array(brs, dtype=np.int64))
    N = max(m.shape[0] for m in mats)
    rel = np.zeros((len(mats), N, N), dtype=np.int64)                    # ArraysToTensor zero padding (data.py:113-124)
    for b, m in enumerate(mats):
        rel[b, :m.shape[0], :m.shape[1]] = m
    rel = rel.transpose(2, 1, 0).copy()                                  # transpose_(0, 2): relation[j][i][b] = path i -> j
    R = len(bank)
    Lmax = max(len(k) for k in bank)
    bank_t = np.zeros((Lmax, R), dtype=np.int64)
    lengths = np.zeros(R, dtype=np.int64)
    for k, v in bank.items():
        bank_t[:len(k), v] = k
        lengths[v] = len(k)
    return rel, bank_t, lengths


def enumerate_paths(graphs, max_len, K, self_id, tl_id, n_max=None):
    """what gtos_graph_all_paths writes: all_paths [B,n_max,n_max,K,max_len] int32 (depth-first adjacency order),
    pcount [B,n_max,n_max] int32 saturated at K + 1; <SELF> / <TL> pairs hold a single entry (data.py:197-199)."""
    B = len(graphs)
    n_max = n_max or max(len(g) for g in graphs)
    all_paths = np.zeros((B, n_max, n_max, K, max_len), dtype=np.int32)
    pcount = np.zeros((B, n_max, n_max), dtype=np.int32)
    for b, adj in enumerate(graphs):
        n = len(adj)
        for j in range(n):
            dj = _bfs_dist_to(adj, j)
            for i in range(n):
                d = dj[i]
                if d == 0:
                    all_paths[b, i, j, 0, 0], pcount[b, i, j] = self_id, 1
                elif d < 0 or d > max_len:
                    all_paths[b, i, j, 0, 0], pcount[b, i, j] = tl_id, 1
                else:
                    ps = all_shortest_label_paths(adj, i, j)
                    pcount[b, i, j] = min(len(ps), K + 1)
                    for k, p in enumerate(ps[:K]):
                        all_paths[b, i, j, k, :len(p)] = p
    return all_paths, pcount


def assemble_eval_first_seen(all_chosen, pad_id, cls_id, rcls_id, self_id):
    """the evaluation branch of batchify, data.py:176-225: all_chosen[b][i][j] = list of label tuples (substitutions done,
    <SELF> / <TL> pairs already cut to one entry).  Returns (relation [N,N,B,K], relation_bank [Lmax,R],
    relation_length [R]) numpy int64; bank rows 0..3 = <PAD>, <CLS>, <rCLS>, <SELF> (data.py:183-186)."""
    bank = {(pad_id,): 0, (cls_id,): 1, (rcls_id,): 2, (self_id,): 3}
    per_graph, num_concepts, num_paths = [], 0, 0
    for g in all_chosen:
        n = len(g)
        num_concepts = max(n + 1, num_concepts)
        brs = [[[3]] + [[1]] * n]                                           # data.py:194
        for i in range(n):
            rs = [[2]]                                                      # data.py:196
            for j in range(n):
                all_r = []
                for p in g[i][j]:
                    p = tuple(p)
                    r = bank.get(p, len(bank))
                    if r == len(bank):
                        bank[p] = r
                    all_r.append(r)
                num_paths = max(len(all_r), num_paths)
                rs.append(all_r)
            brs.append(rs)
        per_graph.append(brs)
    mat = np.zeros((len(per_graph), num_concepts, num_concepts, num_paths), dtype=np.int64)
    for b, x in enumerate(per_graph):
        for i, y in enumerate(x):
            for j, z in enumerate(y):
                for k, r in enumerate(z):
                    mat[b, i, j, k] = r
    rel = mat.transpose(2, 1, 0, 3).copy()                                   # transpose_(0, 2), data.py:221
    R = len(bank)
    Lmax = max(len(k) for k in bank)
    bank_t = np.zeros((Lmax, R), dtype=np.int64)
    lengths = np.zeros(R, dtype=np.int64)
    for k, v in bank.items():
        bank_t[:len(k), v] = k
        lengths[v] = len(k)
    return rel, bank_t, lengths


def bfs_order(adj, root):
    """AMRGraph.bfs, generator/AMRGraph.py:82-98: (queue order, depths, is_connected); adj[v] in neighbour iteration order"""
    queue, depths, visited = [root], [0], {root}
    step = 0
    while step < len(queue):
        u, depth = queue[step], depths[step]
        step += 1
        for v, _ in adj[u]:
            if v not in visited:
                queue.append(v)
                depths.append(depth + 1)
                visited.add(v)
    return queue, depths, len(queue) == len(adj)


def relabel(adj, order):
    """adjacency renumbered by the BFS order (what collect_concepts_and_relations indexes its relations by, AMRGraph.py:103)"""
    pos = {v: k for k, v in enumerate(order)}
    return [[(pos[u], l) for u, l in adj[v]] for v in order]
