"""SURVEY.md §8 f-3 (batch construction) on CPU:
  * oracle/paths_oracle.py against the golden run of the reference's own AMRGraph + batchify
    (tests/golden/golden_paths.json, made by tests/golden/make_golden_paths.py);
  * the kernel source itself (gtos_b200/csrc/graph_paths_core.h), compiled as plain C++ with every barrier-separated
    phase run as a loop over 128 emulated threads (tests/emu/graph_paths_emu.cpp), against the oracle bit for bit;
  * the host logic of gtos_b200/paths.py (adjacency packing, bank / index assembly) - torch index arithmetic that runs on
    any device.
The CUDA build of the same kernel is compared with the same oracle in tests/test_gpu_paths.py."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import paths_oracle as PO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 19940117


@pytest.fixture(scope="module")
def gold():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_paths.json")))
    voc = g["relation_vocab"]
    graphs = [[[(u, voc[l]) for u, l in a] for a in gr["adjacency"]] for gr in g["graphs"]]
    return g, voc, graphs


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "graph_paths_emu.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "emu", "graph_paths_emu.cpp")])
    lib = C.CDLL(so)
    lib.emu_graph_paths.restype = C.c_int
    lib.emu_graph_paths.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 6 + [C.c_uint64, C.c_void_p, C.c_void_p]

    def run(n_nodes, deg, nbr, lab, max_len, self_id, tl_id, seed):
        B, n_max, deg_max = nbr.shape
        paths = np.full((B, n_max, n_max, max_len), -7, dtype=np.int32)
        plen = np.full((B, n_max, n_max), -7, dtype=np.int32)
        arrs = [np.ascontiguousarray(x, dtype=np.int32) for x in (n_nodes, deg, nbr, lab)]
        assert lib.emu_graph_paths(*[x.ctypes.data for x in arrs], B, n_max, deg_max, max_len, self_id, tl_id,
                                   seed & PO.M64, paths.ctypes.data, plen.ctypes.data) == 0
        return paths, plen

    return run


def _ids(voc):
    return voc["<CLS>"], voc["<rCLS>"], voc["<SELF>"], voc["<TL>"]


def test_oracle_enumerates_the_reference_paths(gold):
    g, voc, graphs = gold
    for gr, adj in zip(g["graphs"], graphs):
        n = len(adj)
        for i in range(n):
            for j in range(n):
                ref = sorted(tuple(voc[l] for l in p) for p in gr["all_paths"][i][j])
                assert sorted(PO.all_shortest_label_paths(adj, i, j)) == ref, (i, j)


def test_oracle_sampler_draws_reference_paths_uniformly(gold):
    g, voc, graphs = gold
    cls_id, rcls_id, self_id, tl_id = _ids(voc)
    n_nodes, deg, nbr, lab = PO.pack_adjacency(graphs)
    for seed in (SEED, SEED + 1):
        paths, plen = PO.sample_paths(n_nodes, deg, nbr, lab, 8, self_id, tl_id, seed)
        again, _ = PO.sample_paths(n_nodes, deg, nbr, lab, 8, self_id, tl_id, seed)
        assert np.array_equal(paths, again)                              # reproducible from the seed
        for b, (gr, adj) in enumerate(zip(g["graphs"], graphs)):
            n = len(adj)
            assert (plen[b, n:, :] == 0).all() and (plen[b, :, n:] == 0).all()
            for i in range(n):
                for j in range(n):
                    ref = [tuple(voc[l] for l in p) for p in gr["all_paths"][i][j]]
                    got = tuple(int(x) for x in paths[b, i, j, :plen[b, i, j]])
                    if len(ref[0]) == 0:
                        assert got == (self_id,)                         # data.py:151-152
                    elif len(ref[0]) > 8:
                        assert got == (tl_id,)                           # data.py:153-154
                    else:
                        assert got in ref
    # uniform over NODE paths: graph 4 has pairs with 3 equally short paths
    b = 4
    adj = graphs[b]
    n = len(adj)
    multi = [(i, j) for i in range(n) for j in range(n) if len(g["graphs"][b]["all_paths"][i][j]) == 3]
    assert multi
    i, j = multi[0]
    ref = [tuple(voc[l] for l in p) for p in g["graphs"][b]["all_paths"][i][j]]
    one = PO.pack_adjacency([adj])
    counts = {}
    draws = 600
    for s in range(draws):
        paths, plen = PO.sample_paths(*one, 8, self_id, tl_id, 1000 + s)
        got = tuple(int(x) for x in paths[0, i, j, :plen[0, i, j]])
        counts[got] = counts.get(got, 0) + 1
    assert sum(counts.values()) == draws and set(counts) <= set(ref)
    for p in set(ref):                                                   # multiplicity / 3 each, 5 sigma wide
        expect = draws * ref.count(p) / 3.0
        assert abs(counts.get(p, 0) - expect) < 5 * (expect * (1 - ref.count(p) / 3.0)) ** 0.5 + 1


def test_oracle_assembly_equals_reference_batchify(gold):
    g, voc, graphs = gold
    cls_id, rcls_id, self_id, tl_id = _ids(voc)
    chosen = []
    for gr in g["graphs"]:
        n = len(gr["nodes"])
        per = []
        for i in range(n):
            row = []
            for j in range(n):
                p = [voc[l] for l in gr["all_paths"][i][j][0]]           # the golden run replaced random.choice by "first"
                if len(p) == 0:
                    p = [self_id]
                if len(p) > 8:
                    p = [tl_id]
                row.append(tuple(p))
            per.append(row)
        chosen.append(per)
    rel, bank, length = PO.assemble_first_seen(chosen, cls_id, rcls_id, self_id)
    ref = g["batchify_first_choice"]
    assert np.array_equal(rel, np.array(ref["relation"]))
    assert np.array_equal(bank, np.array(ref["relation_bank"]))
    assert np.array_equal(length, np.array(ref["relation_length"]))


def _layered(width, layers, n_labels, rng):
    """complete bipartite connections between consecutive layers: width ** (layers - 1) shortest paths end to end"""
    n = width * layers
    adj = [[] for _ in range(n)]
    for l in range(layers - 1):
        for a in range(width):
            for c in range(width):
                u, v = l * width + a, (l + 1) * width + c
                k = int(rng.integers(n_labels))
                adj[u].append((v, 6 + 2 * k))
                adj[v].append((u, 7 + 2 * k))
    return adj


def test_kernel_source_emulated_on_cpu_equals_oracle(gold, emu):
    g, voc, graphs = gold
    cls_id, rcls_id, self_id, tl_id = _ids(voc)
    # golden graphs, padded shapes (n_max and deg_max larger than any graph needs)
    n_nodes, deg, nbr, lab = PO.pack_adjacency(graphs, n_max=14, deg_max=6)
    for seed, max_len in ((SEED, 8), (SEED + 5, 4), (0, 8), ((1 << 63) + 12345, 8)):
        want = PO.sample_paths(n_nodes, deg, nbr, lab, max_len, self_id, tl_id, seed)
        got = emu(n_nodes, deg, nbr, lab, max_len, self_id, tl_id, seed)
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    # random larger graphs with many re-entrancies, more nodes than emulated threads
    rng = np.random.default_rng(SEED)
    big = []
    for n in (150, 40, 97):
        adj = [dict() for _ in range(n)]
        for v in range(1, n):
            u = int(rng.integers(max(0, v - 6), v))
            k = int(rng.integers(20))
            adj[u][v], adj[v][u] = 6 + 2 * k, 7 + 2 * k
        for _ in range(n // 2):
            u, v = int(rng.integers(n)), int(rng.integers(n))
            if u != v:
                k = int(rng.integers(20))
                adj[u][v], adj[v][u] = 6 + 2 * k, 7 + 2 * k
        big.append([list(a.items()) for a in adj])
    packed = PO.pack_adjacency(big)
    want = PO.sample_paths(*packed, 8, self_id, tl_id, SEED)
    got = emu(*packed, 8, self_id, tl_id, SEED)
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    assert (want[1] > 1).any() and (want[0][..., 0] == tl_id).any()


def test_path_counts_beyond_float_range_are_rescaled(emu):
    """4 ** 63 = 8.5e37 shortest paths end to end: the per-level rescaling keeps the counts finite and the draw valid"""
    rng = np.random.default_rng(3)
    adj = _layered(4, 64, 5, rng)
    packed = PO.pack_adjacency([adj])
    want = PO.sample_paths(*packed, 16, 4, 5, 77)
    got = emu(*packed, 16, 4, 5, 77)
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    # pairs at distance <= 16 carry real paths whose labels are edges of the graph, one layer per step
    paths, plen = got
    n = len(adj)
    lab_of = {(u, v): l for u in range(n) for v, l in adj[u]}
    checked = 0
    for i in range(0, n, 37):
        for j in range(n):
            d = abs(i // 4 - j // 4)
            if 0 < d <= 16:
                assert plen[0, i, j] == d
                labels = [int(x) for x in paths[0, i, j, :d]]
                step = 1 if j > i else -1
                layer = i // 4
                ok_nodes = {i}
                for s, l in enumerate(labels):                       # some node of the next layer is reached by label l
                    nxt = {v for u in ok_nodes for v in range((layer + step) * 4, (layer + step) * 4 + 4)
                           if lab_of.get((u, v)) == l}
                    assert nxt
                    ok_nodes, layer = nxt, layer + step
                assert j in ok_nodes or d > 1
                checked += 1
    assert checked > 100


def test_pack_adjacency_and_assembly_host_logic(gold):
    from gtos_b200 import paths as P
    g, voc, graphs = gold
    cls_id, rcls_id, self_id, tl_id = _ids(voc)
    with pytest.raises(ValueError):
        P.pack_adjacency([[[(1, 6)], []]])                               # edge 0 -> 1 without its twin
    n_nodes, deg, nbr, lab = P.pack_adjacency([[[(1, 6), (1, 8)], [(0, 7)]]])
    assert deg.tolist() == [[1, 1]] and lab[0, 0, 0].item() == 8          # repeated neighbour: last label wins
    n_nodes, deg, nbr, lab = P.pack_adjacency(graphs)
    o = PO.pack_adjacency(graphs)
    for a, b in zip((n_nodes, deg, nbr, lab), o):
        assert np.array_equal(a.numpy(), b)
    paths, plen = PO.sample_paths(*o, 8, self_id, tl_id, SEED)
    out = P.assemble_relation_batch(torch.from_numpy(paths), torch.from_numpy(plen), n_nodes, cls_id, rcls_id, self_id)
    rel, bank, length = out["relation"], out["relation_bank"], out["relation_length"]
    chosen = [[[tuple(int(x) for x in paths[b, i, j, :plen[b, i, j]]) for j in range(len(adj))] for i in range(len(adj))]
              for b, adj in enumerate(graphs)]
    rel_o, bank_o, length_o = PO.assemble_first_seen(chosen, cls_id, rcls_id, self_id)
    # same tensors up to a permutation of the bank rows >= 3
    assert rel.shape == rel_o.shape and bank.shape == bank_o.shape and length.shape == length_o.shape
    assert bank[:, :3].tolist() == bank_o[:, :3].tolist() and length[:3].tolist() == [1, 1, 1]
    seq = lambda bk, ln, r: tuple(int(x) for x in bk[:int(ln[r]), r])
    N = rel.shape[0]
    for b in range(len(graphs)):
        for x in range(N):
            for y in range(N):
                assert seq(bank, length, int(rel[y, x, b])) == seq(bank_o, length_o, int(rel_o[y, x, b]))
    assert len({seq(bank, length, r) for r in range(bank.shape[1])}) == bank.shape[1]      # no duplicate rows
    assert sorted(seq(bank, length, r) for r in range(bank.shape[1])) == sorted(seq(bank_o, length_o, r) for r in range(bank_o.shape[1]))


def test_paths_product_has_no_cpu_path():
    from gtos_b200 import _lib, paths as P
    z = torch.zeros(1, 2, dtype=torch.int32)
    with pytest.raises(_lib.GtosLibraryError):
        P.shortest_label_paths(torch.zeros(1, dtype=torch.int32), z, torch.zeros(1, 2, 1, dtype=torch.int32),
                               torch.zeros(1, 2, 1, dtype=torch.int32), 4, 4, 5)


# ---------------------------------------------------------------------------------------------------------------------
# evaluation batches: every shortest path of a pair (data.py:176-225)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu_all(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu_all") / "graph_paths_emu.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "emu", "graph_paths_emu.cpp")])
    lib = C.CDLL(so)
    lib.emu_graph_all_paths.restype = C.c_int
    lib.emu_graph_all_paths.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 7 + [C.c_void_p, C.c_void_p]

    def run(n_nodes, deg, nbr, lab, max_len, K, self_id, tl_id):
        B, n_max, deg_max = nbr.shape
        allp = np.full((B, n_max, n_max, K, max_len), -7, dtype=np.int32)
        cnt = np.full((B, n_max, n_max), -7, dtype=np.int32)
        arrs = [np.ascontiguousarray(x, dtype=np.int32) for x in (n_nodes, deg, nbr, lab)]
        assert lib.emu_graph_all_paths(*[x.ctypes.data for x in arrs], B, n_max, deg_max, max_len, K, self_id, tl_id,
                                       allp.ctypes.data, cnt.ctypes.data) == 0
        return allp, cnt

    return run


def _golden_all_chosen(g, voc, self_id, tl_id):
    out = []
    for gr in g["graphs"]:
        n = len(gr["nodes"])
        per = []
        for i in range(n):
            row = []
            for j in range(n):
                ps = [[voc[l] for l in p] for p in gr["all_paths"][i][j]]
                if len(ps[0]) == 0 or len(ps[0]) > 8:                    # data.py:197-199
                    ps = ps[:1]
                row.append([tuple([self_id] if len(p) == 0 else [tl_id] if len(p) > 8 else p) for p in ps])
            per.append(row)
        out.append(per)
    return out


def test_oracle_eval_assembly_equals_reference_batchify(gold):
    g, voc, graphs = gold
    cls_id, rcls_id, self_id, tl_id = _ids(voc)
    rel, bank, length = PO.assemble_eval_first_seen(_golden_all_chosen(g, voc, self_id, tl_id), voc["<PAD>"], cls_id, rcls_id,
                                                    self_id)
    ref = g["batchify_eval"]
    assert np.array_equal(rel, np.array(ref["relation"]))
    assert np.array_equal(bank, np.array(ref["relation_bank"]))
    assert np.array_equal(length, np.array(ref["relation_length"]))


def test_all_paths_kernel_source_emulated_on_cpu_equals_oracle(gold, emu_all):
    g, voc, graphs = gold
    cls_id, rcls_id, self_id, tl_id = _ids(voc)
    packed = PO.pack_adjacency(graphs, n_max=14, deg_max=6)
    for max_len, K in ((8, 4), (4, 3), (8, 1)):                          # K = 1: pairs with 2-3 paths saturate at K + 1
        want = PO.enumerate_paths(graphs, max_len, K, self_id, tl_id, n_max=14)
        got = emu_all(*packed, max_len, K, self_id, tl_id)
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    assert (want[1] == 2).any()
    # the enumerated sets are the reference's (as multisets of label sequences), pair by pair
    allp, cnt = emu_all(*packed, 8, 4, self_id, tl_id)
    for b, (gr, adj) in enumerate(zip(g["graphs"], graphs)):
        for i in range(len(adj)):
            for j in range(len(adj)):
                ref = sorted(tuple(voc[l] for l in p) for p in gr["all_paths"][i][j])
                if len(ref[0]) == 0:
                    ref = [(self_id,)]
                elif len(ref[0]) > 8:
                    ref = [(tl_id,)]
                got = sorted(tuple(int(x) for x in allp[b, i, j, k] if x != 0) for k in range(cnt[b, i, j]))
                assert got == ref
    # many re-entrancies, more nodes than emulated threads
    rng = np.random.default_rng(SEED + 2)
    n = 140
    adj = [dict() for _ in range(n)]
    for v in range(1, n):
        u = int(rng.integers(max(0, v - 5), v))
        k = int(rng.integers(20))
        adj[u][v], adj[v][u] = 6 + 2 * k, 7 + 2 * k
    for _ in range(n):
        u, v = int(rng.integers(n)), int(rng.integers(n))
        if u != v:
            k = int(rng.integers(20))
            adj[u][v], adj[v][u] = 6 + 2 * k, 7 + 2 * k
    big = [[list(a.items()) for a in adj]]
    want = PO.enumerate_paths(big, 6, 8, self_id, tl_id)
    got = emu_all(*PO.pack_adjacency(big), 6, 8, self_id, tl_id)
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    assert want[1].max() >= 3


def test_eval_assembly_host_logic(gold):
    from gtos_b200 import paths as P
    g, voc, graphs = gold
    cls_id, rcls_id, self_id, tl_id = _ids(voc)
    K = 4
    allp, cnt = PO.enumerate_paths(graphs, 8, K, self_id, tl_id)
    n_nodes = torch.tensor([len(a) for a in graphs], dtype=torch.int32)
    out = P.assemble_eval_relation_batch(torch.from_numpy(allp), torch.from_numpy(cnt), n_nodes, voc["<PAD>"], cls_id, rcls_id,
                                         self_id)
    rel, bank, length = out["relation"], out["relation_bank"], out["relation_length"]
    ref = g["batchify_eval"]
    rel_o, bank_o, length_o = np.array(ref["relation"]), np.array(ref["relation_bank"]), np.array(ref["relation_length"])
    assert tuple(rel.shape) == rel_o.shape and tuple(bank.shape) == bank_o.shape
    assert bank[:, :4].tolist() == bank_o[:, :4].tolist() and length[:4].tolist() == length_o[:4].tolist()
    seq = lambda bk, ln, r: tuple(int(x) for x in bk[:int(ln[r]), r])
    N = rel.shape[0]
    for b in range(len(graphs)):
        for x in range(N):
            for y in range(N):
                mine = sorted(seq(bank, length, int(r)) for r in rel[y, x, b] if int(r) != 0)
                theirs = sorted(seq(bank_o, length_o, int(r)) for r in rel_o[y, x, b] if int(r) != 0)
                assert mine == theirs, (b, x, y)
    assert sorted(seq(bank, length, r) for r in range(bank.shape[1])) == sorted(seq(bank_o, length_o, r) for r in range(bank_o.shape[1]))


def test_translator_dependency_trees_are_reproduced_exactly(gold, emu):
    """translator/dependencyGraph.py:54-74 keeps the one path nx.single_source_shortest_path returns; in a tree it is the
    only shortest path, so enumeration, the uniform draw (any seed) and the emulated kernel must all return exactly it."""
    g, _, _ = gold
    trees = g["translator_trees"]
    labels = sorted({l for t in trees for a in t["adjacency"] for _, l in a})
    voc = {l: 6 + k for k, l in enumerate(labels)}
    graphs = [[[(u, voc[l]) for u, l in a] for a in t["adjacency"]] for t in trees]
    packed = PO.pack_adjacency(graphs)
    for seed in (1, SEED):
        paths, plen = PO.sample_paths(*packed, 8, 4, 5, seed)
        got = emu(*packed, 8, 4, 5, seed)
        assert np.array_equal(got[0], paths) and np.array_equal(got[1], plen)
        for b, (t, adj) in enumerate(zip(trees, graphs)):
            n = len(adj)
            for i in range(n):
                for j in range(n):
                    ref = tuple(voc[l] for l in t["paths"][i][j])
                    assert PO.all_shortest_label_paths(adj, i, j) == [ref]
                    want = (4,) if len(ref) == 0 else (5,) if len(ref) > 8 else ref      # translator/data.py:151-155
                    assert tuple(int(x) for x in paths[b, i, j, :plen[b, i, j]]) == want


def test_pack_edges_matches_pack_adjacency(gold):
    """vectorised edge-list packing: same graph as the adjacency-list packer up to the order of a node's neighbours (which
    the path distribution does not depend on), last label wins, asymmetric input is rejected"""
    from gtos_b200 import paths as P
    g, voc, graphs = gold
    gi, src, dst, lab = [], [], [], []
    for b, adj in enumerate(graphs):
        for v, a in enumerate(adj):
            for u, l in a:
                gi.append(b); src.append(v); dst.append(u); lab.append(l)
    n_nodes = [len(a) for a in graphs]
    t = P.pack_edges(n_nodes, gi, src, dst, lab)
    ref = P.pack_adjacency(graphs)
    assert torch.equal(t[0], ref[0]) and torch.equal(t[1], ref[1])
    for b, adj in enumerate(graphs):
        for v in range(len(adj)):
            d = int(t[1][b, v])
            assert sorted(zip(t[2][b, v, :d].tolist(), t[3][b, v, :d].tolist())) == sorted((int(u), int(l)) for u, l in adj[v])
    # same shortest paths (as sets) whichever neighbour order
    a1 = [[(int(u), int(l)) for u, l in zip(t[2][1, v, :int(t[1][1, v])].tolist(), t[3][1, v, :int(t[1][1, v])].tolist())]
          for v in range(n_nodes[1])]
    for i in range(n_nodes[1]):
        for j in range(n_nodes[1]):
            assert sorted(PO.all_shortest_label_paths(a1, i, j)) == sorted(PO.all_shortest_label_paths(graphs[1], i, j))
    dup = P.pack_edges([2], [0, 0, 0], [0, 0, 1], [1, 1, 0], [6, 8, 7])
    assert dup[1].tolist() == [[1, 1]] and dup[3][0, 0, 0].item() == 8
    with pytest.raises(ValueError):
        P.pack_edges([2], [0], [0], [1], [6])
    with pytest.raises(ValueError):
        P.pack_edges([2], [0, 0], [0, 2], [2, 0], [6, 7])
    empty = P.pack_edges([1, 1], [], [], [], [])
    assert empty[1].tolist() == [[0], [0]] and tuple(empty[2].shape) == (2, 1, 1)


def test_emulated_kernel_draw_is_uniform_over_the_reference_lists(gold, emu):
    """random.choice over the reference's enumerated list (data.py:150) = every NODE path equally likely.  3000 seeds through
    the kernel source: for every pair with more than one shortest path, each label sequence must come up in proportion to
    the number of node paths that spell it (5 sigma of the binomial)."""
    g, voc, graphs = gold
    cls_id, rcls_id, self_id, tl_id = _ids(voc)
    packed = PO.pack_adjacency(graphs)
    draws = 3000
    counts = {}
    for s in range(draws):
        paths, plen = emu(*packed, 8, self_id, tl_id, 7_000_000 + 977 * s)
        for b, gr in enumerate(g["graphs"]):
            n = len(gr["nodes"])
            for i in range(n):
                for j in range(n):
                    if len(gr["all_paths"][i][j]) > 1:
                        key = (b, i, j, tuple(int(x) for x in paths[b, i, j, :plen[b, i, j]]))
                        counts[key] = counts.get(key, 0) + 1
    pairs = 0
    for b, gr in enumerate(g["graphs"]):
        n = len(gr["nodes"])
        for i in range(n):
            for j in range(n):
                ref = [tuple(voc[l] for l in p) for p in gr["all_paths"][i][j]]
                if len(ref) < 2:
                    continue
                pairs += 1
                assert sum(c for k, c in counts.items() if k[:3] == (b, i, j)) == draws
                for p in set(ref):
                    q = ref.count(p) / len(ref)
                    got = counts.get((b, i, j, p), 0)
                    assert abs(got - draws * q) < 5 * (draws * q * (1 - q)) ** 0.5 + 1, (b, i, j, p, got, q)
    assert pairs >= 30


def test_adjacency_recovered_from_the_reference_json_item(gold):
    """the reference's JSON keeps enumerated paths, not edges (extract.py:173-180): the edges are the one-label paths"""
    from gtos_b200 import paths as P
    g, voc, graphs = gold
    for gr, adj in zip(g["graphs"], graphs):
        n = len(gr["nodes"])
        item = dict(concept=gr["nodes"],
                    relation={str(i): {str(j): [dict(edge=p, length=len(p)) for p in gr["all_paths"][i][j]] for j in range(n)}
                              for i in range(n)})
        got = P.adjacency_from_reference_item(item, lambda t: voc[t])
        assert [sorted(a) for a in got] == [sorted((int(u), int(l)) for u, l in a) for a in adj]


def test_bfs_order_oracle_emulation_and_relabelling(gold, tmp_path):
    """AMRGraph.bfs (AMRGraph.py:82-98): node order, depths and connectivity of the reference's own run; the kernel source
    (one thread per graph) equals it; the relabelled adjacency equals the adjacency the reference's relations are indexed by"""
    from gtos_b200 import paths as P
    g, voc, graphs = gold
    so = str(tmp_path / "emu.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "emu", "graph_paths_emu.cpp")])
    lib = C.CDLL(so)
    lib.emu_graph_bfs.restype = C.c_int
    lib.emu_graph_bfs.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p] * 4
    orig = [[[(u, voc[l]) for u, l in a] for a in gr["orig_adjacency"]] for gr in g["graphs"]]
    for gr, adj in zip(g["graphs"], orig):
        order, depths, ok = PO.bfs_order(adj, gr["root"])
        assert order == gr["bfs_order"] and depths == gr["bfs_depths"] and ok
        assert [sorted(a) for a in PO.relabel(adj, order)] == [sorted((u, voc[l]) for u, l in a) for a in gr["adjacency"]]
    # a disconnected graph: node 3 has no edges
    orig.append([[(1, 6)], [(0, 7), (2, 6)], [(1, 7)], []])
    roots = [gr["root"] for gr in g["graphs"]] + [1]
    n_nodes, deg, nbr, lab = PO.pack_adjacency(orig, n_max=14, deg_max=6)
    B = len(orig)
    root = np.array(roots, dtype=np.int32)
    order, depth, pos = (np.full((B, 14), -9, dtype=np.int32) for _ in range(3))
    reached = np.full(B, -9, dtype=np.int32)
    assert lib.emu_graph_bfs(n_nodes.ctypes.data, deg.ctypes.data, nbr.ctypes.data, root.ctypes.data, B, 14, 6, order.ctypes.data,
                             depth.ctypes.data, pos.ctypes.data, reached.ctypes.data) == 0
    for b, adj in enumerate(orig):
        want_o, want_d, ok = PO.bfs_order(adj, roots[b])
        m = len(want_o)
        assert reached[b] == m and (reached[b] == len(adj)) == ok
        assert order[b, :m].tolist() == want_o and depth[b, :m].tolist() == want_d
        assert (order[b, m:] == -1).all() and (depth[b, m:] == 0).all()
        for v in range(14):
            assert pos[b, v] == (want_o.index(v) if v in want_o else -1)
    assert reached[-1] == 3 and pos[-1, 3] == -1
    # relabelling on tensors == the oracle's relabel == the adjacency the golden relations are indexed by (connected graphs)
    t = [torch.from_numpy(x[:-1].copy()) for x in (deg, nbr, lab, order, pos)]
    deg2, nbr2, lab2 = P.relabel_adjacency(*t)
    for b, gr in enumerate(g["graphs"]):
        for k in range(len(gr["nodes"])):
            d = int(deg2[b, k])
            got = sorted(zip(nbr2[b, k, :d].tolist(), lab2[b, k, :d].tolist()))
            assert got == sorted((u, voc[l]) for u, l in gr["adjacency"][k])
        assert int(deg2[b, len(gr["nodes"]):].sum()) == 0


def test_empty_and_degenerate_batches(emu):
    """an empty graph inside a batch, a batch of single nodes, two isolated nodes (unreachable pair -> <TL>)"""
    from gtos_b200 import paths as P
    graphs = [[], [[]], [[], []], [[(1, 6)], [(0, 7)]]]
    packed = PO.pack_adjacency(graphs, n_max=3, deg_max=2)
    paths, plen = emu(*packed, 4, 4, 5, 1)
    want = PO.sample_paths(*packed, 4, 4, 5, 1)
    assert np.array_equal(paths, want[0]) and np.array_equal(plen, want[1])
    assert (plen[0] == 0).all() and (paths[0] == 0).all()                               # empty graph: all padding
    assert plen[1, 0, 0] == 1 and paths[1, 0, 0, 0] == 4 and plen[1].sum() == 1          # single node: <SELF>
    assert paths[2, 0, 1, 0] == 5 and paths[2, 1, 0, 0] == 5                             # isolated nodes: <TL>
    assert paths[3, 0, 1].tolist() == [6, 0, 0, 0] and paths[3, 1, 0].tolist() == [7, 0, 0, 0]
    n_nodes = torch.tensor([0, 1, 2, 2], dtype=torch.int32)
    out = P.assemble_relation_batch(torch.from_numpy(paths), torch.from_numpy(plen), n_nodes, 2, 3, 4)
    rel, bank, length = out["relation"], out["relation_bank"], out["relation_length"]
    assert bank[0, :3].tolist() == [2, 3, 4] and sorted(bank[0, 3:].tolist()) == [5, 6, 7] and length.tolist() == [1] * 6
    assert rel[:, :, 0].tolist() == [[2, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]]   # only the <SELF> corner
    assert rel[1, 1, 1].item() == 2 and rel[0, 1, 1].item() == 1 and rel[1, 0, 1].item() == 0
    tl = (bank[0] == 5).nonzero().item()
    assert rel[2, 1, 2].item() == tl and rel[1, 2, 2].item() == tl
    none = P.assemble_relation_batch(torch.zeros(0, 3, 3, 4, dtype=torch.int32), torch.zeros(0, 3, 3, dtype=torch.int32),
                                     torch.zeros(0, dtype=torch.int32), 2, 3, 4)
    assert tuple(none["relation"].shape) == (4, 4, 0) and none["relation_bank"].shape[1] == 3


def test_unique_keys_equals_torch_unique_rows():
    from gtos_b200 import paths as P
    gen = torch.Generator().manual_seed(5)
    for n, span in ((0, 5), (1, 5), (2000, 7), (5000, 1 << 40)):
        hi = torch.randint(0, span, (n,), generator=gen)
        lo = torch.randint(0, span, (n,), generator=gen)
        uniq, inv = P._unique_keys(hi, lo)
        if n == 0:
            assert tuple(uniq.shape) == (0, 2) and inv.numel() == 0
            continue
        ref_u, ref_i = torch.unique(torch.stack([hi, lo], 1), dim=0, return_inverse=True)
        assert torch.equal(uniq, ref_u) and torch.equal(inv, ref_i)
