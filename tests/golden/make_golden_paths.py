"""Golden runs of the reference's OWN graph code for SURVEY.md §8 f-3: AMRGraph.bfs / collect_concepts_and_relations
(generator/AMRGraph.py:82-115, networkx all_shortest_paths) and the relation part of batchify (generator/data.py:126-176)
on small random AMR-shaped graphs.  Build container only (needs /root/reference and networkx):
    python tests/golden/make_golden_paths.py        -> tests/golden/golden_paths.json

Shims (SURVEY.md §8c style, none touches the logic under test):
  * numpy 2 removed np.int, which data.py:118,160,171 still use                   -> np.int = int
  * an AMRGraph is normally built from a smatch AMR object (AMRGraph.py:24-70); here the instance is allocated with
    __new__ and filled through the reference's own _add_edge (AMRGraph.py:76-80), which creates the `_reverse_` twins
  * batchify's random.choice (data.py:150) is replaced by "first of the list" for the deterministic bank / index run
"""
import json
import os
import random
import sys
import tempfile

import numpy as np

np.int = int
sys.path.insert(0, "/root/reference/generator")
import networkx as nx                              # noqa: E402
import AMRGraph as ref_graph                       # noqa: E402
import data as ref_data                            # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N_LABELS = 6


def random_edges(n, rng, extra):
    """(label, src name, des name): a random tree over n named nodes plus `extra` re-entrancy edges"""
    names = [f"c{k}" for k in range(n)]
    edges = []
    for v in range(1, n):
        u = rng.randrange(max(0, v - 3), v)
        edges.append((f"ARG{rng.randrange(N_LABELS)}", names[u], names[v]))
    for _ in range(extra):
        u, v = rng.randrange(n), rng.randrange(n)
        if u != v:
            edges.append((f"ARG{rng.randrange(N_LABELS)}", names[u], names[v]))
    return names, edges


def build(names, edges):
    g = ref_graph.AMRGraph.__new__(ref_graph.AMRGraph)
    g.graph = nx.DiGraph()
    g.name2concept = {n: n for n in names}
    g.root = names[0]
    for n in names:
        g.graph.add_node(n)
    for rel, src, des in edges:
        g._add_edge(rel, src, des)                 # the reference's own edge + `_reverse_` twin
    return g


def translator_trees(rng):
    """translator/dependencyGraph.py (the reference's own constructor, _add_edge with `_r_` twins, bfs and
    collect_concepts_and_relations, dependencyGraph.py:8-74): dependency trees, ONE path per pair
    (nx.single_source_shortest_path) - unique in a tree, so the GPU draw must reproduce it exactly."""
    sys.path.append("/root/reference/translator")
    import dependencyGraph as ref_dep              # noqa: E402
    out = []
    for n in (1, 6, 13, 10):
        head = [0] + [rng.randrange(1, v + 1) for v in range(1, n)]          # head[v] in 1..v (1-based), 0 = root
        dep = ["root"] + [f"dep{rng.randrange(5)}" for _ in range(1, n)]
        tok = [f"t{k}" for k in range(n)]
        g = ref_dep.dependencyGraph(dep, head, tok, ["x"])                   # the reference, unmodified
        concepts, _, relations, connected = g.collect_concepts_and_relations()
        assert connected
        order, _, _ = g.bfs()
        pos = {name: k for k, name in enumerate(order)}
        adj = [[(pos[u], g.graph[v][u]["label"]) for u in g.graph.neighbors(v)] for v in order]
        paths = [[relations[i][j][0]["edge"] for j in range(n)] for i in range(n)]
        out.append(dict(head=head, dep=dep, nodes=order, adjacency=adj, paths=paths))
    return out


def main():
    rng = random.Random(19940117)
    specs = [(5, 1), (9, 3), (12, 5), (11, 0), (7, 4), (1, 0)]
    graphs = []
    for n, extra in specs:
        names, edges = random_edges(n, rng, extra)
        graphs.append((names, edges))
    # a chain of 11 nodes: paths of up to 10 labels, so the > 8 -> <TL> rule of data.py:153 fires
    names = [f"c{k}" for k in range(11)]
    graphs.append((names, [(f"ARG{k % N_LABELS}", names[k], names[k + 1]) for k in range(10)]))

    # relation vocabulary file: every label and twin with a count >= 5 (data.py:343 Vocab(..., 5, [CLS, rCLS, SEL, TL]))
    tmp = tempfile.mkdtemp()
    labels = [f"ARG{k}" for k in range(N_LABELS)] + [f"ARG{k}_reverse_" for k in range(N_LABELS)]
    with open(os.path.join(tmp, "rel"), "w") as f:
        for l in labels:
            f.write(f"{l}\t10\n")
    with open(os.path.join(tmp, "tok"), "w") as f:
        for k in range(12):
            f.write(f"c{k}\t10\n")
        f.write("w\t10\n")
    V = ref_data.Vocab
    vocabs = dict(concept=V(os.path.join(tmp, "tok"), 5, [ref_data.CLS]), token=V(os.path.join(tmp, "tok"), 5, [ref_data.STR, ref_data.END]),
                  predictable_token=V(os.path.join(tmp, "tok"), 5, [ref_data.END]),
                  token_char=V(os.path.join(tmp, "tok"), 5, [ref_data.STR, ref_data.END]),
                  concept_char=V(os.path.join(tmp, "tok"), 5, [ref_data.STR, ref_data.END]),
                  relation=V(os.path.join(tmp, "rel"), 5, [ref_data.CLS, ref_data.rCLS, ref_data.SEL, ref_data.TL]))
    rel_vocab = {t: vocabs["relation"].token2idx(t) for t in [ref_data.PAD, ref_data.UNK, ref_data.CLS, ref_data.rCLS, ref_data.SEL,
                                                             ref_data.TL] + labels}

    out_graphs, items = [], []
    for names, edges in graphs:
        g = build(names, edges)
        concepts, depths, relations, connected = g.collect_concepts_and_relations()      # the reference, unmodified
        assert connected
        order, _, _ = g.bfs()
        pos = {name: k for k, name in enumerate(order)}
        # adjacency in BFS indices, neighbours in networkx iteration order, label = the reference's edge attribute
        adj = [[(pos[u], g.graph[v][u]["label"]) for u in g.graph.neighbors(v)] for v in order]
        n = len(order)
        all_paths = [[[p["edge"] for p in relations[i][j]] for j in range(n)] for i in range(n)]
        # the graph as the reference holds it BEFORE the BFS renumbering: node k = names[k], neighbours in networkx order
        idx = {name: k for k, name in enumerate(names)}
        orig_adj = [[(idx[u], g.graph[v][u]["label"]) for u in g.graph.neighbors(v)] for v in names]
        _, bfs_depths, _ = g.bfs()
        out_graphs.append(dict(nodes=order, edges=edges, adjacency=adj, all_paths=all_paths, orig_adjacency=orig_adj,
                               root=idx[g.root], bfs_order=[idx[v] for v in order], bfs_depths=bfs_depths))
        items.append(dict(concept=concepts, depth=depths, relation=json.loads(json.dumps(relations)), token=["w"],
                          token2idx={}, idx2token={}, cp_seq=concepts, abstract=[]))

    ref_data.random.choice = lambda xs: xs[0]                                            # deterministic choice
    batch = ref_data.batchify(items, vocabs, train=True)                                  # the reference, unmodified
    ev = ref_data.batchify(items, vocabs, train=False)                                    # evaluation branch: all paths
    out = dict(relation_vocab=rel_vocab, graphs=out_graphs,
               batchify_first_choice=dict(relation=batch["relation"].tolist(), relation_bank=batch["relation_bank"].tolist(),
                                          relation_length=batch["relation_length"].tolist()),
               batchify_eval=dict(relation=ev["relation"].tolist(), relation_bank=ev["relation_bank"].tolist(),
                                  relation_length=ev["relation_length"].tolist()))
    out["translator_trees"] = translator_trees(rng)
    with open(os.path.join(HERE, "golden_paths.json"), "w") as f:
        json.dump(out, f)
    print("graphs", [len(g["nodes"]) for g in out_graphs], "bank", len(out["batchify_first_choice"]["relation_length"]),
          "bytes", os.path.getsize(os.path.join(HERE, "golden_paths.json")))


if __name__ == "__main__":
    main()
