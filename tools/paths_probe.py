#!/usr/bin/env python
"""Time SURVEY.md §8 f-3 on a B200: gtos_graph_bfs / gtos_graph_paths / gtos_graph_all_paths and the bank / index assembly at
the config-2 / 3 / 4 batch sizes, next to the host pipeline they replace (the reference's networkx enumeration is stood in
for by this repo's BFS-per-source host code, gtos_b200/synthetic.py, which is already far cheaper than all_shortest_paths).
Also checks the drawn paths against the CPU oracle on a sample of the batch.  First thing to run next round:
    GTOS_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_paths.py -q && python tools/paths_probe.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtos_b200 import paths as P, synthetic          # noqa: E402
from oracle import paths_oracle as PO                # noqa: E402

CLS, RCLS, SELF, TL = synthetic.CLS, synthetic.RCLS, synthetic.SELF, synthetic.TL


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = torch.device("cuda:0")
    vocab = synthetic.RelVocab(100)
    for name, B, n_max, max_len in (("cfg2", 64, 40, 4), ("cfg3 shard", 16, 60, 8), ("cfg4 shard", 32, 256, 8)):
        rng = np.random.default_rng(19940117)
        graphs = []
        t0 = time.perf_counter()
        for b in range(B):
            n = n_max if b == 0 else int(rng.integers(n_max // 2, n_max + 1))
            adj = synthetic._random_graph(n, rng, vocab)
            graphs.append([list(dict(a).items()) for a in adj])
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        for adj in graphs[:8]:
            for i in range(len(adj)):
                synthetic._shortest_label_paths(adj, i)
        t_host = (time.perf_counter() - t0) / 8 * B
        t0 = time.perf_counter()
        n_nodes, deg, nbr, lab = P.pack_adjacency(graphs, device=dev)
        t_pack = time.perf_counter() - t0
        root = torch.zeros(B, dtype=torch.int32, device=dev)
        seed = torch.tensor([19940117], dtype=torch.int64, device=dev)
        us_bfs = timeit(lambda: P.bfs_order(n_nodes, deg, nbr, root))
        us_paths = timeit(lambda: P.shortest_label_paths(n_nodes, deg, nbr, lab, max_len, SELF, TL, seed=seed, seed_off=0))
        paths, plen = P.shortest_label_paths(n_nodes, deg, nbr, lab, max_len, SELF, TL, seed=seed, seed_off=0)
        us_asm = timeit(lambda: P.assemble_relation_batch(paths, plen, n_nodes, CLS, RCLS, SELF), n=5)
        K = 8
        us_all = timeit(lambda: P.all_shortest_label_paths(n_nodes, deg, nbr, lab, max_len, K, SELF, TL), n=5)
        out = P.assemble_relation_batch(paths, plen, n_nodes, CLS, RCLS, SELF)
        # oracle check on the first two graphs
        sub = PO.pack_adjacency(graphs[:2], n_max=nbr.shape[1], deg_max=nbr.shape[2])
        want = PO.sample_paths(*sub, max_len, SELF, TL, 19940117)
        ok = np.array_equal(paths[:2].cpu().numpy(), want[0]) and np.array_equal(plen[:2].cpu().numpy(), want[1])
        pairs = int((n_nodes.to(torch.int64) ** 2).sum())
        print(f"{name:11s} B={B:3d} n<={n_max:3d}: bfs {us_bfs:7.1f} us | paths {us_paths:8.1f} us ({pairs / us_paths:6.1f} M pairs/s) | "
              f"all paths (K={K}) {us_all:8.1f} us | assemble {us_asm:8.1f} us (R={out['relation_bank'].shape[1]}) | "
              f"host: BFS per source {t_host * 1e3:7.1f} ms, pack {t_pack * 1e3:6.1f} ms, generate {t_gen * 1e3:6.1f} ms | "
              f"oracle match: {ok}")


if __name__ == "__main__":
    main()
