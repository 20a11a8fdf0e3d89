#!/usr/bin/env python
"""Dry run of tests/test_gpu_paths.py WITHOUT a GPU: gtos_b200.paths.{shortest_label_paths, all_shortest_label_paths} are
replaced by the CPU emulation of the kernel source (tests/emu/graph_paths_emu.cpp, compiled here with g++), so the test
LOGIC (oracle equality, full-batch properties, bank / index round trips) is exercised end to end.  It says nothing about
the CUDA build - that is what the gated tests are for (GTOS_TEST_EXPERIMENTAL=1 on a B200)."""
import subprocess, tempfile
import ctypes as C, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["GTOS_TEST_EXPERIMENTAL"] = "1"
from oracle import paths_oracle as PO
from gtos_b200 import paths as P
SO = os.path.join(tempfile.mkdtemp(), "graph_paths_emu.so")
subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", SO, os.path.join(ROOT, "tests", "emu", "graph_paths_emu.cpp")])
lib = C.CDLL(SO)
lib.emu_graph_paths.restype = C.c_int
lib.emu_graph_paths.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 6 + [C.c_uint64, C.c_void_p, C.c_void_p]
def fake(n_nodes, deg, nbr, lab, max_len, self_id, tl_id, seed_off=0, seed=None):
    B, n_max, deg_max = nbr.shape
    arrs = [np.ascontiguousarray(t.cpu().numpy(), dtype=np.int32) for t in (n_nodes, deg, nbr, lab)]
    paths = np.zeros((B, n_max, n_max, max_len), dtype=np.int32); plen = np.zeros((B, n_max, n_max), dtype=np.int32)
    s = (int(seed.item()) if seed is not None else 0) + seed_off
    lib.emu_graph_paths(*[x.ctypes.data for x in arrs], B, n_max, deg_max, max_len, self_id, tl_id, s & PO.M64, paths.ctypes.data, plen.ctypes.data)
    return torch.from_numpy(paths), torch.from_numpy(plen)
P.shortest_label_paths = fake
torch.cuda.synchronize = lambda *a, **k: None
import test_gpu_paths as T
dev = torch.device("cpu")
T.test_graph_paths_equals_oracle_on_reference_graphs(dev); print("ref graphs ok")
T.test_graph_paths_full_batch_properties_and_assembly(dev); print("full batch ok")
T.test_graph_paths_equals_oracle_on_larger_graphs(dev); print("larger ok")
lib.emu_graph_all_paths.restype = C.c_int
lib.emu_graph_all_paths.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 7 + [C.c_void_p, C.c_void_p]
def fake_all(n_nodes, deg, nbr, lab, max_len, K, self_id, tl_id):
    B, n_max, deg_max = nbr.shape
    arrs = [np.ascontiguousarray(t.cpu().numpy(), dtype=np.int32) for t in (n_nodes, deg, nbr, lab)]
    allp = np.zeros((B, n_max, n_max, K, max_len), dtype=np.int32); cnt = np.zeros((B, n_max, n_max), dtype=np.int32)
    lib.emu_graph_all_paths(*[x.ctypes.data for x in arrs], B, n_max, deg_max, max_len, K, self_id, tl_id, allp.ctypes.data, cnt.ctypes.data)
    if cnt.max() > K: raise ValueError("K")
    return torch.from_numpy(allp), torch.from_numpy(cnt)
P.all_shortest_label_paths = fake_all
T.test_graph_all_paths_equals_oracle_and_eval_assembly(dev); print("all paths ok")
lib.emu_graph_bfs.restype = C.c_int
lib.emu_graph_bfs.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p] * 4
def fake_bfs(n_nodes, deg, nbr, root):
    B, n_max, deg_max = nbr.shape
    arrs = [np.ascontiguousarray(t.cpu().numpy(), dtype=np.int32) for t in (n_nodes, deg, nbr, root)]
    order, depth, pos = (np.zeros((B, n_max), dtype=np.int32) for _ in range(3)); reached = np.zeros(B, dtype=np.int32)
    lib.emu_graph_bfs(*[x.ctypes.data for x in arrs], B, n_max, deg_max, order.ctypes.data, depth.ctypes.data, pos.ctypes.data, reached.ctypes.data)
    return tuple(torch.from_numpy(x) for x in (order, depth, pos, reached))
P.bfs_order = fake_bfs
T.test_graph_bfs_equals_reference_order(dev); print("bfs ok")
