// CPU build of gtos_b200/csrc/graph_paths_core.h - the SAME source the CUDA kernel runs, with every barrier-separated
// phase executed as a loop over 128 emulated thread ids.  Test infrastructure only (tests/test_paths_cpu.py compiles it
// with g++ and compares it with oracle/paths_oracle.py); nothing in gtos_b200/ loads it.
#include <stdint.h>

#include <vector>

#include "../../gtos_b200/csrc/graph_paths_core.h"

extern "C" int emu_graph_paths(const int32_t* n_nodes, const int32_t* deg, const int32_t* nbr, const int32_t* lab, int32_t B,
                               int32_t n_max, int32_t deg_max, int32_t max_len, int32_t self_id, int32_t tl_id, uint64_t seed,
                               int32_t* paths, int32_t* plen) {
  gtos::GraphPathsArgs a;
  a.n_nodes = n_nodes; a.deg = deg; a.nbr = nbr; a.lab = lab;
  a.B = B; a.n_max = n_max; a.deg_max = deg_max; a.max_len = max_len;
  a.self_id = self_id; a.tl_id = tl_id; a.seed = seed;
  a.paths = paths; a.plen = plen;
  std::vector<int32_t> smem((gtos::graph_paths_smem_bytes(n_max) + 3) / 4);
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < n_max; ++j) gtos::graph_paths_cta(a, b, j, smem.data());   // grid (n_max, B)
  return 0;
}

extern "C" int emu_graph_all_paths(const int32_t* n_nodes, const int32_t* deg, const int32_t* nbr, const int32_t* lab, int32_t B,
                                   int32_t n_max, int32_t deg_max, int32_t max_len, int32_t K, int32_t self_id, int32_t tl_id,
                                   int32_t* all_paths, int32_t* pcount) {
  gtos::GraphAllPathsArgs a;
  a.n_nodes = n_nodes; a.deg = deg; a.nbr = nbr; a.lab = lab;
  a.B = B; a.n_max = n_max; a.deg_max = deg_max; a.max_len = max_len; a.K = K;
  a.self_id = self_id; a.tl_id = tl_id;
  a.all_paths = all_paths; a.pcount = pcount;
  std::vector<int32_t> smem((gtos::graph_paths_smem_bytes(n_max) + 3) / 4);
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < n_max; ++j) gtos::graph_all_paths_cta(a, b, j, smem.data());
  return 0;
}

extern "C" int emu_graph_bfs(const int32_t* n_nodes, const int32_t* deg, const int32_t* nbr, const int32_t* root, int32_t B,
                             int32_t n_max, int32_t deg_max, int32_t* order, int32_t* depth, int32_t* pos, int32_t* reached) {
  gtos::GraphBfsArgs a;
  a.n_nodes = n_nodes; a.deg = deg; a.nbr = nbr; a.root = root;
  a.B = B; a.n_max = n_max; a.deg_max = deg_max;
  a.order = order; a.depth = depth; a.pos = pos; a.reached = reached;
  for (int b = 0; b < B; ++b) gtos::graph_bfs_one(a, b);                      // one thread per graph
  return 0;
}
