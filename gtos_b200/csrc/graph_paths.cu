// gtos_b200 -- SURVEY.md §8 f-3: shortest label paths of a batch of graphs on the GPU (graph_paths_core.h holds the
// algorithm; this file is the launch).  Integer / byte work on tiny graphs (n <= ~260 nodes): one CTA per
// (graph, target) keeps its BFS state in shared memory and reads the padded adjacency of its graph (L1/L2 resident:
// n_max * deg_max * 8 B per graph) - latency-bound, sized so that a batch is B * n_max CTAs (config 2: 2 560, config 4:
// 65 536), i.e. many waves of 148 x 16 resident CTAs.
#include "elementwise.cuh"
#include "graph_paths_core.h"

namespace gtos {

__global__ void __launch_bounds__(128) graph_paths_kernel(GraphPathsArgs a, const unsigned long long* seed_ptr,
                                                          unsigned long long seed_off) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ __align__(16) unsigned char gp_smem[];
  a.seed = (seed_ptr ? seed_ptr[0] : 0ull) + seed_off;
  graph_paths_cta(a, (int)blockIdx.y, (int)blockIdx.x, gp_smem);
}

int graph_paths(const GraphPathsArgs& a, const void* seed_ptr, unsigned long long seed_off, cudaStream_t st) {
  GTOS_REQUIRE(a.B >= 0 && a.n_max >= 1 && a.deg_max >= 1 && a.max_len >= 1 && a.max_len <= 16,
               "graph_paths: bad sizes (B=%d n_max=%d deg_max=%d max_len=%d)", a.B, a.n_max, a.deg_max, a.max_len);
  GTOS_REQUIRE(a.n_nodes && a.deg && a.nbr && a.lab && a.paths && a.plen, "graph_paths: null argument");
  GTOS_REQUIRE(a.B <= 65535, "graph_paths: at most 65535 graphs per call");
  if (a.B == 0) return GTOS_OK;
  const size_t smem = graph_paths_smem_bytes(a.n_max);
  GTOS_REQUIRE(smem <= 48 * 1024, "graph_paths: %d nodes per graph do not fit the shared-memory working set", a.n_max);
  GTOS_KLAUNCH(graph_paths_kernel, dim3((unsigned)a.n_max, (unsigned)a.B), dim3(128), smem, st, a,
               reinterpret_cast<const unsigned long long*>(seed_ptr), seed_off);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

__global__ void __launch_bounds__(128) graph_all_paths_kernel(const GraphAllPathsArgs a) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ __align__(16) unsigned char gp_smem[];
  graph_all_paths_cta(a, (int)blockIdx.y, (int)blockIdx.x, gp_smem);
}

int graph_all_paths(const GraphAllPathsArgs& a, cudaStream_t st) {
  GTOS_REQUIRE(a.B >= 0 && a.n_max >= 1 && a.deg_max >= 1 && a.max_len >= 1 && a.max_len <= GTOS_PATHS_MAX_LEN && a.K >= 1,
               "graph_all_paths: bad sizes (B=%d n_max=%d deg_max=%d max_len=%d K=%d)", a.B, a.n_max, a.deg_max, a.max_len,
               a.K);
  GTOS_REQUIRE(a.n_nodes && a.deg && a.nbr && a.lab && a.all_paths && a.pcount, "graph_all_paths: null argument");
  GTOS_REQUIRE(a.B <= 65535, "graph_all_paths: at most 65535 graphs per call");
  if (a.B == 0) return GTOS_OK;
  const size_t smem = graph_paths_smem_bytes(a.n_max);
  GTOS_REQUIRE(smem <= 48 * 1024, "graph_all_paths: %d nodes per graph do not fit the shared-memory working set", a.n_max);
  GTOS_KLAUNCH(graph_all_paths_kernel, dim3((unsigned)a.n_max, (unsigned)a.B), dim3(128), smem, st, a);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

__global__ void __launch_bounds__(128) graph_bfs_kernel(const GraphBfsArgs a) {
  GTOS_PDL_PROLOGUE();
  const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (b < a.B) graph_bfs_one(a, b);
}

int graph_bfs(const GraphBfsArgs& a, cudaStream_t st) {
  GTOS_REQUIRE(a.B >= 0 && a.n_max >= 1 && a.deg_max >= 1, "graph_bfs: bad sizes (B=%d n_max=%d deg_max=%d)", a.B, a.n_max,
               a.deg_max);
  GTOS_REQUIRE(a.n_nodes && a.deg && a.nbr && a.root && a.order && a.depth && a.pos && a.reached, "graph_bfs: null argument");
  if (a.B == 0) return GTOS_OK;
  GTOS_KLAUNCH(graph_bfs_kernel, dim3((unsigned)((a.B + 127) / 128)), dim3(128), 0, st, a);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

}  // namespace gtos
