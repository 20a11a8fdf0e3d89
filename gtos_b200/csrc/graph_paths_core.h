// gtos_b200 -- SURVEY.md §8 f-3: all-pairs shortest label paths of a batch of small graphs, one path per ordered pair
// sampled uniformly among the equally short ones.
//
// Reference: AMRGraph.collect_concepts_and_relations (generator/AMRGraph.py:100-115) enumerates, with networkx, every
// shortest NODE path src -> tgt of the bidirected labelled graph (each edge has a `_reverse_` twin, AMRGraph.py:76-80;
// `_r_` in translator/dependencyGraph.py:30-34) and keeps the edge labels; batchify (generator/data.py:148-154) then
// draws one of them with random.choice, replaces the empty path by <SELF> and a path of more than `max_len` labels by
// <TL>.  Enumeration is exponential in the worst case; the same distribution comes from counting:
//   sigma[v] = number of shortest paths v -> tgt  (BFS from tgt; the structure is symmetric, so dist(v -> tgt) is the BFS
//              depth of v), and a walk from src that moves from v to a neighbour u one level closer with probability
//   sigma[u] / sum of sigma over such neighbours
// visits every shortest path src -> tgt with probability 1 / sigma[src].
//
// One CTA owns one (graph, target) and produces the paths of ALL sources to that target.  The code below is written in
// barrier-separated phases (GTOS_PHASE ... GTOS_PHASE_END) so that the SAME source also compiles as plain C++, where a
// phase is a loop over the emulated thread ids (tests/emu/graph_paths_emu.cpp: the CPU check of this very code against
// the oracle; the CUDA build is what ships).  Phases are pull-based - a thread writes only its own nodes / sources - so
// there is no intra-phase ordering to get wrong, and the floating-point sums run in adjacency order: results are
// bit-identical on both sides and reproducible from (seed, pair).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GTOS_HD __host__ __device__ __forceinline__
#else
#define GTOS_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define GTOS_PHASE(tid, nthr) { const int tid = (int)threadIdx.x; const int nthr = (int)blockDim.x;
#define GTOS_PHASE_END } __syncthreads();
#else
#ifndef GTOS_EMU_THREADS
#define GTOS_EMU_THREADS 128
#endif
#define GTOS_PHASE(tid, nthr) for (int tid = 0, nthr = GTOS_EMU_THREADS; tid < nthr; ++tid) {
#define GTOS_PHASE_END }
#endif

namespace gtos {

// the path sampler's counter-based uniform (64-bit splitmix finalizer; the dropout generator of common.cuh is a separate,
// cheaper function) - pinned bit for bit to oracle/paths_oracle.py, usable from host code too
GTOS_HD float paths_uniform(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<float>(static_cast<uint32_t>(z >> 40)) * (1.0f / 16777216.0f);
}

struct GraphPathsArgs {
  const int32_t* n_nodes;   // [B]
  const int32_t* deg;       // [B, n_max]            out-degree of every node (<= deg_max)
  const int32_t* nbr;       // [B, n_max, deg_max]   neighbour of edge k of node v
  const int32_t* lab;       // [B, n_max, deg_max]   label id of the edge v -> nbr
  int32_t B, n_max, deg_max, max_len;
  int32_t self_id, tl_id;
  uint64_t seed;            // already combined: *seed_ptr + seed_off
  int32_t* paths;           // [B, n_max, n_max, max_len]  labels of the path i -> j in walking order, 0 padded
  int32_t* plen;            // [B, n_max, n_max]           number of labels (1 for <SELF> / <TL>); 0 = pair outside the graph
};

// shared-memory working set of one CTA: dist [n_max] int32, sigma [n_max] float, mark [n_max] int32, flag [4] int32
GTOS_HD size_t graph_paths_smem_bytes(int n_max) { return (size_t)n_max * 12 + 16; }

// everything one CTA does for (graph b, target j); `smem` = graph_paths_smem_bytes(n_max) bytes, 4-byte aligned
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void graph_paths_cta(const GraphPathsArgs& a, int b, int j, void* smem) {
  int32_t* dist = reinterpret_cast<int32_t*>(smem);
  float* sigma = reinterpret_cast<float*>(dist + a.n_max);
  int32_t* mark = reinterpret_cast<int32_t*>(sigma + a.n_max);
  int32_t* flag = mark + a.n_max;                                   // flag[0]: the next level is not empty
  const int n = a.n_nodes[b];
  const int32_t* deg = a.deg + (long)b * a.n_max;
  const int32_t* nbr = a.nbr + (long)b * a.n_max * a.deg_max;
  const int32_t* lab = a.lab + (long)b * a.n_max * a.deg_max;
  int32_t* paths = a.paths + ((long)b * a.n_max * a.n_max) * a.max_len;
  int32_t* plen = a.plen + (long)b * a.n_max * a.n_max;

  if (j >= n) {                                                      // target outside the graph: the whole column is padding
    GTOS_PHASE(tid, nthr)
      for (int i = tid; i < a.n_max; i += nthr) {
        plen[(long)i * a.n_max + j] = 0;
        for (int s = 0; s < a.max_len; ++s) paths[((long)i * a.n_max + j) * a.max_len + s] = 0;
      }
    GTOS_PHASE_END
    return;
  }

  GTOS_PHASE(tid, nthr)
    for (int v = tid; v < a.n_max; v += nthr) {
      dist[v] = (v == j) ? 0 : -1;
      sigma[v] = (v == j) ? 1.f : 0.f;
      mark[v] = 0;
    }
    if (tid == 0) {
      flag[0] = 1;
      flag[1] = 0;                                                   // a node joined the next level
      flag[2] = 0;                                                   // a count of the next level left the safe range
    }
  GTOS_PHASE_END

  // ---- level-synchronous BFS from the target with shortest-path counts ----
  for (int level = 0; flag[0] != 0; ++level) {                       // flags are only written between barriers
    GTOS_PHASE(tid, nthr)
      // pull: an unvisited node joins level + 1 if one of its neighbours sits on `level` (symmetric structure)
      for (int v = tid; v < n; v += nthr) {
        if (dist[v] >= 0) continue;
        float s = 0.f;
        bool hit = false;
        for (int k = 0; k < deg[v]; ++k) {
          const int u = nbr[(long)v * a.deg_max + k];
          if (dist[u] == level) {
            s += sigma[u];
            hit = true;
          }
        }
        if (hit) {
          mark[v] = 1;
          sigma[v] = s;
          if (s > 1.0e30f) flag[2] = 1;                              // benign same-value race
        }
      }
    GTOS_PHASE_END
    GTOS_PHASE(tid, nthr)
      for (int v = tid; v < n; v += nthr)
        if (mark[v]) {
          mark[v] = 0;
          dist[v] = level + 1;
          flag[1] = 1;                                               // benign same-value race
        }
    GTOS_PHASE_END
    // keep the counts of the new level in float range: only ratios inside one level are ever used
    GTOS_PHASE(tid, nthr)
      if (flag[2] != 0)
        for (int v = tid; v < n; v += nthr)
          if (dist[v] == level + 1) sigma[v] = sigma[v] * (1.0f / 1.0e30f);
    GTOS_PHASE_END
    GTOS_PHASE(tid, nthr)
      if (tid == 0) {
        flag[0] = flag[1];
        flag[1] = 0;
        flag[2] = 0;
      }
    GTOS_PHASE_END
  }

  // ---- one uniformly drawn shortest path i -> j for every source i ----
  GTOS_PHASE(tid, nthr)
    for (int i = tid; i < a.n_max; i += nthr) {
      int32_t* out = paths + ((long)i * a.n_max + j) * a.max_len;
      for (int s = 0; s < a.max_len; ++s) out[s] = 0;
      if (i >= n) {
        plen[(long)i * a.n_max + j] = 0;
        continue;
      }
      const int d = dist[i];
      if (d == 0) {                                                  // data.py:151-152  [] -> <SELF>
        out[0] = a.self_id;
        plen[(long)i * a.n_max + j] = 1;
      } else if (d < 0 || d > a.max_len) {                           // data.py:153-154  too long (or unreachable) -> <TL>
        out[0] = a.tl_id;
        plen[(long)i * a.n_max + j] = 1;
      } else {
        int v = i;
        for (int s = 0; s < d; ++s) {
          const int want = d - s - 1;
          float total = 0.f;
          for (int k = 0; k < deg[v]; ++k) {
            const int u = nbr[(long)v * a.deg_max + k];
            if (dist[u] == want) total += sigma[u];
          }
          const uint64_t e = (((uint64_t)b * a.n_max + i) * a.n_max + j) * a.max_len + s;
          const float r = paths_uniform(a.seed, e) * total;
          float cum = 0.f;
          int pick = -1, last = -1;
          for (int k = 0; k < deg[v]; ++k) {
            const int u = nbr[(long)v * a.deg_max + k];
            if (dist[u] != want) continue;
            last = k;
            cum += sigma[u];
            if (cum > r) {
              pick = k;
              break;
            }
          }
          if (pick < 0) pick = last;                                 // r == total after rounding
          out[s] = lab[(long)v * a.deg_max + pick];
          v = nbr[(long)v * a.deg_max + pick];
        }
        plen[(long)i * a.n_max + j] = d;
      }
    }
  GTOS_PHASE_END
}

// -------------------------------------------------------------------------------------------------------------------
// Evaluation batches keep EVERY shortest path of a pair (generator/data.py:176-225; the model averages their encodings,
// generator/generator.py:83-88).  Same BFS, exact saturating counts, then one thread per source enumerates its paths
// depth first in adjacency order - the order of oracle/paths_oracle.py::all_shortest_label_paths - up to K per pair.
//   all_paths[b][i][j][k][0..len)  labels of the k-th path, 0 padded          (k < min(pcount, K))
//   pcount[b][i][j]                number of shortest paths, saturated at K + 1  (0 = pair outside the graph);
//                                  <SELF> / <TL> pairs hold one entry (data.py:197-199 keeps all_path[:1])
// A pair with more than K paths reports K + 1 and carries its first K: the caller decides (gtos_b200/paths.py raises).
// -------------------------------------------------------------------------------------------------------------------
struct GraphAllPathsArgs {
  const int32_t* n_nodes;
  const int32_t* deg;
  const int32_t* nbr;
  const int32_t* lab;
  int32_t B, n_max, deg_max, max_len, K;
  int32_t self_id, tl_id;
  int32_t* all_paths;       // [B, n_max, n_max, K, max_len]
  int32_t* pcount;          // [B, n_max, n_max]
};

static const int GTOS_PATHS_MAX_LEN = 16;

// shared-memory working set: dist [n_max] int32, count [n_max] uint32, mark [n_max] int32, flag [4] int32
#if defined(__CUDACC__)
__host__ __device__
#endif
inline void graph_all_paths_cta(const GraphAllPathsArgs& a, int b, int j, void* smem) {
  int32_t* dist = reinterpret_cast<int32_t*>(smem);
  uint32_t* count = reinterpret_cast<uint32_t*>(dist + a.n_max);
  int32_t* mark = reinterpret_cast<int32_t*>(count + a.n_max);
  int32_t* flag = mark + a.n_max;
  const int n = a.n_nodes[b];
  const int32_t* deg = a.deg + (long)b * a.n_max;
  const int32_t* nbr = a.nbr + (long)b * a.n_max * a.deg_max;
  const int32_t* lab = a.lab + (long)b * a.n_max * a.deg_max;
  const long pair_stride = (long)a.K * a.max_len;
  int32_t* all_paths = a.all_paths + ((long)b * a.n_max * a.n_max) * pair_stride;
  int32_t* pcount = a.pcount + (long)b * a.n_max * a.n_max;
  const uint32_t cap = (uint32_t)a.K + 1u;

  if (j >= n) {
    GTOS_PHASE(tid, nthr)
      for (int i = tid; i < a.n_max; i += nthr) {
        pcount[(long)i * a.n_max + j] = 0;
        int32_t* out = all_paths + ((long)i * a.n_max + j) * pair_stride;
        for (long s = 0; s < pair_stride; ++s) out[s] = 0;
      }
    GTOS_PHASE_END
    return;
  }

  GTOS_PHASE(tid, nthr)
    for (int v = tid; v < a.n_max; v += nthr) {
      dist[v] = (v == j) ? 0 : -1;
      count[v] = (v == j) ? 1u : 0u;
      mark[v] = 0;
    }
    if (tid == 0) flag[0] = 1;
  GTOS_PHASE_END

  for (int level = 0; flag[0] != 0; ++level) {
    GTOS_PHASE(tid, nthr)
      if (tid == 0) flag[1] = 0;
      for (int v = tid; v < n; v += nthr) {
        if (dist[v] >= 0) continue;
        uint32_t c = 0;
        bool hit = false;
        for (int k = 0; k < deg[v]; ++k) {
          const int u = nbr[(long)v * a.deg_max + k];
          if (dist[u] == level) {
            c += count[u];                                           // both <= cap: no wrap before the clamp
            if (c > cap) c = cap;
            hit = true;
          }
        }
        if (hit) {
          mark[v] = 1;
          count[v] = c;
        }
      }
    GTOS_PHASE_END
    GTOS_PHASE(tid, nthr)
      for (int v = tid; v < n; v += nthr)
        if (mark[v]) {
          mark[v] = 0;
          dist[v] = level + 1;
          flag[1] = 1;
        }
    GTOS_PHASE_END
    GTOS_PHASE(tid, nthr)
      if (tid == 0) flag[0] = flag[1];
    GTOS_PHASE_END
  }

  GTOS_PHASE(tid, nthr)
    for (int i = tid; i < a.n_max; i += nthr) {
      int32_t* out = all_paths + ((long)i * a.n_max + j) * pair_stride;
      for (long s = 0; s < pair_stride; ++s) out[s] = 0;
      if (i >= n) {
        pcount[(long)i * a.n_max + j] = 0;
        continue;
      }
      const int d = dist[i];
      if (d == 0) {
        out[0] = a.self_id;
        pcount[(long)i * a.n_max + j] = 1;
        continue;
      }
      if (d < 0 || d > a.max_len) {
        out[0] = a.tl_id;
        pcount[(long)i * a.n_max + j] = 1;
        continue;
      }
      // depth-first enumeration in adjacency order; the stack is at most max_len deep
      int node[GTOS_PATHS_MAX_LEN + 1], next_k[GTOS_PATHS_MAX_LEN + 1], labs[GTOS_PATHS_MAX_LEN];
      int depth = 0, found = 0;
      node[0] = i;
      next_k[0] = 0;
      while (depth >= 0 && found < a.K) {
        const int v = node[depth];
        if (depth == d) {                                            // v == j
          for (int s = 0; s < d; ++s) out[(long)found * a.max_len + s] = labs[s];
          ++found;
          --depth;
          continue;
        }
        int k = next_k[depth];
        const int want = d - depth - 1;
        while (k < deg[v] && dist[nbr[(long)v * a.deg_max + k]] != want) ++k;
        if (k >= deg[v]) {
          --depth;
          continue;
        }
        next_k[depth] = k + 1;
        labs[depth] = lab[(long)v * a.deg_max + k];
        node[depth + 1] = nbr[(long)v * a.deg_max + k];
        next_k[depth + 1] = 0;
        ++depth;
      }
      pcount[(long)i * a.n_max + j] = (int32_t)count[i];
    }
  GTOS_PHASE_END
}

// -------------------------------------------------------------------------------------------------------------------
// Node order of a batch: AMRGraph.bfs (generator/AMRGraph.py:82-98; translator/dependencyGraph.py:36-52) - a queue BFS
// from the root that visits neighbours in adjacency (insertion) order; the batch uses the queue order as node order and
// the BFS depths as `concept_depth` (data.py:129).  The queue discipline is sequential by definition, the graphs are
// tiny: one THREAD per graph replays it exactly.
//   order[b][k] = k-th node of the queue, depth[b][k] = its depth, pos[b][v] = position of node v (-1: not reached),
//   reached[b]  = queue length (== n_nodes[b] iff the graph is connected, the reference's `is_connected`)
// -------------------------------------------------------------------------------------------------------------------
struct GraphBfsArgs {
  const int32_t* n_nodes;   // [B]
  const int32_t* deg;       // [B, n_max]
  const int32_t* nbr;       // [B, n_max, deg_max]
  const int32_t* root;      // [B]
  int32_t B, n_max, deg_max;
  int32_t* order;           // [B, n_max]  (-1 beyond the queue)
  int32_t* depth;           // [B, n_max]  (0 beyond the queue)
  int32_t* pos;             // [B, n_max]
  int32_t* reached;         // [B]
};

#if defined(__CUDACC__)
__host__ __device__
#endif
inline void graph_bfs_one(const GraphBfsArgs& a, int b) {
  const int n = a.n_nodes[b];
  const int32_t* deg = a.deg + (long)b * a.n_max;
  const int32_t* nbr = a.nbr + (long)b * a.n_max * a.deg_max;
  int32_t* order = a.order + (long)b * a.n_max;
  int32_t* depth = a.depth + (long)b * a.n_max;
  int32_t* pos = a.pos + (long)b * a.n_max;
  for (int v = 0; v < a.n_max; ++v) {
    order[v] = -1;
    depth[v] = 0;
    pos[v] = -1;
  }
  int tail = 0;
  const int r = a.root[b];
  if (n > 0 && r >= 0 && r < n) {
    order[0] = r;
    pos[r] = 0;
    tail = 1;
  }
  for (int head = 0; head < tail; ++head) {                          // AMRGraph.py:88-96
    const int u = order[head];
    const int du = depth[head];
    for (int k = 0; k < deg[u]; ++k) {
      const int v = nbr[(long)u * a.deg_max + k];
      if (pos[v] < 0) {
        pos[v] = tail;
        order[tail] = v;
        depth[tail] = du + 1;
        ++tail;
      }
    }
  }
  a.reached[b] = tail;
}

}  // namespace gtos
