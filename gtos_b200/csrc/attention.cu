// gtos_b200 -- attention core kernels (softmax + masks + dropout + PV and their backward).
//
// Two users:
//   * encoder (generator/graph_transformer.py:136-159): the scores already contain the relation
//     terms and arrive from the fused tcgen05 kernel as [B,H,S(j),T(i)];
//   * decoder / vanilla MHA (generator/transformer.py:131-155): scores = scale * q k^T computed here.
// Sequences are short (<= ~260), so each CTA owns one (batch, head) and a block of 32 query (or
// key) rows, stages 64-wide feature chunks of the other operand in shared memory and keeps the
// [32 x S] score block in shared memory.  fp32 throughout; the contraction sizes here are <1% of
// the layer FLOPs.
#include "elementwise.cuh"

namespace gtos {

static constexpr int AT_THREADS = 256;
static constexpr int AT_WARPS = 8;
static constexpr int AT_ROWS_MAX = 64;  // query (or key) rows per CTA: all rows when the sequence is short
static constexpr int AT_DC = 64;    // feature chunk

// shared-memory working set of one CTA.  Row strides are multiples of 4 floats (16-byte vector loads):
//   xs [rows][dstr]  query-side chunk (q or dO)       dstr = round4(dc) + 4  (=68 for dc=64: conflict-free LDS.128)
//   ys [L4][dstr]    key-side chunk (K, V, dO or q)   L4 = round4(L), pad rows zero
//   sc [rows][lstr]  score block                      lstr = L4 + 4
struct AttnSmem {
  float* xs;
  float* ys;
  float* sc;
  int dstr, lstr, L4;
};

__host__ __device__ inline int r4(int n) { return (n + 3) & ~3; }

__device__ __forceinline__ AttnSmem carve(float* base, int L, int dc, int AT_ROWS) {
  AttnSmem s;
  s.dstr = r4(dc) + 4;
  s.L4 = r4(L);
  s.lstr = s.L4 + 4;
  s.xs = base;
  s.ys = s.xs + (size_t)AT_ROWS * s.dstr;
  s.sc = s.ys + (size_t)s.L4 * s.dstr;
  return s;
}

// rows [r0, r0+nr) x dims [c0, c0+dc) of a [len, B, ld] projection for (b, h) -> dst[nr][dstr]; rows >= len and the
// pad columns are zero-filled
__device__ __forceinline__ void load_rows(float* dst, int dstr, const float* src, long ld, int B, int b, int hoff,
                                          int r0, int nr, int len, int c0, int dc) {
  const int dc4 = r4(dc);
  const bool vec = ((dc & 3) == 0) && ((ld & 3) == 0) && (((hoff + c0) & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec) {
    // 16-byte loads, all of a thread's loads in flight before the first store (latency-bound otherwise)
    const int q = dc4 >> 2, total = nr * q;
#pragma unroll 4
    for (int idx = threadIdx.x; idx < total; idx += AT_THREADS) {
      const int r = idx / q, d4 = idx - r * q;
      const int t = r0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < len) v = *reinterpret_cast<const float4*>(src + ((long)t * B + b) * ld + hoff + c0 + 4 * d4);
      *reinterpret_cast<float4*>(dst + r * dstr + 4 * d4) = v;
    }
    return;
  }
  for (int idx = threadIdx.x; idx < nr * dc4; idx += AT_THREADS) {
    int r = idx / dc4, d = idx - r * dc4;
    int t = r0 + r;
    dst[r * dstr + d] = (t < len && d < dc) ? src[((long)t * B + b) * ld + hoff + c0 + d] : 0.f;
  }
}

constexpr int AT_RB = 8;  // rows per warp pass (register blocking)

// sc[r][j] += sum_d xs[r][d] * ys[j][d]     lanes over j, AT_RB rows per pass, 16-byte smem loads along d
__device__ __forceinline__ void nt_accumulate(const AttnSmem& s, int L, int dc, int nrows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dq = r4(dc) >> 2;
  for (int r0 = warp * AT_RB; r0 < nrows; r0 += AT_WARPS * AT_RB) {
    for (int j = lane; j < L; j += 32) {
      const float4* y = reinterpret_cast<const float4*>(s.ys + (size_t)j * s.dstr);
      float acc[AT_RB];
#pragma unroll
      for (int rr = 0; rr < AT_RB; ++rr) acc[rr] = 0.f;
      for (int d4 = 0; d4 < dq; ++d4) {
        const float4 yv = y[d4];
#pragma unroll
        for (int rr = 0; rr < AT_RB; ++rr) {
          const float4 xv = reinterpret_cast<const float4*>(s.xs + (size_t)(r0 + rr) * s.dstr)[d4];
          acc[rr] = fmaf(xv.x, yv.x, fmaf(xv.y, yv.y, fmaf(xv.z, yv.z, fmaf(xv.w, yv.w, acc[rr]))));
        }
      }
#pragma unroll
      for (int rr = 0; rr < AT_RB; ++rr)
        if (r0 + rr < nrows) s.sc[(size_t)(r0 + rr) * s.lstr + j] += acc[rr];
    }
  }
}

// out[r][d] = sum_j sc[r][j] * ys[j][d]   lanes over d (two per lane for dc = 64), AT_RB rows per pass
template <class F>
__device__ __forceinline__ void nn_product(const AttnSmem& s, int L, int dc, int nrows, F&& store) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d0 = lane, d1 = lane + 32;
  const bool has1 = d1 < dc, has0 = d0 < dc;
  for (int r0 = warp * AT_RB; r0 < nrows; r0 += AT_WARPS * AT_RB) {
    float a0[AT_RB], a1[AT_RB];
#pragma unroll
    for (int rr = 0; rr < AT_RB; ++rr) a0[rr] = a1[rr] = 0.f;
    for (int j4 = 0; j4 < s.L4; j4 += 4) {
      float y0[4], y1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* yr = s.ys + (size_t)(j4 + u) * s.dstr;
        y0[u] = has0 ? yr[d0] : 0.f;
        y1[u] = has1 ? yr[d1] : 0.f;
      }
#pragma unroll
      for (int rr = 0; rr < AT_RB; ++rr) {
        const float4 w = *reinterpret_cast<const float4*>(s.sc + (size_t)(r0 + rr) * s.lstr + j4);
        a0[rr] = fmaf(w.x, y0[0], fmaf(w.y, y0[1], fmaf(w.z, y0[2], fmaf(w.w, y0[3], a0[rr]))));
        a1[rr] = fmaf(w.x, y1[0], fmaf(w.y, y1[1], fmaf(w.z, y1[2], fmaf(w.w, y1[3], a1[rr]))));
      }
    }
#pragma unroll
    for (int rr = 0; rr < AT_RB; ++rr) {
      if (r0 + rr < nrows) {
        if (has0) store(r0 + rr, d0, a0[rr]);
        if (has1) store(r0 + rr, d1, a1[rr]);
      }
    }
  }
}

// zero the pad columns [L, lstr) of every score row and the pad rows [L, L4) of ys (so vector loops can run to L4)
__device__ __forceinline__ void zero_pads(const AttnSmem& s, int L, int rows) {
  const int padc = s.lstr - L;
  for (int idx = threadIdx.x; idx < rows * padc; idx += AT_THREADS) s.sc[(size_t)(idx / padc) * s.lstr + L + idx % padc] = 0.f;
  const int padr = s.L4 - L;
  for (int idx = threadIdx.x; idx < padr * s.dstr; idx += AT_THREADS) s.ys[(size_t)L * s.dstr + idx] = 0.f;
}

__device__ __forceinline__ bool is_masked(const AttnArgs& a, int b, int t, int j) {
  if (a.key_pad && a.key_pad[(long)j * a.B + b]) return true;
  if (a.attn_mask && a.attn_mask[(long)t * a.S + j]) return true;
  return false;
}

// -inf the masked entries of the score block in one cooperative, coalesced pass (instead of dependent byte loads
// inside every softmax row loop)
__device__ __forceinline__ void apply_masks(const AttnArgs& a, const AttnSmem& s, int b, int t0, int nrows, float sscale) {
  const int S = a.S;
  for (int idx = threadIdx.x; idx < nrows * S; idx += AT_THREADS) {
    const int r = idx / S, j = idx - r * S;
    float* p = s.sc + (size_t)r * s.lstr + j;
    *p = is_masked(a, b, t0 + r, j) ? -INFINITY : *p * sscale;
  }
}

__global__ void __launch_bounds__(AT_THREADS) attn_fwd_kernel(const AttnArgs a, const int rows_per_cta) {
  extern __shared__ float smem_f[];
  const int dc = a.hd < AT_DC ? a.hd : AT_DC;
  const int AT_ROWS = rows_per_cta;
  const AttnSmem s = carve(smem_f, a.S, dc, AT_ROWS);
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int t0 = blockIdx.y * AT_ROWS;
  const int nrows = (a.T - t0) < AT_ROWS ? (a.T - t0) : AT_ROWS;
  const int hoff = h * a.hd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.S;
  zero_pads(s, S, AT_ROWS);

  if (a.scores_jt) {
    for (int idx = threadIdx.x; idx < S * AT_ROWS; idx += AT_THREADS) {
      int j = idx / AT_ROWS, r = idx % AT_ROWS;
      s.sc[(size_t)r * s.lstr + j] = (r < nrows) ? a.scores_jt[((long)bh * S + j) * a.T + t0 + r] : 0.f;
    }
    __syncthreads();
  } else {
    for (int idx = threadIdx.x; idx < AT_ROWS * s.lstr; idx += AT_THREADS) s.sc[idx] = 0.f;
    for (int c0 = 0; c0 < a.hd; c0 += dc) {
      __syncthreads();
      load_rows(s.xs, s.dstr, a.q, a.ldq, a.B, b, hoff, t0, AT_ROWS, a.T, c0, dc);
      load_rows(s.ys, s.dstr, a.k, a.ldk, a.B, b, hoff, 0, S, S, c0, dc);
      __syncthreads();
      nt_accumulate(s, S, dc, nrows);
    }
    __syncthreads();
  }

  // ---- masked softmax (+ dropout) per row ----
  const unsigned long long seed = a.p_drop > 0.f ? reinterpret_cast<const unsigned long long*>(a.seed_ptr)[0] + a.seed_off : 0ull;
  const float ks = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const float sscale = a.scores_jt ? 1.f : a.scale;
  apply_masks(a, s, b, t0, nrows, sscale);
  __syncthreads();
  for (int r = warp; r < nrows; r += AT_WARPS) {
    const int t = t0 + r;
    float* w = s.sc + (size_t)r * s.lstr;
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, w[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      float e = (w[j] == -INFINITY) ? 0.f : __expf(w[j] - mx);
      w[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    const long prow = ((long)bh * a.T + t) * S;
    for (int j = lane; j < S; j += 32) {
      float p = w[j] * inv;
      a.probs[prow + j] = p;
      if (a.p_drop > 0.f) p = (rng_uniform(seed, (unsigned long long)(prow + j)) >= a.p_drop) ? p * ks : 0.f;
      if (a.probs_dropped) a.probs_dropped[prow + j] = p;
      w[j] = p;
    }
  }

  // ---- PV ----
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  for (int c0 = 0; c0 < a.hd; c0 += dc) {
    __syncthreads();
    load_rows(s.ys, s.dstr, a.v, a.ldv, a.B, b, hoff, 0, S, S, c0, dc);
    __syncthreads();
    nn_product(s, S, dc, nrows, [&](int r, int d, float acc) {
      long o = ((long)(t0 + r) * a.B + b) * a.ldo + hoff + c0 + d;
      a.out[o] = acc;
      if (ob) ob[o] = __float2bfloat16(acc);
    });
  }
}

// rows per CTA: the whole sequence when there are plenty of (batch, head) CTAs, otherwise split the rows so
// that the grid still covers ~2 CTAs per SM (e.g. the 1-head alignment attention, decoder.py:13)
static int attn_rows(int len, int bh) {
  int want_y = (2 * 148 + bh - 1) / bh;
  int rows = (len + want_y - 1) / want_y;
  rows = (rows + 7) & ~7;  // register blocking reads whole groups of AT_RB = 8 rows
  if (rows > AT_ROWS_MAX) rows = 32;
  return rows;
}
static size_t attn_smem_bytes(int L, int hd, int rows) {
  int dc = hd < AT_DC ? hd : AT_DC;
  int dstr = r4(dc) + 4, L4 = r4(L);
  return sizeof(float) * ((size_t)rows * dstr + (size_t)L4 * dstr + (size_t)rows * (L4 + 4));
}

int attn_fwd(const AttnArgs& a, cudaStream_t st) {
  GTOS_REQUIRE(a.hd >= 1 && (a.hd <= AT_DC || a.hd % AT_DC == 0), "attention: unsupported head_dim %d", a.hd);
  GTOS_REQUIRE(a.p_drop == 0.f || a.seed_ptr, "attention dropout needs a device seed pointer");
  if (a.T == 0 || a.B == 0) return GTOS_OK;
  const int rows = attn_rows(a.T, a.B * a.H);
  size_t smem = attn_smem_bytes(a.S, a.hd, rows);
  GTOS_REQUIRE(smem <= 227 * 1024, "attention: source length %d too long for shared memory", a.S);
  GTOS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(a.B * a.H, (a.T + rows - 1) / rows);
  attn_fwd_kernel<<<grid, AT_THREADS, smem, st>>>(a, rows);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// backward, query side: dS (and dq in decoder mode)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_q_kernel(const AttnBwdArgs g, const int rows_per_cta) {
  extern __shared__ float smem_f[];
  const AttnArgs& a = g.f;
  const int dc = a.hd < AT_DC ? a.hd : AT_DC;
  const int AT_ROWS = rows_per_cta;
  const AttnSmem s = carve(smem_f, a.S, dc, AT_ROWS);
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int t0 = blockIdx.y * AT_ROWS;
  const int nrows = (a.T - t0) < AT_ROWS ? (a.T - t0) : AT_ROWS;
  const int hoff = h * a.hd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.S;
  zero_pads(s, S, AT_ROWS);
  __syncthreads();

  // dPd[t][j] = dO[t] . V[j]
  for (int idx = threadIdx.x; idx < AT_ROWS * s.lstr; idx += AT_THREADS) s.sc[idx] = 0.f;
  for (int c0 = 0; c0 < a.hd; c0 += dc) {
    __syncthreads();
    load_rows(s.xs, s.dstr, g.dout, g.lddo, a.B, b, hoff, t0, AT_ROWS, a.T, c0, dc);
    load_rows(s.ys, s.dstr, a.v, a.ldv, a.B, b, hoff, 0, S, S, c0, dc);
    __syncthreads();
    nt_accumulate(s, S, dc, nrows);
  }
  __syncthreads();

  const unsigned long long seed = a.p_drop > 0.f ? reinterpret_cast<const unsigned long long*>(a.seed_ptr)[0] + a.seed_off : 0ull;
  const float ks = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  for (int r = warp; r < nrows; r += AT_WARPS) {
    const int t = t0 + r;
    float* w = s.sc + (size_t)r * s.lstr;
    const long prow = ((long)bh * a.T + t) * S;
    float dot = 0.f;
    for (int j = lane; j < S; j += 32) {
      float p = a.probs[prow + j];
      float dp = w[j];
      if (g.dprobs_extra) dp += g.dprobs_extra[prow + j];
      if (a.p_drop > 0.f) dp = (rng_uniform(seed, (unsigned long long)(prow + j)) >= a.p_drop) ? dp * ks : 0.f;
      w[j] = dp;
      dot += dp * p;
    }
    dot = warp_sum(dot);
    for (int j = lane; j < S; j += 32) {
      float p = a.probs[prow + j];
      float ds = p * (w[j] - dot);
      g.dscores_ts[prow + j] = ds;
      w[j] = ds;
    }
  }
  __syncthreads();
  if (g.dscores_jt) {
    // transposed store for the fused relation backward kernel: [B,H,S(j),T(i)], coalesced along i
    for (int idx = threadIdx.x; idx < S * AT_ROWS; idx += AT_THREADS) {
      int j = idx / AT_ROWS, r = idx % AT_ROWS;
      if (r < nrows) g.dscores_jt[((long)bh * S + j) * a.T + t0 + r] = s.sc[(size_t)r * s.lstr + j];
    }
  }
  if (g.dq) {
    for (int c0 = 0; c0 < a.hd; c0 += dc) {
      __syncthreads();
      load_rows(s.ys, s.dstr, a.k, a.ldk, a.B, b, hoff, 0, S, S, c0, dc);
      __syncthreads();
      nn_product(s, S, dc, nrows, [&](int r, int d, float acc) {
        g.dq[((long)(t0 + r) * a.B + b) * g.lddq + hoff + c0 + d] = acc * a.scale;
      });
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward, key side: dV = Pd^T dO ; dK = scale * dS^T q      (CTA = 32 key rows of one (b,h))
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_kv_kernel(const AttnBwdArgs g, const int rows_per_cta) {
  extern __shared__ float smem_f[];
  const AttnArgs& a = g.f;
  const int dc = a.hd < AT_DC ? a.hd : AT_DC;
  const int T = a.T, S = a.S;
  const int AT_ROWS = rows_per_cta;
  const AttnSmem s = carve(smem_f, T, dc, AT_ROWS);
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int j0 = blockIdx.y * AT_ROWS;
  const int nrows = (S - j0) < AT_ROWS ? (S - j0) : AT_ROWS;
  const int hoff = h * a.hd;
  const unsigned long long seed = a.p_drop > 0.f ? reinterpret_cast<const unsigned long long*>(a.seed_ptr)[0] + a.seed_off : 0ull;
  const float ks = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;

  zero_pads(s, T, AT_ROWS);
  // sc[jr][t] = Pd[t][j0+jr]
  for (int idx = threadIdx.x; idx < T * AT_ROWS; idx += AT_THREADS) {
    int t = idx / AT_ROWS, jr = idx % AT_ROWS;
    float p = 0.f;
    if (jr < nrows) {
      long pi = ((long)bh * T + t) * S + j0 + jr;
      p = a.probs[pi];
      if (a.p_drop > 0.f) p = (rng_uniform(seed, (unsigned long long)pi) >= a.p_drop) ? p * ks : 0.f;
    }
    s.sc[(size_t)jr * s.lstr + t] = p;
  }
  for (int c0 = 0; c0 < a.hd; c0 += dc) {
    __syncthreads();
    load_rows(s.ys, s.dstr, g.dout, g.lddo, a.B, b, hoff, 0, T, T, c0, dc);
    __syncthreads();
    nn_product(s, T, dc, nrows, [&](int r, int d, float acc) {
      g.dv[((long)(j0 + r) * a.B + b) * g.lddv + hoff + c0 + d] = acc;
    });
  }
  if (g.dk) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * AT_ROWS; idx += AT_THREADS) {
      int t = idx / AT_ROWS, jr = idx % AT_ROWS;
      s.sc[(size_t)jr * s.lstr + t] = (jr < nrows) ? g.dscores_ts[((long)bh * T + t) * S + j0 + jr] * a.scale : 0.f;
    }
    for (int c0 = 0; c0 < a.hd; c0 += dc) {
      __syncthreads();
      load_rows(s.ys, s.dstr, a.q, a.ldq, a.B, b, hoff, 0, T, T, c0, dc);
      __syncthreads();
      nn_product(s, T, dc, nrows, [&](int r, int d, float acc) {
        g.dk[((long)(j0 + r) * a.B + b) * g.lddk + hoff + c0 + d] = acc;
      });
    }
  }
}

int attn_bwd(const AttnBwdArgs& g, cudaStream_t st) {
  const AttnArgs& a = g.f;
  GTOS_REQUIRE(a.hd >= 1 && (a.hd <= AT_DC || a.hd % AT_DC == 0), "attention: unsupported head_dim %d", a.hd);
  if (a.T == 0 || a.B == 0) return GTOS_OK;
  const int rq = attn_rows(a.T, a.B * a.H), rk = attn_rows(a.S, a.B * a.H);
  size_t smq = attn_smem_bytes(a.S, a.hd, rq), smk = attn_smem_bytes(a.T, a.hd, rk);
  GTOS_REQUIRE(smq <= 227 * 1024 && smk <= 227 * 1024, "attention: sequence too long for shared memory");
  GTOS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smq));
  GTOS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smk));
  dim3 gq(a.B * a.H, (a.T + rq - 1) / rq);
  attn_bwd_q_kernel<<<gq, AT_THREADS, smq, st>>>(g, rq);
  GTOS_LAUNCH_CHECK();
  dim3 gk(a.B * a.H, (a.S + rk - 1) / rk);
  attn_bwd_kv_kernel<<<gk, AT_THREADS, smk, st>>>(g, rk);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

}  // namespace gtos
