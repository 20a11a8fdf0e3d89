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

struct AttnSmem {
  float* xs;  // [rows][dc+1]
  float* ys;  // [L][dc+1]
  float* sc;  // [rows][L+1]
};

__device__ __forceinline__ AttnSmem carve(float* base, int L, int dc, int AT_ROWS) {
  AttnSmem s;
  s.xs = base;
  s.ys = s.xs + AT_ROWS * (dc + 1);
  s.sc = s.ys + (size_t)L * (dc + 1);
  return s;
}

// rows [r0, r0+nr) x dims [c0, c0+dc) of a [len, B, ld] projection for (b, h) -> dst[nr][dc+1]
__device__ __forceinline__ void load_rows(float* dst, const float* src, long ld, int B, int b, int hoff, int r0, int nr,
                                          int len, int c0, int dc) {
  for (int idx = threadIdx.x; idx < nr * dc; idx += AT_THREADS) {
    int r = idx / dc, d = idx - r * dc;
    int t = r0 + r;
    dst[r * (dc + 1) + d] = (t < len) ? src[((long)t * B + b) * ld + hoff + c0 + d] : 0.f;
  }
}

// sc[r][j] += sum_d xs[r][d] * ys[j][d]     (lanes over j)
__device__ __forceinline__ void nt_accumulate(const AttnSmem& s, int L, int dc, int nrows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < nrows; r += AT_WARPS) {
    const float* x = s.xs + r * (dc + 1);
    for (int j = lane; j < L; j += 32) {
      const float* y = s.ys + (size_t)j * (dc + 1);
      float acc = 0.f;
#pragma unroll 8
      for (int d = 0; d < dc; ++d) acc = fmaf(x[d], y[d], acc);
      s.sc[r * (L + 1) + j] += acc;
    }
  }
}

// out[(row)*ld + d] = sum_j sc[r][j] * ys[j][d]   (lanes over d)
template <class F>
__device__ __forceinline__ void nn_product(const AttnSmem& s, int L, int dc, int nrows, F&& store) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < nrows; r += AT_WARPS) {
    const float* w = s.sc + r * (L + 1);
    for (int d = lane; d < dc; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < L; ++j) acc = fmaf(w[j], s.ys[(size_t)j * (dc + 1) + d], acc);
      store(r, d, acc);
    }
  }
}

__device__ __forceinline__ bool is_masked(const AttnArgs& a, int b, int t, int j) {
  if (a.key_pad && a.key_pad[(long)j * a.B + b]) return true;
  if (a.attn_mask && a.attn_mask[(long)t * a.S + j]) return true;
  return false;
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

  if (a.scores_jt) {
    for (int idx = threadIdx.x; idx < S * AT_ROWS; idx += AT_THREADS) {
      int j = idx / AT_ROWS, r = idx % AT_ROWS;
      s.sc[r * (S + 1) + j] = (r < nrows) ? a.scores_jt[((long)bh * S + j) * a.T + t0 + r] : 0.f;
    }
    __syncthreads();
  } else {
    for (int idx = threadIdx.x; idx < AT_ROWS * (S + 1); idx += AT_THREADS) s.sc[idx] = 0.f;
    for (int c0 = 0; c0 < a.hd; c0 += dc) {
      __syncthreads();
      load_rows(s.xs, a.q, a.ldq, a.B, b, hoff, t0, AT_ROWS, a.T, c0, dc);
      load_rows(s.ys, a.k, a.ldk, a.B, b, hoff, 0, S, S, c0, dc);
      __syncthreads();
      nt_accumulate(s, S, dc, nrows);
    }
    __syncthreads();
  }

  // ---- masked softmax (+ dropout) per row ----
  const unsigned long long seed = a.p_drop > 0.f ? reinterpret_cast<const unsigned long long*>(a.seed_ptr)[0] + a.seed_off : 0ull;
  const float ks = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const float sscale = a.scores_jt ? 1.f : a.scale;
  for (int r = warp; r < nrows; r += AT_WARPS) {
    const int t = t0 + r;
    float* w = s.sc + r * (S + 1);
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) {
      float v = is_masked(a, b, t, j) ? -INFINITY : w[j] * sscale;
      w[j] = v;
      mx = fmaxf(mx, v);
    }
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
    load_rows(s.ys, a.v, a.ldv, a.B, b, hoff, 0, S, S, c0, dc);
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
  if (rows < 8) rows = 8;
  if (rows > AT_ROWS_MAX) rows = 32;
  return rows;
}
static size_t attn_smem_bytes(int L, int hd, int rows) {
  int dc = hd < AT_DC ? hd : AT_DC;
  return sizeof(float) * ((size_t)rows * (dc + 1) + (size_t)L * (dc + 1) + (size_t)rows * (L + 1));
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

  // dPd[t][j] = dO[t] . V[j]
  for (int idx = threadIdx.x; idx < AT_ROWS * (S + 1); idx += AT_THREADS) s.sc[idx] = 0.f;
  for (int c0 = 0; c0 < a.hd; c0 += dc) {
    __syncthreads();
    load_rows(s.xs, g.dout, g.lddo, a.B, b, hoff, t0, AT_ROWS, a.T, c0, dc);
    load_rows(s.ys, a.v, a.ldv, a.B, b, hoff, 0, S, S, c0, dc);
    __syncthreads();
    nt_accumulate(s, S, dc, nrows);
  }
  __syncthreads();

  const unsigned long long seed = a.p_drop > 0.f ? reinterpret_cast<const unsigned long long*>(a.seed_ptr)[0] + a.seed_off : 0ull;
  const float ks = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  for (int r = warp; r < nrows; r += AT_WARPS) {
    const int t = t0 + r;
    float* w = s.sc + r * (S + 1);
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
      if (r < nrows) g.dscores_jt[((long)bh * S + j) * a.T + t0 + r] = s.sc[r * (S + 1) + j];
    }
  }
  if (g.dq) {
    for (int c0 = 0; c0 < a.hd; c0 += dc) {
      __syncthreads();
      load_rows(s.ys, a.k, a.ldk, a.B, b, hoff, 0, S, S, c0, dc);
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

  // sc[jr][t] = Pd[t][j0+jr]
  for (int idx = threadIdx.x; idx < T * AT_ROWS; idx += AT_THREADS) {
    int t = idx / AT_ROWS, jr = idx % AT_ROWS;
    float p = 0.f;
    if (jr < nrows) {
      long pi = ((long)bh * T + t) * S + j0 + jr;
      p = a.probs[pi];
      if (a.p_drop > 0.f) p = (rng_uniform(seed, (unsigned long long)pi) >= a.p_drop) ? p * ks : 0.f;
    }
    s.sc[jr * (T + 1) + t] = p;
  }
  for (int c0 = 0; c0 < a.hd; c0 += dc) {
    __syncthreads();
    load_rows(s.ys, g.dout, g.lddo, a.B, b, hoff, 0, T, T, c0, dc);
    __syncthreads();
    nn_product(s, T, dc, nrows, [&](int r, int d, float acc) {
      g.dv[((long)(j0 + r) * a.B + b) * g.lddv + hoff + c0 + d] = acc;
    });
  }
  if (g.dk) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * AT_ROWS; idx += AT_THREADS) {
      int t = idx / AT_ROWS, jr = idx % AT_ROWS;
      s.sc[jr * (T + 1) + t] = (jr < nrows) ? g.dscores_ts[((long)bh * T + t) * S + j0 + jr] * a.scale : 0.f;
    }
    for (int c0 = 0; c0 < a.hd; c0 += dc) {
      __syncthreads();
      load_rows(s.ys, a.q, a.ldq, a.B, b, hoff, 0, T, T, c0, dc);
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
