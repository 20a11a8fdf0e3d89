// gtos_b200 -- attention core kernels (masks + softmax + dropout + PV and their backward).
//
// Two users:
//   * encoder (generator/graph_transformer.py:136-159): the scores already contain the relation
//     terms and arrive from the fused tcgen05 kernel as [B,H,S(j),T(i)];
//   * decoder / vanilla MHA (generator/transformer.py:131-155): scores = scale * q k^T computed here.
// Sequences are short (<= ~260 keys), so one CTA owns one (batch, head) and a block of <= 64 query (or key)
// rows.  The three small products of each kernel (QK^T / dO V^T, PV / dS K, P^T dO / dS^T q) run on the
// tensor cores as 16x16x16 bf16 tiles with fp32 accumulation (warp-level mma; these are 60x60x64-sized
// problems, far below a 128-row tcgen05 tile), operands converted to bf16 while they are staged in shared
// memory in 64-wide feature chunks.  Softmax, masks, dropout and the dS algebra stay in fp32 on the score
// block in shared memory.  This is < 1 % of the layer FLOPs: the kernels are
// latency-bound, so the design goal is few instructions and few dependent global round trips.
#include <mma.h>

#include "elementwise.cuh"

namespace gtos {

using namespace nvcuda;

// GTOS_DBG=2: clock64 timestamps of the phases of CTA 0 (timing experiments; attn_debug_read_trace)
__device__ unsigned long long g_attn_trace[3 * 16];
__device__ int g_attn_trace_on = 0;
#define ATT_TRACE(k, slot) do { if (g_attn_trace_on && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_attn_trace[(k) * 16 + (slot)] = clock64(); } while (0)

static constexpr int AT_THREADS = 256;
static constexpr int AT_WARPS = 8;
// 64 registers per thread: four CTAs per SM, so the 512 (batch, head) CTAs of this path are resident in ONE wave.  The
// compiler otherwise takes 128 registers (ncu: occupancy limit 2 CTAs per SM, 23 % of the warp slots active, two waves).
static constexpr int AT_MIN_CTAS = 4;
static constexpr int AT_ROWS_MAX = 64;  // query (or key) rows per CTA
static constexpr int AT_DC = 64;        // feature chunk

__host__ __device__ inline int r16(int n) { return (n + 15) & ~15; }

// shared-memory working set of one CTA (all tile extents rounded up to 16, pads zero-filled):
//   sc [R16][lstr]  fp32  score block (also the fp32 staging tile of the second product)
//   xb [R16][dstr]  bf16  query-side chunk (q or dO)
//   yb [L16][dstr]  bf16  key-side chunk (K, V, dO or q)
//   pb [R16][lstr]  bf16  probabilities / dS as MMA operand
//
// fp32 mode (AttnArgs::precise, north star: 1e-3 against the fp32 reference): every MMA operand x is staged as a PAIR of
// bf16 planes, x_hi = bf16(x) and x_lo = bf16(x - x_hi) (16 significand bits together), and every product runs as three
// tensor-core passes into the same fp32 accumulator: A_hi B_hi + A_lo B_hi + A_hi B_lo (the dropped A_lo B_lo term is
// ~2^-18 relative).  xl / yl / pl are the lo planes (null in bf16 mode); softmax uses expf instead of __expf.
struct AttnSmem {
  __nv_bfloat16* xb;
  __nv_bfloat16* yb;
  __nv_bfloat16* pb;
  __nv_bfloat16* xl;
  __nv_bfloat16* yl;
  __nv_bfloat16* pl;
  float* sc;
  int dstr, lstr, L16, R16;
};

__host__ __device__ inline size_t attn_smem_layout(int L, int dc, int rows, int* dstr, int* lstr, int* L16, int* R16,
                                                   bool split = false) {
  *dstr = r16(dc) + 8;
  *L16 = r16(L);
  *R16 = r16(rows);
  int ls = *L16 > r16(dc) ? *L16 : r16(dc);   // sc doubles as the [R16 x dc16] output staging tile
  *lstr = ls + 8;
  const size_t planes = (size_t)(*R16) * (*dstr) * 2 + (size_t)(*L16) * (*dstr) * 2 + (size_t)(*R16) * (*lstr) * 2;
  return planes * (split ? 2 : 1) + (size_t)(*R16) * (*lstr) * 4;
}

template <bool SPLIT>
__device__ __forceinline__ AttnSmem carve(uint8_t* base, int L, int dc, int rows) {
  AttnSmem s;
  attn_smem_layout(L, dc, rows, &s.dstr, &s.lstr, &s.L16, &s.R16, SPLIT);
  s.sc = reinterpret_cast<float*>(base);
  s.xb = reinterpret_cast<__nv_bfloat16*>(s.sc + (size_t)s.R16 * s.lstr);
  s.yb = s.xb + (size_t)s.R16 * s.dstr;
  s.pb = s.yb + (size_t)s.L16 * s.dstr;
  s.xl = s.yl = s.pl = nullptr;
  if (SPLIT) {
    s.xl = s.pb + (size_t)s.R16 * s.lstr;
    s.yl = s.xl + (size_t)s.R16 * s.dstr;
    s.pl = s.yl + (size_t)s.L16 * s.dstr;
  }
  return s;
}

// one probability / dS value as MMA operand: hi plane (and lo plane in fp32 mode)
template <bool SPLIT>
__device__ __forceinline__ void put_operand(__nv_bfloat16* hi, __nv_bfloat16* lo, int j, float v) {
  const __nv_bfloat16 h = __float2bfloat16(v);
  hi[j] = h;
  if (SPLIT) lo[j] = __float2bfloat16(v - __bfloat162float(h));
}

// These kernels are instruction-bound (ncu: ~14 M warp instructions per launch, 40 % of them integer division of
// a flat thread index).  All 2-D loops therefore split the thread index with shifts: `qs` = log2 of the number
// of lanes per row rounded up to a power of two (lanes beyond the row width idle).
__device__ __forceinline__ int log2_ceil(int n) { return n <= 1 ? 0 : 32 - __clz(n - 1); }

// rows [r0, r0+nr16) x dims [c0, c0+dc) of a [len, B, ld] fp32 projection for (b, h) -> bf16 dst[nr16][dstr];
// rows >= len and columns >= dc (up to r16(dc)) are zero-filled.  16-byte global loads when alignment allows.
__device__ __forceinline__ void load_rows_bf16(__nv_bfloat16* dst, int dstr, const float* src, long ld, int B, int b,
                                               int hoff, int r0, int nr16, int len, int c0, int dc,
                                               __nv_bfloat16* dst_lo = nullptr) {
  const int dc16 = r16(dc);
  const bool vec = ((dc & 3) == 0) && ((ld & 3) == 0) && (((hoff + c0) & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const int q = dc16 >> 2;                       // float4 groups per row
  const int qs = log2_ceil(q);
  const int rstep = AT_THREADS >> qs;
  const int d = (threadIdx.x & ((1 << qs) - 1)) * 4;
  if (d >= dc16) return;
  const float* colp = src + (long)b * ld + hoff + c0 + d;
#pragma unroll 2
  for (int r = threadIdx.x >> qs; r < nr16; r += rstep) {
    const int t = r0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < len && d < dc) {
      const float* p = colp + (long)t * B * ld;
      if (vec) {
        v = *reinterpret_cast<const float4*>(p);
      } else {
        v.x = p[0];
        if (d + 1 < dc) v.y = p[1];
        if (d + 2 < dc) v.z = p[2];
        if (d + 3 < dc) v.w = p[3];
      }
    }
    const uint2 hi = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(dst + (size_t)r * dstr + d) = hi;
    if (dst_lo) {                                  // fp32 mode: the residual plane x - bf16(x)
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hi);
      const float2 h01 = __bfloat1622float2(h2[0]), h23 = __bfloat1622float2(h2[1]);
      *reinterpret_cast<uint2*>(dst_lo + (size_t)r * dstr + d) =
          make_uint2(pack_bf16x2(v.x - h01.x, v.y - h01.y), pack_bf16x2(v.z - h23.x, v.w - h23.y));
    }
  }
}

// sc[r][0..dc) (fp32, smem, row stride lstr) * scale -> dst[((row0 + r) * B + b) * ld + col + d] (+ optional bf16 copy)
__device__ __forceinline__ void store_rows_f32(float* dst, __nv_bfloat16* dst_b, long ld, int B, int b, int col, int row0,
                                               int nrows, int dc, const float* sc, int lstr, float scale) {
  const bool vec = ((dc & 3) == 0) && ((ld & 3) == 0) && ((col & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                   (!dst_b || (reinterpret_cast<uintptr_t>(dst_b) & 7) == 0);
  if (vec) {
    const int q = dc >> 2;
    const int qs = log2_ceil(q);
    const int rstep = AT_THREADS >> qs;
    const int d = (threadIdx.x & ((1 << qs) - 1)) * 4;
    if (d >= dc) return;
    for (int r = threadIdx.x >> qs; r < nrows; r += rstep) {
      float4 v = *reinterpret_cast<const float4*>(sc + (size_t)r * lstr + d);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      const long o = ((long)(row0 + r) * B + b) * ld + col + d;
      *reinterpret_cast<float4*>(dst + o) = v;
      if (dst_b) *reinterpret_cast<uint2*>(dst_b + o) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
  } else {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < nrows; r += AT_WARPS)
      for (int d = lane; d < dc; d += 32) {
        const float v = sc[(size_t)r * lstr + d] * scale;
        const long o = ((long)(row0 + r) * B + b) * ld + col + d;
        dst[o] = v;
        if (dst_b) dst_b[o] = __float2bfloat16(v);
      }
  }
}

// C[M16 x N16] (fp32, ldc) (+)= A[M16 x K16] (bf16 row-major, lda) * B
//   B_COL = true : B given as Y[N16 x K16] row-major (i.e. C = A * Y^T)
//   B_COL = false: B given as Y[K16 x N16] row-major
//   Alo / Blo: the residual planes of the operands (fp32 mode: three passes hi*hi + lo*hi + hi*lo), or null
template <bool B_COL>
__device__ __forceinline__ void tile_mm(const __nv_bfloat16* A, int lda, const __nv_bfloat16* Bm, int ldb, float* C, int ldc,
                                        int M16, int N16, int K16, bool accumulate, const __nv_bfloat16* Alo = nullptr,
                                        const __nv_bfloat16* Blo = nullptr) {
  const int warp = threadIdx.x >> 5;
  const int nt = N16 >> 4, tiles = (M16 >> 4) * nt;
  for (int tile = warp; tile < tiles; tile += AT_WARPS) {
    const int m0 = (tile / nt) << 4, n0 = (tile % nt) << 4;
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
    if (accumulate)
      wmma::load_matrix_sync(acc, C + (size_t)m0 * ldc + n0, ldc, wmma::mem_row_major);
    else
      wmma::fill_fragment(acc, 0.f);
    for (int k0 = 0; k0 < K16; k0 += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> fa;
      wmma::load_matrix_sync(fa, A + (size_t)m0 * lda + k0, lda);
      if (B_COL) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::col_major> fb;
        wmma::load_matrix_sync(fb, Bm + (size_t)n0 * ldb + k0, ldb);
        wmma::mma_sync(acc, fa, fb, acc);
        if (Alo) {
          wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::col_major> fl;
          wmma::load_matrix_sync(fl, Blo + (size_t)n0 * ldb + k0, ldb);
          wmma::mma_sync(acc, fa, fl, acc);                                  // hi * lo
          wmma::load_matrix_sync(fa, Alo + (size_t)m0 * lda + k0, lda);
          wmma::mma_sync(acc, fa, fb, acc);                                  // lo * hi
        }
      } else {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fb;
        wmma::load_matrix_sync(fb, Bm + (size_t)k0 * ldb + n0, ldb);
        wmma::mma_sync(acc, fa, fb, acc);
        if (Alo) {
          wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fl;
          wmma::load_matrix_sync(fl, Blo + (size_t)k0 * ldb + n0, ldb);
          wmma::mma_sync(acc, fa, fl, acc);
          wmma::load_matrix_sync(fa, Alo + (size_t)m0 * lda + k0, lda);
          wmma::mma_sync(acc, fa, fb, acc);
        }
      }
    }
    wmma::store_matrix_sync(C + (size_t)m0 * ldc + n0, acc, ldc, wmma::mem_row_major);
  }
}

// C[M16 x N16] (fp32, ldc) = A^T * B with A given as [K16 x M16] row-major (the dS block as stored: rows = queries) and
// B as [K16 x N16] row-major: the dS^T q product of the key gradient, taken straight from the query-side tiles
__device__ __forceinline__ void tile_mm_at(const __nv_bfloat16* A, int lda, const __nv_bfloat16* Bm, int ldb, float* C, int ldc,
                                           int M16, int N16, int K16, const __nv_bfloat16* Alo = nullptr,
                                           const __nv_bfloat16* Blo = nullptr) {
  const int warp = threadIdx.x >> 5;
  const int nt = N16 >> 4, tiles = (M16 >> 4) * nt;
  for (int tile = warp; tile < tiles; tile += AT_WARPS) {
    const int m0 = (tile / nt) << 4, n0 = (tile % nt) << 4;
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
    wmma::fill_fragment(acc, 0.f);
    for (int k0 = 0; k0 < K16; k0 += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> fa;
      wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fb;
      wmma::load_matrix_sync(fa, A + (size_t)k0 * lda + m0, lda);
      wmma::load_matrix_sync(fb, Bm + (size_t)k0 * ldb + n0, ldb);
      wmma::mma_sync(acc, fa, fb, acc);
      if (Alo) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fl;
        wmma::load_matrix_sync(fl, Blo + (size_t)k0 * ldb + n0, ldb);
        wmma::mma_sync(acc, fa, fl, acc);
        wmma::load_matrix_sync(fa, Alo + (size_t)k0 * lda + m0, lda);
        wmma::mma_sync(acc, fa, fb, acc);
      }
    }
    wmma::store_matrix_sync(C + (size_t)m0 * ldc + n0, acc, ldc, wmma::mem_row_major);
  }
}

__device__ __forceinline__ bool is_masked(const AttnArgs& a, int b, int t, int j) {
  if (a.key_pad && a.key_pad[(long)j * a.B + b]) return true;
  if (a.attn_mask && a.attn_mask[(long)t * a.S + j]) return true;
  return false;
}

// ============================================================================================
// forward
// ============================================================================================
template <bool SPLIT>
__global__ void __launch_bounds__(AT_THREADS, SPLIT ? 2 : AT_MIN_CTAS) attn_fwd_kernel(const AttnArgs a, const int rows) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ __align__(128) uint8_t smem_u8[];
  const int dc = a.hd < AT_DC ? a.hd : AT_DC;
  const int dc16 = r16(dc);
  const AttnSmem s = carve<SPLIT>(smem_u8, a.S, dc, rows);
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int t0 = blockIdx.y * rows;
  const int nrows = (a.T - t0) < rows ? (a.T - t0) : rows;
  const int hoff = h * a.hd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.S;
  ATT_TRACE(0, 0);

  // ---- scores into sc ----
  if (a.scores_jt) {
    for (int j = warp; j < s.L16; j += AT_WARPS)
      for (int r = lane; r < s.R16; r += 32)
        s.sc[(size_t)r * s.lstr + j] = (r < nrows && j < S) ? a.scores_jt[((long)bh * S + j) * a.T + t0 + r] : 0.f;
  } else {
    for (int c0 = 0; c0 < a.hd; c0 += dc) {
      __syncthreads();
      load_rows_bf16(s.xb, s.dstr, a.q, a.ldq, a.B, b, hoff, t0, s.R16, a.T, c0, dc, s.xl);
      load_rows_bf16(s.yb, s.dstr, a.k, a.ldk, a.B, b, hoff, 0, s.L16, S, c0, dc, s.yl);
      __syncthreads();
      tile_mm<true>(s.xb, s.dstr, s.yb, s.dstr, s.sc, s.lstr, s.R16, s.L16, dc16, c0 > 0, s.xl, s.yl);
    }
  }
  __syncthreads();
  ATT_TRACE(0, 1);

  // ---- masks, softmax, dropout; probabilities -> global (fp32) and pb (bf16 operand) ----
  const unsigned long long seed = a.p_drop > 0.f ? reinterpret_cast<const unsigned long long*>(a.seed_ptr)[0] + a.seed_off : 0ull;
  const float ks = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const float sscale = a.scores_jt ? 1.f : a.scale;
  for (int r = warp; r < s.R16; r += AT_WARPS) {
    float* w = s.sc + (size_t)r * s.lstr;
    __nv_bfloat16* pw = s.pb + (size_t)r * s.lstr;
    __nv_bfloat16* pl = SPLIT ? s.pl + (size_t)r * s.lstr : nullptr;
    if (r >= nrows) {
      for (int j = lane; j < s.L16; j += 32) put_operand<SPLIT>(pw, pl, j, 0.f);
      continue;
    }
    const int t = t0 + r;
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) {         // masks + scale fused into the max pass (same lane owns w[j] below)
      const float v = is_masked(a, b, t, j) ? -INFINITY : w[j] * sscale;
      w[j] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      float e = (w[j] == -INFINITY) ? 0.f : (SPLIT ? expf(w[j] - mx) : __expf(w[j] - mx));
      w[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    const long prow = ((long)bh * a.T + t) * S;
    for (int j = lane; j < s.L16; j += 32) {
      float p = 0.f;
      if (j < S) {
        p = w[j] * inv;
        a.probs[prow + j] = p;
        if (a.p_drop > 0.f) p = (rng_uniform(seed, (unsigned long long)(prow + j)) >= a.p_drop) ? p * ks : 0.f;
        if (a.probs_dropped) a.probs_dropped[prow + j] = p;
      }
      put_operand<SPLIT>(pw, pl, j, p);
    }
  }

  // ---- PV: out[r][d] = sum_j pb[r][j] * V[j][d], staged in sc, written coalesced ----
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  ATT_TRACE(0, 2);
  for (int c0 = 0; c0 < a.hd; c0 += dc) {
    __syncthreads();
    load_rows_bf16(s.yb, s.dstr, a.v, a.ldv, a.B, b, hoff, 0, s.L16, S, c0, dc, s.yl);
    __syncthreads();
    ATT_TRACE(0, 3);
    tile_mm<false>(s.pb, s.lstr, s.yb, s.dstr, s.sc, s.lstr, s.R16, dc16, s.L16, false, s.pl, s.yl);
    __syncthreads();
    ATT_TRACE(0, 4);
    store_rows_f32(a.out, ob, a.ldo, a.B, b, hoff + c0, t0, nrows, dc, s.sc, s.lstr, 1.f);
  }
  ATT_TRACE(0, 5);
}

// rows per CTA: the whole sequence when there are plenty of (batch, head) CTAs, otherwise split the rows so
// that the grid still covers ~2 CTAs per SM (e.g. the 1-head alignment attention, decoder.py:13)
static int attn_rows(int len, int bh) {
  int want_y = (2 * 148 + bh - 1) / bh;
  int rows = (len + want_y - 1) / want_y;
  rows = r16(rows);
  if (rows > AT_ROWS_MAX) rows = 32;
  return rows;
}
static size_t attn_smem_bytes(int L, int hd, int rows, bool split) {
  int dc = hd < AT_DC ? hd : AT_DC;
  int a, b2, c, d;
  return attn_smem_layout(L, dc, rows, &a, &b2, &c, &d, split) + 128;
}

int attn_debug_read_trace(unsigned long long* host_out, int enable) {
  GTOS_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_attn_trace, sizeof(unsigned long long) * 48));
  GTOS_CHECK_CUDA(cudaMemcpyToSymbol(g_attn_trace_on, &enable, sizeof(int)));
  return GTOS_OK;
}

int attn_fwd(const AttnArgs& a, cudaStream_t st) {
  GTOS_REQUIRE(a.hd >= 1 && (a.hd <= AT_DC || a.hd % AT_DC == 0), "attention: unsupported head_dim %d", a.hd);
  GTOS_REQUIRE(a.p_drop == 0.f || a.seed_ptr, "attention dropout needs a device seed pointer");
  if (a.T == 0 || a.B == 0) return GTOS_OK;
  const bool split = a.precise != 0;
  int rows = attn_rows(a.T, a.B * a.H);
  size_t smem = attn_smem_bytes(a.S, a.hd, rows, split);
  if (split && smem > 227 * 1024 && rows > 16) {      // the lo planes double the operand tiles: fall back to 16-row blocks
    rows = 16;
    smem = attn_smem_bytes(a.S, a.hd, rows, split);
  }
  GTOS_REQUIRE(smem <= 227 * 1024, "attention: source length %d too long for shared memory", a.S);
  dim3 grid(a.B * a.H, (a.T + rows - 1) / rows);
  if (split) {
    GTOS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GTOS_KLAUNCH(attn_fwd_kernel<true>, dim3(grid), dim3(AT_THREADS), smem, st, a, rows);
  } else {
    GTOS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GTOS_KLAUNCH(attn_fwd_kernel<false>, dim3(grid), dim3(AT_THREADS), smem, st, a, rows);
  }
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ============================================================================================
// backward, query side: dS (and dq in decoder mode)
// ============================================================================================
template <bool SPLIT>
__global__ void __launch_bounds__(AT_THREADS, SPLIT ? 2 : AT_MIN_CTAS) attn_bwd_q_kernel(const AttnBwdArgs g, const int rows) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ __align__(128) uint8_t smem_u8[];
  const AttnArgs& a = g.f;
  const int dc = a.hd < AT_DC ? a.hd : AT_DC;
  const int dc16 = r16(dc);
  const AttnSmem s = carve<SPLIT>(smem_u8, a.S, dc, rows);
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int t0 = blockIdx.y * rows;
  const int nrows = (a.T - t0) < rows ? (a.T - t0) : rows;
  const int hoff = h * a.hd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.S;

  // dPd[t][j] = dO[t] . V[j]
  ATT_TRACE(1, 0);
  for (int c0 = 0; c0 < a.hd; c0 += dc) {
    __syncthreads();
    load_rows_bf16(s.xb, s.dstr, g.dout, g.lddo, a.B, b, hoff, t0, s.R16, a.T, c0, dc, s.xl);
    load_rows_bf16(s.yb, s.dstr, a.v, a.ldv, a.B, b, hoff, 0, s.L16, S, c0, dc, s.yl);
    __syncthreads();
    ATT_TRACE(1, 1);
    tile_mm<true>(s.xb, s.dstr, s.yb, s.dstr, s.sc, s.lstr, s.R16, s.L16, dc16, c0 > 0, s.xl, s.yl);
  }
  __syncthreads();
  ATT_TRACE(1, 2);

  const unsigned long long seed = a.p_drop > 0.f ? reinterpret_cast<const unsigned long long*>(a.seed_ptr)[0] + a.seed_off : 0ull;
  const float ks = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  for (int r = warp; r < s.R16; r += AT_WARPS) {
    float* w = s.sc + (size_t)r * s.lstr;
    __nv_bfloat16* pw = s.pb + (size_t)r * s.lstr;
    __nv_bfloat16* pl = SPLIT ? s.pl + (size_t)r * s.lstr : nullptr;
    if (r >= nrows) {
      for (int j = lane; j < s.L16; j += 32) {
        put_operand<SPLIT>(pw, pl, j, 0.f);
        w[j] = 0.f;
      }
      continue;
    }
    const int t = t0 + r;
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
    for (int j = lane; j < s.L16; j += 32) {
      float ds = 0.f;
      if (j < S) {
        ds = a.probs[prow + j] * (w[j] - dot);
        g.dscores_ts[prow + j] = ds;
      }
      w[j] = ds;
      put_operand<SPLIT>(pw, pl, j, ds);
    }
  }
  __syncthreads();
  ATT_TRACE(1, 3);
  if (g.dscores_jt) {
    // transposed store for the fused relation backward kernel: [B,H,S(j),T(i)], coalesced along i
    for (int j = warp; j < S; j += AT_WARPS)
      for (int r = lane; r < nrows; r += 32)
        g.dscores_jt[((long)bh * S + j) * a.T + t0 + r] = s.sc[(size_t)r * s.lstr + j];
  }
  ATT_TRACE(1, 4);
  if (g.dq) {
    for (int c0 = 0; c0 < a.hd; c0 += dc) {
      __syncthreads();
      load_rows_bf16(s.yb, s.dstr, a.k, a.ldk, a.B, b, hoff, 0, s.L16, S, c0, dc, s.yl);
      __syncthreads();
      tile_mm<false>(s.pb, s.lstr, s.yb, s.dstr, s.sc, s.lstr, s.R16, dc16, s.L16, false, s.pl, s.yl);
      __syncthreads();
      store_rows_f32(g.dq, reinterpret_cast<__nv_bfloat16*>(g.dq_bf16), g.lddq, a.B, b, hoff + c0, t0, nrows, dc, s.sc, s.lstr, a.scale);
    }
  }
  // dK = scale * dS^T q right here when this CTA holds EVERY query row of its (batch, head) (a single row block, and the
  // [keys x dc] result fits the score tile): the key-side kernel is then left with dV = Pd^T dO, which needs nothing from
  // this kernel - the two run side by side instead of one after the other (attn_bwd_dk_on_query_side)
  if (g.dk_in_q && g.dk) {
    for (int c0 = 0; c0 < a.hd; c0 += dc) {
      __syncthreads();
      load_rows_bf16(s.xb, s.dstr, a.q, a.ldq, a.B, b, hoff, 0, s.R16, a.T, c0, dc, s.xl);
      __syncthreads();
      tile_mm_at(s.pb, s.lstr, s.xb, s.dstr, s.sc, s.lstr, s.L16, dc16, s.R16, s.pl, s.xl);
      __syncthreads();
      store_rows_f32(g.dk, reinterpret_cast<__nv_bfloat16*>(g.dk_bf16), g.lddk, a.B, b, hoff + c0, 0, S, dc, s.sc, s.lstr, a.scale);
    }
  }
}

// ============================================================================================
// backward, key side: dV = Pd^T dO ; dK = scale * dS^T q      (CTA = `rows` key rows of one (b,h))
// ============================================================================================
template <bool SPLIT>
__global__ void __launch_bounds__(AT_THREADS, SPLIT ? 2 : AT_MIN_CTAS) attn_bwd_kv_kernel(const AttnBwdArgs g, const int rows) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ __align__(128) uint8_t smem_u8[];
  const AttnArgs& a = g.f;
  const int dc = a.hd < AT_DC ? a.hd : AT_DC;
  const int dc16 = r16(dc);
  const int T = a.T, S = a.S;
  const AttnSmem s = carve<SPLIT>(smem_u8, T, dc, rows);   // L = T: the contraction runs over the queries
  const int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  const int j0 = blockIdx.y * rows;
  const int nrows = (S - j0) < rows ? (S - j0) : rows;
  const int hoff = h * a.hd;
  const unsigned long long seed = a.p_drop > 0.f ? reinterpret_cast<const unsigned long long*>(a.seed_ptr)[0] + a.seed_off : 0ull;
  const float ks = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;

  // pb[jr][t] = Pd[t][j0+jr]  (transposed while loading; coalesced over jr in global)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ATT_TRACE(2, 0);
  for (int t = warp; t < s.L16; t += AT_WARPS)
    for (int jr = lane; jr < s.R16; jr += 32) {
      float p = 0.f;
      if (jr < nrows && t < T) {
        const long pi = ((long)bh * T + t) * S + j0 + jr;
        p = a.probs[pi];
        if (a.p_drop > 0.f) p = (rng_uniform(seed, (unsigned long long)pi) >= a.p_drop) ? p * ks : 0.f;
      }
      put_operand<SPLIT>(s.pb, s.pl, jr * s.lstr + t, p);
    }
  ATT_TRACE(2, 1);
  for (int c0 = 0; c0 < a.hd; c0 += dc) {
    __syncthreads();
    load_rows_bf16(s.yb, s.dstr, g.dout, g.lddo, a.B, b, hoff, 0, s.L16, T, c0, dc, s.yl);
    __syncthreads();
    ATT_TRACE(2, 2);
    tile_mm<false>(s.pb, s.lstr, s.yb, s.dstr, s.sc, s.lstr, s.R16, dc16, s.L16, false, s.pl, s.yl);
    __syncthreads();
    ATT_TRACE(2, 3);
    store_rows_f32(g.dv, reinterpret_cast<__nv_bfloat16*>(g.dv_bf16), g.lddv, a.B, b, hoff + c0, j0, nrows, dc, s.sc, s.lstr, 1.f);
  }
  ATT_TRACE(2, 4);
  if (g.dk) {
    __syncthreads();
    for (int t = warp; t < s.L16; t += AT_WARPS)
      for (int jr = lane; jr < s.R16; jr += 32) {
        const float v = (jr < nrows && t < T) ? g.dscores_ts[((long)bh * T + t) * S + j0 + jr] * a.scale : 0.f;
        put_operand<SPLIT>(s.pb, s.pl, jr * s.lstr + t, v);
      }
    for (int c0 = 0; c0 < a.hd; c0 += dc) {
      __syncthreads();
      load_rows_bf16(s.yb, s.dstr, a.q, a.ldq, a.B, b, hoff, 0, s.L16, T, c0, dc, s.yl);
      __syncthreads();
      tile_mm<false>(s.pb, s.lstr, s.yb, s.dstr, s.sc, s.lstr, s.R16, dc16, s.L16, false, s.pl, s.yl);
      __syncthreads();
      store_rows_f32(g.dk, reinterpret_cast<__nv_bfloat16*>(g.dk_bf16), g.lddk, a.B, b, hoff + c0, j0, nrows, dc, s.sc, s.lstr, 1.f);
    }
  }
}

// 1 if gtos_attn_bwd's query-side kernel also produces dK for this shape (decoder mode, one query block per (batch, head))
int attn_bwd_dk_on_query_side(const AttnArgs& a, bool decoder_mode) {
  if (!decoder_mode || a.T == 0 || a.B == 0) return 0;
  static const bool enabled = !(getenv("GTOS_ATTN_DK_Q") && getenv("GTOS_ATTN_DK_Q")[0] == '0');
  if (!enabled) return 0;
  const bool split = a.precise != 0;
  int rq = attn_rows(a.T, a.B * a.H);
  if (split && attn_smem_bytes(a.S, a.hd, rq, split) > 227 * 1024 && rq > 16) rq = 16;
  return (rq >= a.T && r16(a.S) <= r16(rq)) ? 1 : 0;
}

int attn_bwd(const AttnBwdArgs& g_in, int part, cudaStream_t st) {
  AttnBwdArgs g = g_in;
  const AttnArgs& a = g.f;
  // part 0 / 1: the query side takes dK when it can; part 0: the key side then skips it.  part 2 alone computes whatever
  // the caller asks for (dk == NULL: dV only)
  const int dk_q = (part != 2 && g.dq && g.dk) ? attn_bwd_dk_on_query_side(a, true) : 0;
  g.dk_in_q = dk_q;
  GTOS_REQUIRE(a.hd >= 1 && (a.hd <= AT_DC || a.hd % AT_DC == 0), "attention: unsupported head_dim %d", a.hd);
  if (a.T == 0 || a.B == 0) return GTOS_OK;
  const bool split = a.precise != 0;
  int rq = attn_rows(a.T, a.B * a.H), rk = attn_rows(a.S, a.B * a.H);
  size_t smq = attn_smem_bytes(a.S, a.hd, rq, split), smk = attn_smem_bytes(a.T, a.hd, rk, split);
  if (split && smq > 227 * 1024 && rq > 16) { rq = 16; smq = attn_smem_bytes(a.S, a.hd, rq, split); }
  if (split && smk > 227 * 1024 && rk > 16) { rk = 16; smk = attn_smem_bytes(a.T, a.hd, rk, split); }
  GTOS_REQUIRE(smq <= 227 * 1024 && smk <= 227 * 1024, "attention: sequence too long for shared memory");
  if (part != 2) {
    dim3 gq(a.B * a.H, (a.T + rq - 1) / rq);
    if (split) {
      GTOS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_q_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smq));
      GTOS_KLAUNCH(attn_bwd_q_kernel<true>, dim3(gq), dim3(AT_THREADS), smq, st, g, rq);
    } else {
      GTOS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_q_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smq));
      GTOS_KLAUNCH(attn_bwd_q_kernel<false>, dim3(gq), dim3(AT_THREADS), smq, st, g, rq);
    }
    GTOS_LAUNCH_CHECK();
  }
  if (part != 1) {
    if (dk_q) {                                   // dK came from the query side
      g.dk = nullptr;
      g.dk_bf16 = nullptr;
    }
    dim3 gk(a.B * a.H, (a.S + rk - 1) / rk);
    if (split) {
      GTOS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smk));
      GTOS_KLAUNCH(attn_bwd_kv_kernel<true>, dim3(gk), dim3(AT_THREADS), smk, st, g, rk);
    } else {
      GTOS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smk));
      GTOS_KLAUNCH(attn_bwd_kv_kernel<false>, dim3(gk), dim3(AT_THREADS), smk, st, g, rk);
    }
    GTOS_LAUNCH_CHECK();
  }
  return GTOS_OK;
}

}  // namespace gtos
