// gtos_b200 -- bank-factorised relation attention (forward half; caller generator/generator.py:76-90).
//
// relation = bank[idx] and relation_in_proj has no bias (generator/graph_transformer.py:80), so
//     [ra_ij | rb_ij] = relation_in_proj(relation[j][i]) = PB[idx[j][i][b]],      PB = bank * W_r^T   ([R, 2D], one GEMM)
// and the P-row projection GEMM of graph_transformer.py:122 (4 D^2 FLOP per node pair) collapses to an R-row GEMM plus a
// gather of one 2D-wide bf16 row per pair out of an L2-resident table.  With the tensor work gone, the whole attention
// of graph_transformer.py:122-159 fits ONE kernel per layer:
//
//   rel_attn_banked_fwd   CTA = (graph b, QI queries i, all keys j): gather PB rows, s_ij = hd^-1/2 <q_i + ra, k_j + rb>
//                         per head, key-padding / attention mask, softmax over j, dropout, o_i = sum_j w_ij v_j.
//                         Scores live in shared memory only.  L2-gather bound: 2 * 2D bytes per pair.
//   rel_grad_banked       the backward's per-pair gradient rows G = hd^-1/2 ds_ij [k_j + rb | q_i + ra] from the same gather
//                         (replaces the P-row recompute GEMM of gtos_rel_grad), written in gtos_rel_grad's tile-major
//                         layout so rel_segsum / rel_dqk / rel_dw_bank consume it unchanged.
//
// Measured alternatives (config 2, one layer; tools/time_banked.py): rows of a key kept in registers, 4 queries in flight,
// one CTA per SM: forward 84.6 us, gradient 86 us; queries in pairs + 2 CTAs per SM: 84.4 / 100 us; rows fetched by TMA 1-D
// bulk copies (cp.async.bulk, one 2 KB row per instruction, two mbarrier stages of 32 rows): 119 / 119 us - UBLKCP costs
// ~250 cycles per 2 KB row, far below the LDG path.  ncu (profiles/r02_ncu_banked.txt): long-scoreboard bound at 16
// resident warps per SM, 13 B/clk/SM.
//
// Lane layout (both kernels): lane l of a warp owns DL = D/32 consecutive features of head h = l / (32/H); it loads the
// matching DL-element pieces of ra and rb (2 * DL * 2 bytes), so one warp instruction pair covers a whole 2D-wide row and
// a head's dot product is finished with log2(32/H) shuffles.
#include <stdlib.h>

#include "elementwise.cuh"
#include "gemm.cuh"

namespace gtos {

struct RelBankedDev {
  const __nv_bfloat16* PB; long ldpb;
  const long long* idx;
  const __nv_bfloat16* q; const __nv_bfloat16* k; long ldqk;
  const float* v; long ldv;
  const uint8_t* key_pad; const uint8_t* attn_mask;
  float p_drop; const void* seed_ptr; unsigned long long seed_off;
  float* probs; float* probs_dropped;
  float* out; long ldo; __nv_bfloat16* out_bf16;
  const float* dscores;        // [B,H,N(j),N(i)]
  __nv_bfloat16* G;
  RelTiling rt;
  int N, B, D, H, hd, R, Npad;
  float scale;
};

template <int DL>
__device__ __forceinline__ void ld_bf16_raw(const __nv_bfloat16* p, uint32_t* r) {   // DL bf16 -> DL/2 packed words
  if constexpr (DL == 4) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    r[0] = v.x; r[1] = v.y;
  } else {
#pragma unroll
    for (int t = 0; t < DL / 8; ++t) {
      const uint4 v = *reinterpret_cast<const uint4*>(p + 8 * t);
      r[4 * t] = v.x; r[4 * t + 1] = v.y; r[4 * t + 2] = v.z; r[4 * t + 3] = v.w;
    }
  }
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// bank row of pair (query i, key j, graph b); out-of-range rows read row 0 (never happens with a well-formed batch)
__device__ __forceinline__ int bank_row(const RelBankedDev& a, int j, int i, int b) {
  if (i >= a.N) return 0;
  const long long r = a.idx[((long)j * a.N + i) * a.B + b];
  return (r >= 0 && r < a.R) ? (int)r : 0;
}

template <int DL, int QI>
__global__ void __launch_bounds__(256, 2) rel_attn_banked_fwd_kernel(const RelBankedDev a) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ float sm[];
  float* sc = sm;                                                   // [QI][H][Npad] scores -> (dropped) probabilities
  float* part = sm + QI * a.H * a.Npad;                             // [8 warps][QI][D] partial outputs of the PV pass
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * QI, b = blockIdx.y;
  const int N = a.N, B = a.B, H = a.H, hd = a.hd;
  const int lph = 32 / H;                      // lanes per head
  const int h = lane / lph, sub = lane - h * lph;
  const int dcol = h * hd + sub * DL;          // this lane's features inside D (q, k, v)
  const int pcol = h * 2 * hd + sub * DL;      // its ra piece inside the head-interleaved PB row; rb piece at + hd

  uint32_t qraw[QI][DL / 2];                   // q_i pieces, packed bf16 (unpacked where they are used: fewer live registers)
#pragma unroll
  for (int qi = 0; qi < QI; ++qi) {
#pragma unroll
    for (int t = 0; t < DL / 2; ++t) qraw[qi][t] = 0u;
    if (i0 + qi < N) ld_bf16_raw<DL>(a.q + ((long)(i0 + qi) * B + b) * a.ldqk + dcol, qraw[qi]);
  }

  // ---- scores: one warp per key j, QI queries two at a time; the bank rows of the next key are fetched a step ahead ----
  int rows_next[QI];
#pragma unroll
  for (int qi = 0; qi < QI; ++qi) rows_next[qi] = warp < N ? bank_row(a, warp, i0 + qi, b) : 0;
  for (int j = warp; j < N; j += 8) {
    int rows[QI];
#pragma unroll
    for (int qi = 0; qi < QI; ++qi) rows[qi] = rows_next[qi];
    if (j + 8 < N) {
#pragma unroll
      for (int qi = 0; qi < QI; ++qi) rows_next[qi] = bank_row(a, j + 8, i0 + qi, b);
    }
    uint32_t kraw[DL / 2];
    ld_bf16_raw<DL>(a.k + ((long)j * B + b) * a.ldqk + dcol, kraw);
#pragma unroll
    for (int q0 = 0; q0 < QI; q0 += 2) {
      uint32_t ra[2][DL / 2], rb[2][DL / 2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const __nv_bfloat16* pr = a.PB + (long)rows[q0 + u] * a.ldpb + pcol;
        ld_bf16_raw<DL>(pr, ra[u]);
        ld_bf16_raw<DL>(pr + hd, rb[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < DL / 2; ++t) {
          acc = fmaf(bf_lo(qraw[q0 + u][t]) + bf_lo(ra[u][t]), bf_lo(kraw[t]) + bf_lo(rb[u][t]), acc);
          acc = fmaf(bf_hi(qraw[q0 + u][t]) + bf_hi(ra[u][t]), bf_hi(kraw[t]) + bf_hi(rb[u][t]), acc);
        }
        for (int o = lph >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (sub == 0) sc[((q0 + u) * H + h) * a.Npad + j] = acc * a.scale;
      }
    }
  }
  __syncthreads();

  // ---- masks, softmax over keys, dropout (same counter-based draw as the attention core, so gtos_attn_bwd replays it) ----
  const unsigned long long seed = a.p_drop > 0.f ? reinterpret_cast<const unsigned long long*>(a.seed_ptr)[0] + a.seed_off : 0ull;
  const float ks = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  for (int r = warp; r < QI * H; r += 8) {
    const int qi = r / H, hh = r - qi * H, i = i0 + qi;
    float* w = sc + r * a.Npad;
    if (i >= N) {
      for (int j = lane; j < N; j += 32) w[j] = 0.f;
      continue;
    }
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      const bool masked = (a.key_pad && a.key_pad[(long)j * B + b]) || (a.attn_mask && a.attn_mask[(long)i * N + j]);
      const float v = masked ? -INFINITY : w[j];
      w[j] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = (w[j] == -INFINITY) ? 0.f : __expf(w[j] - mx);
      w[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    const long prow = (((long)b * H + hh) * N + i) * N;
    for (int j = lane; j < N; j += 32) {
      float p = w[j] * inv;
      a.probs[prow + j] = p;
      if (a.p_drop > 0.f) p = (rng_uniform(seed, (unsigned long long)(prow + j)) >= a.p_drop) ? p * ks : 0.f;
      if (a.probs_dropped) a.probs_dropped[prow + j] = p;
      w[j] = p;
    }
  }
  __syncthreads();

  // ---- PV: o_i = sum_j w_ij v_j.  Same split as the scores (warp = keys j = warp, warp + 8, ..; lane = DL features of its
  // head): every v row is read once per CTA with all of a warp's loads independent; the 8 partial sums meet in shared memory
  {
    float acc[QI][DL];
#pragma unroll
    for (int qi = 0; qi < QI; ++qi)
#pragma unroll
      for (int t = 0; t < DL; ++t) acc[qi][t] = 0.f;
    for (int j = warp; j < N; j += 8) {
      float vv[DL];
      const float4* vp = reinterpret_cast<const float4*>(a.v + ((long)j * B + b) * a.ldv + dcol);
#pragma unroll
      for (int t = 0; t < DL / 4; ++t) {
        const float4 f = vp[t];
        vv[4 * t] = f.x; vv[4 * t + 1] = f.y; vv[4 * t + 2] = f.z; vv[4 * t + 3] = f.w;
      }
#pragma unroll
      for (int qi = 0; qi < QI; ++qi) {
        const float p = sc[(qi * H + h) * a.Npad + j];
#pragma unroll
        for (int t = 0; t < DL; ++t) acc[qi][t] = fmaf(p, vv[t], acc[qi][t]);
      }
    }
#pragma unroll
    for (int qi = 0; qi < QI; ++qi) {
      float4* dst = reinterpret_cast<float4*>(part + ((long)warp * QI + qi) * a.D + dcol);
#pragma unroll
      for (int t = 0; t < DL / 4; ++t) dst[t] = make_float4(acc[qi][4 * t], acc[qi][4 * t + 1], acc[qi][4 * t + 2], acc[qi][4 * t + 3]);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < QI * (a.D / 2); e += 256) {
    const int qi = e / (a.D / 2), f = 2 * (e - qi * (a.D / 2));
    const int i = i0 + qi;
    if (i >= N) continue;
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) {
      const float2 t2 = *reinterpret_cast<const float2*>(part + ((long)w8 * QI + qi) * a.D + f);
      s2.x += t2.x; s2.y += t2.y;
    }
    const long o = ((long)i * B + b);
    *reinterpret_cast<float2*>(a.out + o * a.ldo + f) = s2;
    if (a.out_bf16) *reinterpret_cast<uint32_t*>(a.out_bf16 + o * a.D + f) = pack_bf16x2(s2.x, s2.y);
  }
}

// One CTA per SM with all four queries' gathers of a key in flight per lane measured faster here (86 us) than two CTAs per SM
// with the queries taken in pairs (100 us): the kernel is bound by loads in flight per warp, not by resident warps.
template <int DL, int QI>
__global__ void __launch_bounds__(256) rel_grad_banked_kernel(const RelBankedDev a) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ float sm[];
  int* s_idx = reinterpret_cast<int*>(sm);                          // [N][QI]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * QI, b = blockIdx.y;
  const int N = a.N, B = a.B, H = a.H, hd = a.hd;
  const int lph = 32 / H;
  const int h = lane / lph, sub = lane - h * lph;
  const int dcol = h * hd + sub * DL;
  const int pcol = h * 2 * hd + sub * DL;
  float qv[QI][DL];
#pragma unroll
  for (int qi = 0; qi < QI; ++qi) {
    const int i = i0 + qi;
    uint32_t raw[DL / 2];
#pragma unroll
    for (int t = 0; t < DL / 2; ++t) raw[t] = 0u;
    if (i < N) ld_bf16_raw<DL>(a.q + ((long)i * B + b) * a.ldqk + dcol, raw);
#pragma unroll
    for (int t = 0; t < DL / 2; ++t) { qv[qi][2 * t] = bf_lo(raw[t]); qv[qi][2 * t + 1] = bf_hi(raw[t]); }
  }
  for (int t = threadIdx.x; t < N * QI; t += 256) {
    const int j = t / QI, qi = t - j * QI, i = i0 + qi;
    long long r = (i < N) ? a.idx[((long)j * N + i) * B + b] : 0;
    s_idx[t] = (r >= 0 && r < a.R) ? (int)r : 0;
  }
  __syncthreads();
  const RelTiling& rt = a.rt;
  for (int j = warp; j < N; j += 8) {
    uint32_t kraw[DL / 2];
    ld_bf16_raw<DL>(a.k + ((long)j * B + b) * a.ldqk + dcol, kraw);
    const int jb = j / rt.bj, jj = j - jb * rt.bj;
    const float* dsp = a.dscores + (((long)b * H + h) * N + j) * N;
#pragma unroll
    for (int qi = 0; qi < QI; ++qi) {
      const int i = i0 + qi;
      if (i >= N) continue;                                        // warp-uniform
      const __nv_bfloat16* pr = a.PB + (long)s_idx[j * QI + qi] * a.ldpb + pcol;
      uint32_t ra[DL / 2], rb[DL / 2];
      ld_bf16_raw<DL>(pr, ra);
      ld_bf16_raw<DL>(pr + hd, rb);
      const float g = dsp[i] * a.scale;
      const int ib = i / rt.bi, ii = i - ib * rt.bi;
      const long row = (((long)b * rt.nj_blk + jb) * rt.ni_blk + ib) * 128 + jj * rt.bi + ii;
      __nv_bfloat16* gx = a.G + row * (2L * a.D) + pcol;           // d(q+ra) = g (k+rb) | d(k+rb) = g (q+ra)
      uint32_t wx[DL / 2], wy[DL / 2];
#pragma unroll
      for (int t = 0; t < DL / 2; ++t) {
        wx[t] = pack_bf16x2(g * (bf_lo(kraw[t]) + bf_lo(rb[t])), g * (bf_hi(kraw[t]) + bf_hi(rb[t])));
        wy[t] = pack_bf16x2(g * (qv[qi][2 * t] + bf_lo(ra[t])), g * (qv[qi][2 * t + 1] + bf_hi(ra[t])));
      }
      if constexpr (DL == 4) {
        *reinterpret_cast<uint2*>(gx) = make_uint2(wx[0], wx[1]);
        *reinterpret_cast<uint2*>(gx + hd) = make_uint2(wy[0], wy[1]);
      } else {
#pragma unroll
        for (int t = 0; t < DL / 8; ++t) {
          *reinterpret_cast<uint4*>(gx + 8 * t) = make_uint4(wx[4 * t], wx[4 * t + 1], wx[4 * t + 2], wx[4 * t + 3]);
          *reinterpret_cast<uint4*>(gx + hd + 8 * t) = make_uint4(wy[4 * t], wy[4 * t + 1], wy[4 * t + 2], wy[4 * t + 3]);
        }
      }
    }
  }
}

// ---- cp.async variant: the gather pipeline shared by both kernels -------------------------------------------------------
// A CTA owns (graph b, QI consecutive queries) and walks the keys in batches of KB = 8: the QI * KB = 32 projected-bank rows
// of a batch are copied global -> shared with 16-byte cp.async (LDGSTS: no registers, any number in flight) into one of two
// stages, so the rows of batch t + 1 are in flight while batch t is consumed (warp w = key w of the batch).  The register
// version keeps the rows of a key in registers and is bound by the loads a warp can have in flight (long scoreboard, 16
// warps per SM, 13 B/clk/SM: profiles/r02_ncu_banked.txt).
static constexpr int BK_KEYS = 8;

struct GatherPipe {
  uint8_t* stage[2];
  const int* s_idx;      // [N][QI] bank rows of the CTA's pairs
  int row_bytes;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <int QI>
__device__ __forceinline__ void gather_issue_ca(const RelBankedDev& a, const GatherPipe& gp, int bt, int i0) {
  // every thread copies 16-byte pieces tid, tid + 256, ... of the stage's 32 rows
  const int cpr = gp.row_bytes >> 4;                      // 16-byte pieces per row
  uint8_t* st = gp.stage[bt & 1];
  for (int c = threadIdx.x; c < 32 * cpr; c += 256) {
    const int row = c / cpr, off = c - row * cpr;
    const int kk = row / QI, qi = row - kk * QI;
    const int j = bt * BK_KEYS + kk;
    if (j < a.N && i0 + qi < a.N)
      cp_async16(st + (size_t)row * gp.row_bytes + off * 16,
                 reinterpret_cast<const uint8_t*>(a.PB + (long)gp.s_idx[j * QI + qi] * a.ldpb) + off * 16);
  }
  cp_async_commit();
}

template <int DL>
__device__ __forceinline__ void lds_bf16_raw(const uint8_t* p, uint32_t* r) {
  if constexpr (DL == 4) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    r[0] = v.x; r[1] = v.y;
  } else {
#pragma unroll
    for (int t = 0; t < DL / 8; ++t) {
      const uint4 v = *reinterpret_cast<const uint4*>(p + 16 * t);
      r[4 * t] = v.x; r[4 * t + 1] = v.y; r[4 * t + 2] = v.z; r[4 * t + 3] = v.w;
    }
  }
}

template <int DL, int QI>
__global__ void __launch_bounds__(256) rel_attn_banked_fwd_ca_kernel(const RelBankedDev a) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ __align__(128) uint8_t smem_u8[];
  const int N = a.N, B = a.B, H = a.H, hd = a.hd;
  GatherPipe gp;
  gp.row_bytes = 4 * a.D;                                            // 2D bf16
  const size_t stage_bytes = (size_t)32 * gp.row_bytes;
  gp.stage[0] = smem_u8;
  gp.stage[1] = smem_u8 + stage_bytes;
  float* sc = reinterpret_cast<float*>(smem_u8 + 2 * stage_bytes);  // [QI][H][Npad] scores -> (dropped) probabilities
  int* s_idx = reinterpret_cast<int*>(sc + QI * H * a.Npad);
  gp.s_idx = s_idx;
  float* part = reinterpret_cast<float*>(smem_u8);                   // [8 warps][QI][D] PV partials: reuses the stages
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * QI, b = blockIdx.y;
  const int lph = 32 / H;                      // lanes per head
  const int h = lane / lph, sub = lane - h * lph;
  const int dcol = h * hd + sub * DL;          // this lane's features inside D (q, k, v)
  const int pcol = h * 2 * hd + sub * DL;      // its ra piece inside the head-interleaved PB row; rb piece at + hd
  const int nb = (N + BK_KEYS - 1) / BK_KEYS;

  for (int t = threadIdx.x; t < N * QI; t += 256) {
    const int j = t / QI, qi = t - j * QI;
    s_idx[t] = (i0 + qi < N) ? bank_row(a, j, i0 + qi, b) : 0;
  }
  __syncthreads();
  gather_issue_ca<QI>(a, gp, 0, i0);
  uint32_t qraw[QI][DL / 2];                   // q_i pieces, packed bf16
#pragma unroll
  for (int qi = 0; qi < QI; ++qi) {
#pragma unroll
    for (int t = 0; t < DL / 2; ++t) qraw[qi][t] = 0u;
    if (i0 + qi < N) ld_bf16_raw<DL>(a.q + ((long)(i0 + qi) * B + b) * a.ldqk + dcol, qraw[qi]);
  }

  // ---- scores ----
  for (int bt = 0; bt < nb; ++bt) {
    const int j = bt * BK_KEYS + warp;
    uint32_t kraw[DL / 2];
#pragma unroll
    for (int t = 0; t < DL / 2; ++t) kraw[t] = 0u;
    if (j < N) ld_bf16_raw<DL>(a.k + ((long)j * B + b) * a.ldqk + dcol, kraw);
    if (bt + 1 < nb) {
      gather_issue_ca<QI>(a, gp, bt + 1, i0);       // next batch in flight while this one is consumed
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();                                // every thread's copies of batch bt have landed
    if (j < N) {
      const uint8_t* rows = gp.stage[bt & 1] + (size_t)(warp * QI) * gp.row_bytes + pcol * 2;
#pragma unroll
      for (int qi = 0; qi < QI; ++qi) {
        uint32_t ra[DL / 2], rb[DL / 2];
        lds_bf16_raw<DL>(rows + (size_t)qi * gp.row_bytes, ra);
        lds_bf16_raw<DL>(rows + (size_t)qi * gp.row_bytes + hd * 2, rb);
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < DL / 2; ++t) {
          acc = fmaf(bf_lo(qraw[qi][t]) + bf_lo(ra[t]), bf_lo(kraw[t]) + bf_lo(rb[t]), acc);
          acc = fmaf(bf_hi(qraw[qi][t]) + bf_hi(ra[t]), bf_hi(kraw[t]) + bf_hi(rb[t]), acc);
        }
        for (int o = lph >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (sub == 0) sc[(qi * H + h) * a.Npad + j] = (i0 + qi < N) ? acc * a.scale : 0.f;
      }
    }
    __syncthreads();                           // every warp is done with this stage before batch bt + 2 overwrites it
  }

  // ---- masks, softmax over keys, dropout (same counter-based draw as the attention core, so gtos_attn_bwd replays it) ----
  const unsigned long long seed = a.p_drop > 0.f ? reinterpret_cast<const unsigned long long*>(a.seed_ptr)[0] + a.seed_off : 0ull;
  const float ks = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  for (int r = warp; r < QI * H; r += 8) {
    const int qi = r / H, hh = r - qi * H, i = i0 + qi;
    float* w = sc + r * a.Npad;
    if (i >= N) {
      for (int j = lane; j < N; j += 32) w[j] = 0.f;
      continue;
    }
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      const bool masked = (a.key_pad && a.key_pad[(long)j * B + b]) || (a.attn_mask && a.attn_mask[(long)i * N + j]);
      const float v = masked ? -INFINITY : w[j];
      w[j] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = (w[j] == -INFINITY) ? 0.f : __expf(w[j] - mx);
      w[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    const long prow = (((long)b * H + hh) * N + i) * N;
    for (int j = lane; j < N; j += 32) {
      float p = w[j] * inv;
      a.probs[prow + j] = p;
      if (a.p_drop > 0.f) p = (rng_uniform(seed, (unsigned long long)(prow + j)) >= a.p_drop) ? p * ks : 0.f;
      if (a.probs_dropped) a.probs_dropped[prow + j] = p;
      w[j] = p;
    }
  }
  __syncthreads();

  // ---- PV: o_i = sum_j w_ij v_j.  Same split as the scores (warp = keys j = warp, warp + 8, ..; lane = DL features of its
  // head): every v row is read once per CTA with all of a warp's loads independent; the 8 partial sums meet in shared memory
  {
    float acc[QI][DL];
#pragma unroll
    for (int qi = 0; qi < QI; ++qi)
#pragma unroll
      for (int t = 0; t < DL; ++t) acc[qi][t] = 0.f;
#pragma unroll 2
    for (int j = warp; j < N; j += 8) {
      float vv[DL];
      const float4* vp = reinterpret_cast<const float4*>(a.v + ((long)j * B + b) * a.ldv + dcol);
#pragma unroll
      for (int t = 0; t < DL / 4; ++t) {
        const float4 f = vp[t];
        vv[4 * t] = f.x; vv[4 * t + 1] = f.y; vv[4 * t + 2] = f.z; vv[4 * t + 3] = f.w;
      }
#pragma unroll
      for (int qi = 0; qi < QI; ++qi) {
        const float p = sc[(qi * H + h) * a.Npad + j];
#pragma unroll
        for (int t = 0; t < DL; ++t) acc[qi][t] = fmaf(p, vv[t], acc[qi][t]);
      }
    }
#pragma unroll
    for (int qi = 0; qi < QI; ++qi) {
      float4* dst = reinterpret_cast<float4*>(part + ((long)warp * QI + qi) * a.D + dcol);
#pragma unroll
      for (int t = 0; t < DL / 4; ++t) dst[t] = make_float4(acc[qi][4 * t], acc[qi][4 * t + 1], acc[qi][4 * t + 2], acc[qi][4 * t + 3]);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < QI * (a.D / 2); e += 256) {
    const int qi = e / (a.D / 2), f = 2 * (e - qi * (a.D / 2));
    const int i = i0 + qi;
    if (i >= N) continue;
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) {
      const float2 t2 = *reinterpret_cast<const float2*>(part + ((long)w8 * QI + qi) * a.D + f);
      s2.x += t2.x; s2.y += t2.y;
    }
    const long o = ((long)i * B + b);
    *reinterpret_cast<float2*>(a.out + o * a.ldo + f) = s2;
    if (a.out_bf16) *reinterpret_cast<uint32_t*>(a.out_bf16 + o * a.D + f) = pack_bf16x2(s2.x, s2.y);
  }
}

template <int DL, int QI>
__global__ void __launch_bounds__(256) rel_grad_banked_ca_kernel(const RelBankedDev a) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ __align__(128) uint8_t smem_u8[];
  const int N = a.N, B = a.B, H = a.H, hd = a.hd;
  GatherPipe gp;
  gp.row_bytes = 4 * a.D;
  const size_t stage_bytes = (size_t)32 * gp.row_bytes;
  gp.stage[0] = smem_u8;
  gp.stage[1] = smem_u8 + stage_bytes;
  int* s_idx = reinterpret_cast<int*>(smem_u8 + 2 * stage_bytes);
  gp.s_idx = s_idx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * QI, b = blockIdx.y;
  const int lph = 32 / H;
  const int h = lane / lph, sub = lane - h * lph;
  const int dcol = h * hd + sub * DL;
  const int pcol = h * 2 * hd + sub * DL;
  const int nb = (N + BK_KEYS - 1) / BK_KEYS;
  for (int t = threadIdx.x; t < N * QI; t += 256) {
    const int j = t / QI, qi = t - j * QI;
    s_idx[t] = (i0 + qi < N) ? bank_row(a, j, i0 + qi, b) : 0;
  }
  __syncthreads();
  gather_issue_ca<QI>(a, gp, 0, i0);
  uint32_t qraw[QI][DL / 2];
#pragma unroll
  for (int qi = 0; qi < QI; ++qi) {
#pragma unroll
    for (int t = 0; t < DL / 2; ++t) qraw[qi][t] = 0u;
    if (i0 + qi < N) ld_bf16_raw<DL>(a.q + ((long)(i0 + qi) * B + b) * a.ldqk + dcol, qraw[qi]);
  }
  const RelTiling& rt = a.rt;
  for (int bt = 0; bt < nb; ++bt) {
    const int j = bt * BK_KEYS + warp;
    uint32_t kraw[DL / 2];
    float g[QI];
#pragma unroll
    for (int t = 0; t < DL / 2; ++t) kraw[t] = 0u;
#pragma unroll
    for (int qi = 0; qi < QI; ++qi) g[qi] = 0.f;
    if (j < N) {
      ld_bf16_raw<DL>(a.k + ((long)j * B + b) * a.ldqk + dcol, kraw);
      const float* dsp = a.dscores + (((long)b * H + h) * N + j) * N;
#pragma unroll
      for (int qi = 0; qi < QI; ++qi)
        if (i0 + qi < N) g[qi] = dsp[i0 + qi] * a.scale;
    }
    if (bt + 1 < nb) {
      gather_issue_ca<QI>(a, gp, bt + 1, i0);       // next batch in flight while this one is consumed
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();                                // every thread's copies of batch bt have landed
    if (j < N) {
      const uint8_t* rows = gp.stage[bt & 1] + (size_t)(warp * QI) * gp.row_bytes + pcol * 2;
      const int jb = j / rt.bj, jj = j - jb * rt.bj;
#pragma unroll
      for (int qi = 0; qi < QI; ++qi) {
        const int i = i0 + qi;
        if (i >= N) continue;                                      // warp-uniform
        uint32_t ra[DL / 2], rb[DL / 2];
        lds_bf16_raw<DL>(rows + (size_t)qi * gp.row_bytes, ra);
        lds_bf16_raw<DL>(rows + (size_t)qi * gp.row_bytes + hd * 2, rb);
        const int ib = i / rt.bi, ii = i - ib * rt.bi;
        const long row = (((long)b * rt.nj_blk + jb) * rt.ni_blk + ib) * 128 + jj * rt.bi + ii;
        __nv_bfloat16* gx = a.G + row * (2L * a.D) + pcol;         // d(q+ra) = g (k+rb) | d(k+rb) = g (q+ra)
        uint32_t wx[DL / 2], wy[DL / 2];
#pragma unroll
        for (int t = 0; t < DL / 2; ++t) {
          wx[t] = pack_bf16x2(g[qi] * (bf_lo(kraw[t]) + bf_lo(rb[t])), g[qi] * (bf_hi(kraw[t]) + bf_hi(rb[t])));
          wy[t] = pack_bf16x2(g[qi] * (bf_lo(qraw[qi][t]) + bf_lo(ra[t])), g[qi] * (bf_hi(qraw[qi][t]) + bf_hi(ra[t])));
        }
        if constexpr (DL == 4) {
          *reinterpret_cast<uint2*>(gx) = make_uint2(wx[0], wx[1]);
          *reinterpret_cast<uint2*>(gx + hd) = make_uint2(wy[0], wy[1]);
        } else {
#pragma unroll
          for (int t = 0; t < DL / 8; ++t) {
            *reinterpret_cast<uint4*>(gx + 8 * t) = make_uint4(wx[4 * t], wx[4 * t + 1], wx[4 * t + 2], wx[4 * t + 3]);
            *reinterpret_cast<uint4*>(gx + hd + 8 * t) = make_uint4(wy[4 * t], wy[4 * t + 1], wy[4 * t + 2], wy[4 * t + 3]);
          }
        }
      }
    }
    __syncthreads();
  }
}

static int banked_check(const RelBankedArgs& a, RelBankedDev* d) {
  GTOS_REQUIRE(a.N > 0 && a.B > 0 && a.H > 0 && a.D % a.H == 0, "rel_banked: bad shape N=%d B=%d D=%d H=%d", a.N, a.B, a.D, a.H);
  const int hd = a.D / a.H;
  GTOS_REQUIRE(a.D % 32 == 0 && 32 % a.H == 0 && (a.D / 32 == 4 || a.D / 32 == 8 || a.D / 32 == 16 || a.D / 32 == 32),
               "rel_banked: needs D in {128,256,512,1024} and H dividing 32 (got D=%d H=%d)", a.D, a.H);
  GTOS_REQUIRE(a.ldpb % 8 == 0 && a.ldqk % 8 == 0 && (reinterpret_cast<uintptr_t>(a.PB) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(a.q) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.k) & 15) == 0,
               "rel_banked: PB / q / k rows must be 16-byte aligned");
  memset(d, 0, sizeof(*d));
  d->PB = reinterpret_cast<const __nv_bfloat16*>(a.PB); d->ldpb = a.ldpb;
  d->idx = a.idx;
  d->q = reinterpret_cast<const __nv_bfloat16*>(a.q); d->k = reinterpret_cast<const __nv_bfloat16*>(a.k); d->ldqk = a.ldqk;
  d->v = a.v; d->ldv = a.ldv;
  d->key_pad = a.key_pad; d->attn_mask = a.attn_mask;
  d->p_drop = a.p_drop; d->seed_ptr = a.seed_ptr; d->seed_off = a.seed_off;
  d->probs = a.probs; d->probs_dropped = a.probs_dropped;
  d->out = a.out; d->ldo = a.ldo; d->out_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  d->dscores = a.dscores; d->G = reinterpret_cast<__nv_bfloat16*>(a.G);
  d->N = a.N; d->B = a.B; d->D = a.D; d->H = a.H; d->hd = hd; d->R = a.R;
  d->Npad = (a.N + 3) & ~3;
  d->scale = 1.0f / sqrtf((float)hd);
  return GTOS_OK;
}

static constexpr int BANKED_QI = 4;

template <int DL>
static int launch_banked_fwd(const RelBankedDev& d, cudaStream_t st) {
  constexpr int QI = BANKED_QI;
  const size_t smem = sizeof(float) * ((size_t)QI * d.H * d.Npad + (size_t)8 * QI * d.D);
  GTOS_REQUIRE(smem <= 200 * 1024, "rel_attn_banked_fwd: N=%d needs %zu bytes of shared memory", d.N, smem);
  static const bool ca = getenv("GTOS_BANKED_CPASYNC") && getenv("GTOS_BANKED_CPASYNC")[0] == '1';   // measured slower (133 / 163 us vs 84 / 85 us): off
  if (ca) {
    const size_t sm2 = (size_t)2 * 32 * 4 * d.D + sizeof(float) * (size_t)QI * d.H * d.Npad + sizeof(int) * (size_t)d.N * QI;
    if (sm2 <= 220 * 1024) {
      auto k2 = rel_attn_banked_fwd_ca_kernel<DL, QI>;
      GTOS_CHECK_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      dim3 g2((d.N + QI - 1) / QI, d.B);
      GTOS_KLAUNCH(k2, g2, dim3(256), sm2, st, d);
      GTOS_LAUNCH_CHECK();
      return GTOS_OK;
    }
  }
  auto kern = rel_attn_banked_fwd_kernel<DL, QI>;
  GTOS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((d.N + QI - 1) / QI, d.B);
  GTOS_KLAUNCH(kern, grid, dim3(256), smem, st, d);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int rel_attn_banked_fwd(const RelBankedArgs& a, cudaStream_t st) {
  RelBankedDev d;
  int e = banked_check(a, &d);
  if (e) return e;
  GTOS_REQUIRE(a.v && a.probs && a.out && a.ldv % 2 == 0 && a.ldo % 2 == 0, "rel_attn_banked_fwd: v / probs / out are required");
  GTOS_REQUIRE(a.p_drop == 0.f || a.seed_ptr, "rel_attn_banked_fwd: dropout needs a device seed pointer");
  switch (a.D / 32) {
    case 4: return launch_banked_fwd<4>(d, st);
    case 8: return launch_banked_fwd<8>(d, st);
    case 16: return launch_banked_fwd<16>(d, st);
    default: return launch_banked_fwd<32>(d, st);
  }
}

template <int DL>
static int launch_banked_grad(const RelBankedDev& d, cudaStream_t st) {
  constexpr int QI = BANKED_QI;
  const size_t smem = sizeof(int) * d.N * QI;
  static const bool ca = getenv("GTOS_BANKED_CPASYNC") && getenv("GTOS_BANKED_CPASYNC")[0] == '1';   // measured slower (133 / 163 us vs 84 / 85 us): off
  if (ca) {
    const size_t sm2 = (size_t)2 * 32 * 4 * d.D + sizeof(int) * (size_t)d.N * QI;
    if (sm2 <= 220 * 1024) {
      auto k2 = rel_grad_banked_ca_kernel<DL, QI>;
      GTOS_CHECK_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      dim3 g2((d.N + QI - 1) / QI, d.B);
      GTOS_KLAUNCH(k2, g2, dim3(256), sm2, st, d);
      GTOS_LAUNCH_CHECK();
      return GTOS_OK;
    }
  }
  auto kern = rel_grad_banked_kernel<DL, QI>;
  GTOS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((d.N + QI - 1) / QI, d.B);
  GTOS_KLAUNCH(kern, grid, dim3(256), smem, st, d);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int rel_grad_banked(const RelBankedArgs& a, cudaStream_t st) {
  RelBankedDev d;
  int e = banked_check(a, &d);
  if (e) return e;
  GTOS_REQUIRE(a.dscores && a.G, "rel_grad_banked: dscores and G are required");
  e = choose_rel_tiling(&d.rt, a.N, a.B, a.D, a.H);
  if (e) return e;
  switch (a.D / 32) {
    case 4: return launch_banked_grad<4>(d, st);
    case 8: return launch_banked_grad<8>(d, st);
    case 16: return launch_banked_grad<16>(d, st);
    default: return launch_banked_grad<32>(d, st);
  }
}

}  // namespace gtos
