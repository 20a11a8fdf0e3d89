// gtos_b200 -- kernels of the widened rows of SURVEY.md §8(f): incremental beam decode (f-1), the work=True
// log-probability table (f-2) and the fused Adam / global-norm clip over flat buffers (f-4).
// All of them are HBM/L2-bound streaming kernels: no tensor-core shapes here (T_q = 1, or one pass over a vector).
#include <math.h>
#include <stdlib.h>

#include "elementwise.cuh"

namespace gtos {

// ---------------------------------------------------------------------------------------
// f-1: single-query attention over a K/V cache.
// The reference's decode step (generator/generator.py:120-167) calls MultiheadAttention (transformer.py:98-173) with
// T_q = 1 and re-projects every key/value row of the graph memory and of the token prefix for every live hypothesis at
// every step.  Here keys/values are projected ONCE into a bf16 cache; a hypothesis reads its rows through an index:
//   cross-attention: cache row (l, source graph of h)          -> slot = src_index[h], slot_ld = 0
//   self-attention : cache row (l, ancestor of h at position l) -> slot = anc[l][h],    slot_ld = row pitch of anc
// so a beam reorder copies no K/V bytes (the reference index_selects every state tensor, search.py:72-76).
// One warp per (hypothesis, head): lanes own keys for q.k, softmax by shuffles, lanes own feature pairs for P.V.
// ---------------------------------------------------------------------------------------
struct DecodeAttn {
  int Hyp, L, H, hd, lpr;
  const float* q; long ldq;
  const __nv_bfloat16* kv; long ld_kv; int v_off; long row_stride;
  const int* slot; long slot_ld;
  const unsigned char* key_pad; long pad_ld;
  float scale;
  float* out; long ldo;
  __nv_bfloat16* out_bf16; long ldob;
  float* probs;
};

__device__ __forceinline__ float dot8_bf16(const float* q, uint4 u) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    s = fmaf(q[2 * i], f.x, s);
    s = fmaf(q[2 * i + 1], f.y, s);
  }
  return s;
}

// grid = (ceil(Hyp / warps per block), H): the warps of a block are CONSECUTIVE hypotheses of one head, i.e. (in a beam
// search) siblings that read the same graph rows / mostly the same ancestor rows -> one L2 read, L1 hits for the rest.
// The kernel is latency-bound (a warp owns ~40 rows), so every phase keeps several independent loads in flight:
//   0. index phase: cache row + mask of every key position -> shared memory (lanes own positions);
//   1. q.k: `lpr` lanes share a key row (16 bytes each: whole 128-byte lines), 32/lpr rows per pass, 4 passes unrolled;
//   2. softmax over the warp's positions (shuffles);
//   3. P.V: a lane owns EPL consecutive features (one 4/8/16-byte load per row), 4 rows unrolled.
template <int EPL>
__global__ void attn_decode_kernel(const DecodeAttn a) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ float sm[];
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hyp = blockIdx.x * wpb + w, head = blockIdx.y;
  if (hyp >= a.Hyp) return;                                    // warps are independent: only __syncwarp below
  const int hd = a.hd, L = a.L, chunks = hd >> 3;
  float* sc = sm + (size_t)w * (2 * L);
  uint32_t* rows = reinterpret_cast<uint32_t*>(sc + L);        // cache-row offsets in 16-byte units (32-bit index math below)
  // ---- 0. rows / masks ----
  const long ld16 = a.ld_kv >> 3;
  for (int l = lane; l < L; l += 32) {
    const int s = a.slot ? __ldg(a.slot + (long)l * a.slot_ld + hyp) : hyp;
    const bool masked = a.key_pad && a.key_pad[(long)l * a.pad_ld + s];
    rows[l] = (uint32_t)(((long)l * a.row_stride + s) * ld16);
    sc[l] = masked ? -INFINITY : 0.f;
  }
  const int lpr = a.lpr, rpi = 32 / lpr, sub = lane & (lpr - 1), rsel = lane / lpr;
  const bool has_a = sub < chunks, has_b = sub + 32 < chunks;
  float qa[8], qb[8];
  {
    const float* q = a.q + (long)hyp * a.ldq + head * hd;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      qa[i] = has_a ? q[sub * 8 + i] * a.scale : 0.f;
      qb[i] = has_b ? q[(sub + 32) * 8 + i] * a.scale : 0.f;
    }
  }
  __syncwarp();
  // ---- 1. scores ----
  const uint4* kbase = reinterpret_cast<const uint4*>(a.kv + head * hd);
  for (int l0 = 0; l0 < L; l0 += 4 * rpi) {
    uint4 ka[4], kb[4];
    int ll[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ll[u] = l0 + u * rpi + rsel;
      ka[u] = make_uint4(0, 0, 0, 0);
      kb[u] = make_uint4(0, 0, 0, 0);
      if (ll[u] < L) {
        const uint4* kr = kbase + rows[ll[u]];
        if (has_a) ka[u] = __ldg(kr + sub);
        if (has_b) kb[u] = __ldg(kr + sub + 32);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float part = dot8_bf16(qa, ka[u]);
      if (has_b) part += dot8_bf16(qb, kb[u]);
      for (int o = lpr >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (ll[u] < L && sub == 0) sc[ll[u]] += part;           // -inf + x stays -inf for masked keys
    }
  }
  __syncwarp();
  // ---- 2. softmax ----
  float mx = -INFINITY;
  for (int l = lane; l < L; l += 32) mx = fmaxf(mx, sc[l]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int l = lane; l < L; l += 32) {
    const float e = (mx == -INFINITY) ? 0.f : __expf(sc[l] - mx);
    sc[l] = e;
    se += e;
  }
  se = warp_sum(se);
  const float inv = se > 0.f ? 1.f / se : 0.f;
  for (int l = lane; l < L; l += 32) {
    const float p = sc[l] * inv;
    sc[l] = p;
    if (a.probs) a.probs[((long)hyp * a.H + head) * L + l] = p;
  }
  __syncwarp();
  // ---- 3. P.V ----
  const int e0 = lane * EPL;
  if (e0 < hd) {
    const char* vbase = reinterpret_cast<const char*>(a.kv + a.v_off + head * hd + e0);
    float acc[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) acc[i] = 0.f;
    constexpr int NW = EPL / 2;                                 // 32-bit words per lane per row
    for (int l0 = 0; l0 < L; l0 += 4) {
      uint32_t wv[4][NW];
      float pw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int l = l0 + u;
        pw[u] = l < L ? sc[l] : 0.f;
        const char* vp = vbase + (size_t)rows[l < L ? l : 0] * 16u;
        if constexpr (NW == 1) {
          wv[u][0] = __ldg(reinterpret_cast<const uint32_t*>(vp));
        } else if constexpr (NW == 2) {
          const uint2 t2 = __ldg(reinterpret_cast<const uint2*>(vp));
          wv[u][0] = t2.x; wv[u][1] = t2.y;
        } else {
#pragma unroll
          for (int c4 = 0; c4 < NW / 4; ++c4) {
            const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(vp) + c4);
            wv[u][4 * c4] = t4.x; wv[u][4 * c4 + 1] = t4.y; wv[u][4 * c4 + 2] = t4.z; wv[u][4 * c4 + 3] = t4.w;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv[u][i]));
          acc[2 * i] = fmaf(pw[u], f.x, acc[2 * i]);
          acc[2 * i + 1] = fmaf(pw[u], f.y, acc[2 * i + 1]);
        }
    }
    if (a.out) {
      float* o = a.out + (long)hyp * a.ldo + head * hd + e0;
#pragma unroll
      for (int i = 0; i < EPL; i += 2) *reinterpret_cast<float2*>(o + i) = make_float2(acc[i], acc[i + 1]);
    }
    if (a.out_bf16) {
      __nv_bfloat16* o = a.out_bf16 + (long)hyp * a.ldob + head * hd + e0;
#pragma unroll
      for (int i = 0; i < EPL; i += 2) *reinterpret_cast<uint32_t*>(o + i) = pack_bf16x2(acc[i], acc[i + 1]);
    }
  }
}

// ---- row-wide variant: ONE warp per hypothesis handles ALL heads --------------------------------------------------
// A lane owns DPL = D/32 consecutive features (a fraction of one head: hd/DPL lanes per head), so a key / value row of
// the cache (all heads, D bf16) is read by the warp as one contiguous 2*D-byte piece and the per-(head, key) instruction
// cost drops ~4x against the warp-per-(hypothesis, head) kernel above (which was issue-bound: 60 instructions per key
// for 64 features).  Used when D is 64/128/256/512; 8 keys are in flight per lane to cover the load latency.
template <int DPL>
__device__ __forceinline__ void ld_chunk(const char* p, uint32_t* w) {
  if constexpr (DPL == 2) {
    w[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
  } else if constexpr (DPL == 4) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    w[0] = t.x; w[1] = t.y;
  } else {
#pragma unroll
    for (int c = 0; c < DPL / 8; ++c) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(p) + c);
      w[4 * c] = t.x; w[4 * c + 1] = t.y; w[4 * c + 2] = t.z; w[4 * c + 3] = t.w;
    }
  }
}

template <int DPL>
__global__ void attn_decode_wide_kernel(const DecodeAttn a) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ float sm[];
  constexpr int NW = DPL / 2, U = 8;
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hyp = blockIdx.x * wpb + w;
  if (hyp >= a.Hyp) return;
  const int hd = a.hd, L = a.L, H = a.H;
  const int per_warp = H * L + L + ((L + 3) >> 2);
  float* sc = sm + (size_t)w * per_warp;                       // [H][L]
  uint32_t* rows = reinterpret_cast<uint32_t*>(sc + H * L);    // [L] cache-row offsets, 16-byte units
  unsigned char* msk = reinterpret_cast<unsigned char*>(rows + L);
  const long ld16 = a.ld_kv >> 3;
  for (int l = lane; l < L; l += 32) {
    const int s = a.slot ? __ldg(a.slot + (long)l * a.slot_ld + hyp) : hyp;
    msk[l] = (a.key_pad && a.key_pad[(long)l * a.pad_ld + s]) ? 1 : 0;
    rows[l] = (uint32_t)(((long)l * a.row_stride + s) * ld16);
  }
  const int d0 = lane * DPL, head = d0 / hd, lph = hd / DPL, sub = lane & (lph - 1);
  float qv[DPL];
  {
    const float* q = a.q + (long)hyp * a.ldq + d0;
#pragma unroll
    for (int i = 0; i < DPL; ++i) qv[i] = q[i] * a.scale;
  }
  __syncwarp();
  float* sch = sc + head * L;
  // ---- scores ----
  const char* kbase = reinterpret_cast<const char*>(a.kv + d0);
  for (int l0 = 0; l0 < L; l0 += U) {
    uint32_t kw[U][NW];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int l = l0 + u < L ? l0 + u : L - 1;
      ld_chunk<DPL>(kbase + (size_t)rows[l] * 16u, kw[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&kw[u][i]));
        part = fmaf(qv[2 * i], f.x, part);
        part = fmaf(qv[2 * i + 1], f.y, part);
      }
      for (int o = lph >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      const int l = l0 + u;
      if (l < L && sub == 0) sch[l] = msk[l] ? -INFINITY : part;
    }
  }
  __syncwarp();
  // ---- softmax over the keys of this lane's head (its lph lanes share the work) ----
  float mx = -INFINITY;
  for (int l = sub; l < L; l += lph) mx = fmaxf(mx, sch[l]);
  for (int o = lph >> 1; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f;
  for (int l = sub; l < L; l += lph) {
    const float e = (mx == -INFINITY) ? 0.f : __expf(sch[l] - mx);
    sch[l] = e;
    se += e;
  }
  for (int o = lph >> 1; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  const float inv = se > 0.f ? 1.f / se : 0.f;
  for (int l = sub; l < L; l += lph) {
    const float pr = sch[l] * inv;
    sch[l] = pr;
    if (a.probs) a.probs[((long)hyp * H + head) * L + l] = pr;
  }
  __syncwarp();
  // ---- P.V ----
  const char* vbase = reinterpret_cast<const char*>(a.kv + a.v_off + d0);
  float acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
  for (int l0 = 0; l0 < L; l0 += U) {
    uint32_t vw[U][NW];
    float pw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int l = l0 + u;
      pw[u] = l < L ? sch[l] : 0.f;
      ld_chunk<DPL>(vbase + (size_t)rows[l < L ? l : L - 1] * 16u, vw[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vw[u][i]));
        acc[2 * i] = fmaf(pw[u], f.x, acc[2 * i]);
        acc[2 * i + 1] = fmaf(pw[u], f.y, acc[2 * i + 1]);
      }
  }
  if (a.out) {
    float* o = a.out + (long)hyp * a.ldo + d0;
#pragma unroll
    for (int i = 0; i < DPL; i += 2) *reinterpret_cast<float2*>(o + i) = make_float2(acc[i], acc[i + 1]);
  }
  if (a.out_bf16) {
    __nv_bfloat16* o = a.out_bf16 + (long)hyp * a.ldob + d0;
#pragma unroll
    for (int i = 0; i < DPL; i += 2) *reinterpret_cast<uint32_t*>(o + i) = pack_bf16x2(acc[i], acc[i + 1]);
  }
}

template <int DPL>
static int launch_attn_decode_wide(const DecodeAttn& a, cudaStream_t st) {
  const size_t per_warp = (size_t)(a.H * a.L + a.L + ((a.L + 3) >> 2)) * sizeof(float);
  int wpb = 4;
  while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
  if (per_warp * wpb > 200 * 1024) return -1;                  // caller falls back to the per-head kernel
  const size_t smem = per_warp * wpb;
  auto kern = attn_decode_wide_kernel<DPL>;
  if (smem > 48 * 1024) GTOS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  GTOS_KLAUNCH(kern, dim3((unsigned)((a.Hyp + wpb - 1) / wpb)), dim3(wpb * 32), smem, st, a);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int attn_decode(int Hyp, int L, int H, int hd, const float* q, long ldq, const void* kv, long ld_kv, int v_off,
                long row_stride, const int* slot, long slot_ld, const unsigned char* key_pad, long pad_ld, float scale,
                float* out, long ldo, void* out_bf16, long ldob, float* probs, cudaStream_t st) {
  if (Hyp == 0 || H == 0) return GTOS_OK;
  GTOS_REQUIRE(L > 0 && hd > 0 && hd % 8 == 0 && hd <= 512, "attn_decode: need L > 0, head_dim %% 8 == 0, head_dim <= 512 (L=%d, hd=%d)", L, hd);
  GTOS_REQUIRE(ld_kv % 8 == 0 && v_off % 8 == 0, "attn_decode: cache row stride / value offset must be multiples of 8");
  GTOS_REQUIRE(((long)L * row_stride + Hyp) * (ld_kv / 8) < (1ll << 32) || slot, "attn_decode: cache larger than 64 GB");
  GTOS_REQUIRE(q && kv && (out || out_bf16), "attn_decode: null argument");
  GTOS_REQUIRE(ldo % 2 == 0 && ldob % 2 == 0 && H <= 65535, "attn_decode: output strides must be even, H <= 65535");
  const int chunks = hd / 8;
  int lpr = 1;
  while (lpr < chunks && lpr < 32) lpr <<= 1;                    // lanes per key row (a lane handles chunks sub, sub + 32)
  const size_t per_warp = (size_t)(2 * L) * sizeof(float);
  int wpb = 8;
  while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
  GTOS_REQUIRE(per_warp * wpb <= 200 * 1024, "attn_decode: L=%d keys do not fit in shared memory", L);
  const size_t smem = per_warp * wpb;
  // features per lane in the P.V phase: hd/32 rounded up to a power of two, at least one bf16 pair
  int epl = 2;
  while (epl * 32 < hd) epl <<= 1;
  GTOS_REQUIRE(hd % epl == 0 && (ld_kv * 2) % (epl * 2 < 16 ? epl * 2 : 16) == 0 && (v_off * 2) % (epl * 2 < 16 ? epl * 2 : 16) == 0,
               "attn_decode: head_dim %d / cache layout not aligned for %d features per lane", hd, epl);
  void (*kern)(const DecodeAttn) = epl == 2 ? attn_decode_kernel<2> : epl == 4 ? attn_decode_kernel<4>
                                 : epl == 8 ? attn_decode_kernel<8> : attn_decode_kernel<16>;
  if (smem > 48 * 1024) GTOS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  DecodeAttn a;
  a.Hyp = Hyp; a.L = L; a.H = H; a.hd = hd; a.lpr = lpr; a.q = q; a.ldq = ldq;
  a.kv = reinterpret_cast<const __nv_bfloat16*>(kv); a.ld_kv = ld_kv; a.v_off = v_off; a.row_stride = row_stride;
  a.slot = slot; a.slot_ld = slot_ld; a.key_pad = key_pad; a.pad_ld = pad_ld; a.scale = scale;
  a.out = out; a.ldo = ldo; a.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); a.ldob = ldob; a.probs = probs;
  {
    // one warp per hypothesis for all heads when a lane can own D/32 features of ONE head
    static const bool wide_ok = !(getenv("GTOS_ATTN_WIDE") && getenv("GTOS_ATTN_WIDE")[0] == '0');
    const int D = H * hd, dpl = D / 32;
    if (wide_ok && D % 32 == 0 && (dpl == 2 || dpl == 4 || dpl == 8 || dpl == 16) && hd % dpl == 0 && ((hd / dpl) & (hd / dpl - 1)) == 0 &&
        hd / dpl <= 32 && (ld_kv * 2) % (dpl * 2 < 16 ? dpl * 2 : 16) == 0 && ldq >= D) {
      int rc = dpl == 2 ? launch_attn_decode_wide<2>(a, st) : dpl == 4 ? launch_attn_decode_wide<4>(a, st)
             : dpl == 8 ? launch_attn_decode_wide<8>(a, st) : launch_attn_decode_wide<16>(a, st);
      if (rc >= 0) return rc;
    }
  }
  GTOS_KLAUNCH(kern, dim3(dim3((unsigned)((Hyp + wpb - 1) / wpb), (unsigned)H)), dim3(wpb * 32), smem, st, a);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// f-2 (work=True half): log-probability table over the batch-extended vocabulary (generator/decoder.py:42-59):
//   table[row, v] = log( gen * softmax(logits)[v] (v < V)  +  cpy * sum_s align[row,s] [copy_seq[s, b(row)] == v]  + 1e-12 )
// replaces softmax / zero-extension cat / scatter_add_ / log (5 passes over [rows, V] and a .item() sync).
// One CTA per row; the row stays in L2 between the three phases.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float blk_reduce(float v, float* sh, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float r = (lane < nw) ? sh[lane] : (is_max ? -INFINITY : 0.f);
  return is_max ? warp_max(r) : warp_sum(r);
}

__global__ void token_logprob_kernel(const float* __restrict__ logits, long ldl, int V, const float* __restrict__ gate_logits,
                                     const float* __restrict__ align, int S, const long long* __restrict__ copy_seq,
                                     int Bsrc, const int* __restrict__ src_index, int B, float* __restrict__ table, long ldt,
                                     int W) {
  GTOS_PDL_PROLOGUE();
  __shared__ float sh[32];
  const long row = blockIdx.x;
  const int b = src_index ? src_index[row] : (int)(row % B);
  const float* lr = logits + row * ldl;
  float* tr = table + row * ldt;
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, lr[v]);
  mx = blk_reduce(mx, sh, true);
  float se = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) se += __expf(lr[v] - mx);
  se = blk_reduce(se, sh, false);
  const float g0 = gate_logits[row * 2], g1 = gate_logits[row * 2 + 1];
  const float gm = fmaxf(g0, g1);
  const float e0 = __expf(g0 - gm), e1 = __expf(g1 - gm);
  const float gen = e0 / (e0 + e1), cpy = e1 / (e0 + e1);
  const float coef = gen / se;
  for (int v = threadIdx.x; v < W; v += blockDim.x) tr[v] = v < V ? coef * __expf(lr[v] - mx) : 0.f;
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const long long slot = copy_seq[(long)s * Bsrc + b];
    if (slot >= 0 && slot < W) atomicAdd(tr + slot, cpy * align[row * S + s]);
  }
  __syncthreads();
  for (int v = threadIdx.x; v < W; v += blockDim.x) tr[v] = logf(tr[v] + 1e-12f);
}

int token_logprob(const float* logits, long ldl, int V, const float* gate_logits, const float* align, int S,
                  const long long* copy_seq, int Bsrc, const int* src_index, long rows, int B, float* table, long ldt, int W,
                  cudaStream_t st) {
  if (rows == 0) return GTOS_OK;
  GTOS_REQUIRE(W >= V && ldt >= W && Bsrc > 0 && B > 0, "token_logprob: need W >= V, ldt >= W (V=%d, W=%d, ldt=%ld)", V, W, ldt);
  GTOS_KLAUNCH(token_logprob_kernel, dim3((unsigned)rows), dim3(256), 0, st, logits, ldl, V, gate_logits, align, S, copy_seq, Bsrc, src_index, B,
                                                       table, ldt, W);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// f-1/f-2: the same row computed in SHARED memory with the top-k taken in the same kernel: the beam step only needs the
// k best tokens of every hypothesis (generator.py:157 torch.topk), so the [Hyp, W] table never goes to HBM
// (torch.topk alone cost 320 us per step at Hyp = 2048, W = 10016: 4 kernels re-reading the table).
// Ranking is done on the probabilities (log is monotonic); ties -> lowest token id.  One CTA per row.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long blk_max_u64(unsigned long long v, unsigned long long* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long x = __shfl_xor_sync(0xffffffffu, v, o);
    v = x > v ? x : v;
  }
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  unsigned long long r = lane < nw ? sh[lane] : 0ull;
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long x = __shfl_xor_sync(0xffffffffu, r, o);
    r = x > r ? x : r;
  }
  return r;
}

__global__ void token_topk_kernel(const float* __restrict__ logits, long ldl, int V, const float* __restrict__ gate_logits,
                                  const float* __restrict__ align, int S, const long long* __restrict__ copy_seq, int Bsrc,
                                  const int* __restrict__ src_index, int B, int W, int K, float* __restrict__ top_val,
                                  int* __restrict__ top_idx, float* __restrict__ table, long ldt) {
  GTOS_PDL_PROLOGUE();
  extern __shared__ float prow[];                     // W probabilities
  __shared__ float sh[32];
  __shared__ unsigned long long shu[32];
  const long row = blockIdx.x;
  const int b = src_index ? src_index[row] : (int)(row % B);
  const float* lr = logits + row * ldl;
  float mx = -INFINITY;
  for (int v0 = threadIdx.x; v0 < V; v0 += 8 * blockDim.x) {    // 8 independent loads in flight per thread
    float x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int v = v0 + u * blockDim.x;
      x[u] = v < V ? __ldg(lr + v) : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int v = v0 + u * blockDim.x;
      if (v < V) prow[v] = x[u];
      mx = fmaxf(mx, x[u]);
    }
  }
  mx = blk_reduce(mx, sh, true);
  float se = 0.f;
#pragma unroll 4
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float e = __expf(prow[v] - mx);
    prow[v] = e;
    se += e;
  }
  se = blk_reduce(se, sh, false);
  const float g0 = gate_logits[row * 2], g1 = gate_logits[row * 2 + 1];
  const float gm = fmaxf(g0, g1);
  const float e0 = __expf(g0 - gm), e1 = __expf(g1 - gm);
  const float gen = e0 / (e0 + e1), cpy = e1 / (e0 + e1);
  const float coef = gen / se;
  for (int v = threadIdx.x; v < W; v += blockDim.x) prow[v] = v < V ? coef * prow[v] : 0.f;
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const long long slot = copy_seq[(long)s * Bsrc + b];
    if (slot >= 0 && slot < W) atomicAdd(prow + slot, cpy * align[row * S + s]);
  }
  __syncthreads();
  if (table)
    for (int v = threadIdx.x; v < W; v += blockDim.x) table[row * ldt + v] = logf(prow[v] + 1e-12f);
  // K rounds of block arg-max; key = (probability bits, ~index) so the max picks the lowest index among equals.
  // Every thread caches the best TWO of its own elements and rescans only after both were taken (the rescan of one
  // thread is what the other 255 wait for at the next barrier).
  unsigned long long m0 = 0ull, m1 = 0ull;
  auto rescan = [&]() {
    m0 = 0ull; m1 = 0ull;
#pragma unroll 4
    for (int v = threadIdx.x; v < W; v += blockDim.x) {
      const float pv = prow[v];
      if (pv >= 0.f) {                                 // taken elements are marked -1
        const unsigned long long key = ((unsigned long long)__float_as_uint(pv) << 32) | (unsigned)(0xffffffffu - (unsigned)v);
        if (key > m0) { m1 = m0; m0 = key; } else if (key > m1) { m1 = key; }
      }
    }
  };
  rescan();
  bool have1 = true;
  for (int k = 0; k < K; ++k) {
    const unsigned long long win = blk_max_u64(m0, shu);
    const int idx = (int)(0xffffffffu - (unsigned)(win & 0xffffffffull));
    if (threadIdx.x == 0) {
      top_val[row * K + k] = logf(__uint_as_float((unsigned)(win >> 32)) + 1e-12f);
      top_idx[row * K + k] = idx;
    }
    if (win != 0ull && idx % (int)blockDim.x == (int)threadIdx.x) {
      prow[idx] = -1.f;
      if (have1) { m0 = m1; m1 = 0ull; have1 = false; } else { rescan(); have1 = true; }
    }
  }
}

int token_topk(const float* logits, long ldl, int V, const float* gate_logits, const float* align, int S,
               const long long* copy_seq, int Bsrc, const int* src_index, long rows, int B, int W, int K, float* top_val,
               int* top_idx, float* table, long ldt, cudaStream_t st) {
  if (rows == 0) return GTOS_OK;
  GTOS_REQUIRE(W >= V && K >= 1 && K <= W && Bsrc > 0 && B > 0 && top_val && top_idx, "token_topk: bad arguments (V=%d W=%d K=%d)", V, W, K);
  GTOS_REQUIRE(!table || ldt >= W, "token_topk: table row pitch %ld < W=%d", ldt, W);
  const size_t smem = (size_t)W * sizeof(float);
  GTOS_REQUIRE(smem <= 200 * 1024, "token_topk: vocabulary row of %d entries does not fit in shared memory", W);
  if (smem > 40 * 1024)
    GTOS_CHECK_CUDA(cudaFuncSetAttribute(token_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  GTOS_KLAUNCH(token_topk_kernel, dim3((unsigned)rows), dim3(256), smem, st, logits, ldl, V, gate_logits, align, S, copy_seq, Bsrc, src_index, B, W, K,
                                                       top_val, top_idx, table, ldt);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// f-1: one `Beam.update` (generator/search.py:57-92) for every beam of the batch in ONE launch; one CTA per source graph,
// one thread per (live slot, rank) candidate.  Same semantics as gtos_b200.decode.BeamState.update (the portable
// implementation the CPU tests pin to golden runs of the reference's Beam class); ~40 small PyTorch kernels otherwise.
// ---------------------------------------------------------------------------------------
struct BeamUpd {
  int B, K, t, Tmin, Tmax, end_id, unk_id;
  const float* top_val; const int* top_idx;          // [B*K, K]
  float* score; unsigned char* live; int* n_done; int* steps;      // [B,K], [B,K], [B], [B]
  int* tok; int* par;                                // [Tmax, B, K]
  float* done_score; int* done_step; int* done_par;  // [B,K]
  int* parent_out; long long* last_tok;              // [B*K]
};

__global__ void beam_update_kernel(const BeamUpd a) {
  GTOS_PDL_PROLOGUE();
  __shared__ float s_val[256];
  __shared__ unsigned char s_valid[256];
  __shared__ int s_rank_c[16];       // candidate id holding rank r (r < K)
  __shared__ float s_ns[16];
  __shared__ int s_nt[16], s_np[16];
  __shared__ int s_counts[2];
  const int b = blockIdx.x, K = a.K, KK = K * K, c = threadIdx.x;
  const int nd = a.n_done[b], st = a.steps[b];
  const bool active = nd < K && st < a.Tmax;
  int* tok_t = a.tok + ((long)a.t * a.B + b) * K;
  int* par_t = a.par + ((long)a.t * a.B + b) * K;
  if (!active) {
    if (c < K) {
      tok_t[c] = 0;
      par_t[c] = c;
      a.parent_out[b * K + c] = b * K + c;
      a.last_tok[b * K + c] = 0;
    }
    return;
  }
  const int i = c / K;
  bool valid = false;
  float val = -INFINITY;
  int tokc = 0;
  if (c < KK) {
    valid = a.live[b * K + i] != 0;
    tokc = a.top_idx[(long)(b * K + i) * K + (c - i * K)];
    val = tokc == a.unk_id ? -INFINITY : a.score[b * K + i] + a.top_val[(long)(b * K + i) * K + (c - i * K)];
    s_val[c] = val;
    s_valid[c] = valid;
  }
  if (c < 16) { s_rank_c[c] = -1; s_ns[c] = 0.f; s_nt[c] = 0; s_np[c] = 0; }
  __syncthreads();
  int rank = 0, n_real = 0;
  if (c < KK) {
    for (int o = 0; o < KK; ++o) {
      if (!s_valid[o]) continue;
      ++n_real;
      const float vo = s_val[o];
      if (valid && (vo > val || (vo == val && o < c))) ++rank;
    }
  }
  const int n_take = min(K - nd, n_real);
  const bool taken = c < KK && valid && rank < n_take;
  if (taken) s_rank_c[rank] = c;
  __syncthreads();
  // walk the taken candidates in rank order (<= K of them): completed / alive bookkeeping
  if (c == 0) {
    int n_c = 0, n_l = 0;
    for (int r = 0; r < n_take; ++r) {
      const int cc = s_rank_c[r];
      const int ii = cc / K;
      const int tk = a.top_idx[(long)(b * K + ii) * K + (cc - ii * K)];
      const float v = s_val[cc];
      if (tk == a.end_id) {
        if (a.t >= a.Tmin) {                                   // len(seq) - 2 >= min_time_step (search.py:85-87)
          const int dr = nd + n_c;
          a.done_score[b * K + dr] = v;
          a.done_step[b * K + dr] = a.t;
          a.done_par[b * K + dr] = ii;
          ++n_c;
        }                                                      // an early <END> is dropped
      } else {
        s_ns[n_l] = v; s_nt[n_l] = tk; s_np[n_l] = ii;
        ++n_l;
      }
    }
    s_counts[0] = n_c;
    s_counts[1] = n_l;
  }
  __syncthreads();
  if (c < K) {
    const int n_l = s_counts[1];
    a.score[b * K + c] = s_ns[c];
    a.live[b * K + c] = c < n_l;
    tok_t[c] = s_nt[c];
    par_t[c] = s_np[c];
    a.parent_out[b * K + c] = b * K + s_np[c];
    a.last_tok[b * K + c] = s_nt[c];
  }
  if (c == 0) {
    a.n_done[b] = nd + s_counts[0];
    a.steps[b] = st + 1;
  }
}

int beam_update(const BeamUpd& a, cudaStream_t st) {
  if (a.B == 0) return GTOS_OK;
  GTOS_REQUIRE(a.K >= 1 && a.K <= 16, "beam_update: beam size %d not in [1, 16]", a.K);
  GTOS_REQUIRE(a.t >= 0 && a.t < a.Tmax, "beam_update: step %d outside [0, %d)", a.t, a.Tmax);
  int threads = ((a.K * a.K + 31) / 32) * 32;
  if (threads < 32) threads = 32;
  GTOS_KLAUNCH(beam_update_kernel, dim3((unsigned)a.B), dim3(threads), 0, st, a);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// f-4: global-norm clip + Adam with decoupled weight decay over FLAT buffers.
// Reference: torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0) (generator/train.py:152) followed by
// AdamWeightDecayOptimizer.step (generator/adam.py:28-87: no bias correction, update = m / (sqrt(v) + eps) + wd * p,
// p -= lr * update; two groups, wd on non-bias / non-LayerNorm parameters, train.py:123-132) -- 182 parameters x ~8 small
// launches there, two launches here.  Elements [0, n_decay) take the weight decay, [n_decay, n) do not.
// ---------------------------------------------------------------------------------------
__global__ void sumsq_partial_kernel(const float* __restrict__ g, long n, float* __restrict__ partials) {
  GTOS_PDL_PROLOGUE();
  __shared__ float sh[32];
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  const long n4 = n / 4;
  float s = 0.f;
  for (long i = tid; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long i = n4 * 4 + tid; i < n; i += stride) s += g[i] * g[i];
  s = blk_reduce(s, sh, false);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void sumsq_final_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  GTOS_PDL_PROLOGUE();
  __shared__ double shd[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partials[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += shd[i];
    out[0] = (float)t;
  }
}

static const int kSumsqBlocks = 148 * 4;

long grad_sumsq_workspace() { return kSumsqBlocks; }

int grad_sumsq(const float* g, long n, float* out, float* workspace, cudaStream_t st) {
  GTOS_REQUIRE(g && out && workspace, "grad_sumsq: null argument");
  GTOS_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "grad_sumsq: buffer must be 16-byte aligned");
  GTOS_KLAUNCH(sumsq_partial_kernel, dim3(kSumsqBlocks), dim3(256), 0, st, g, n, workspace);
  GTOS_LAUNCH_CHECK();
  GTOS_KLAUNCH(sumsq_final_kernel, dim3(1), dim3(256), 0, st, workspace, kSumsqBlocks, out);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float clip, float lr, float b1, float b2,
                                          float eps, float wd) {
  g *= clip;
  m = m * b1 + (1.f - b1) * g;
  v = v * b2 + (1.f - b2) * g * g;
  const float upd = m / (sqrtf(v) + eps) + wd * p;
  p -= lr * upd;
}

__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long n, long n_decay, const float* __restrict__ lr_ptr, float b1,
                                 float b2, float eps, float wd, const float* __restrict__ norm_sq, float max_norm) {
  GTOS_PDL_PROLOGUE();
  const float lr = lr_ptr[0];
  float clip = 1.f;
  if (norm_sq) {
    const float c = max_norm / (sqrtf(norm_sq[0]) + 1e-6f);      // clip_grad_norm_: coef clamped to 1
    clip = c < 1.f ? c : 1.f;
  }
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long)gridDim.x * blockDim.x;
  const long n4 = n / 4;
  for (long i = tid; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    const long e = i * 4;
    adam_elem(pp.x, gg.x, mm.x, vv.x, clip, lr, b1, b2, eps, e + 0 < n_decay ? wd : 0.f);
    adam_elem(pp.y, gg.y, mm.y, vv.y, clip, lr, b1, b2, eps, e + 1 < n_decay ? wd : 0.f);
    adam_elem(pp.z, gg.z, mm.z, vv.z, clip, lr, b1, b2, eps, e + 2 < n_decay ? wd : 0.f);
    adam_elem(pp.w, gg.w, mm.w, vv.w, clip, lr, b1, b2, eps, e + 3 < n_decay ? wd : 0.f);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long i = n4 * 4 + tid; i < n; i += stride) adam_elem(p[i], g[i], m[i], v[i], clip, lr, b1, b2, eps, i < n_decay ? wd : 0.f);
}

int adam_step(float* p, const float* g, float* m, float* v, long n, long n_decay, const float* lr_ptr, float b1, float b2,
              float eps, float wd, const float* norm_sq, float max_norm, cudaStream_t st) {
  if (n == 0) return GTOS_OK;
  GTOS_REQUIRE(p && g && m && v && lr_ptr, "adam_step: null argument");
  GTOS_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
  long blocks = (n / 4 + 255) / 256 + 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  GTOS_KLAUNCH(adam_step_kernel, dim3((unsigned)blocks), dim3(256), 0, st, p, g, m, v, n, n_decay, lr_ptr, b1, b2, eps, wd, norm_sq, max_norm);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// f-1: ancestry table of the token-side caches.  anc[l][h] = cache slot that holds position l of hypothesis h's prefix.
// After a beam step re-parents the live hypotheses (search.py:57-92), new_anc[l][h] = old_anc[l][parent[h]] for l < t and
// new_anc[t][h] = h (the row this step appends).  Tmax x Hyp int32 per step instead of re-gathering every cached state.
// ---------------------------------------------------------------------------------------
__global__ void beam_ancestry_kernel(const int* __restrict__ old_anc, int* __restrict__ new_anc, long ld,
                                     const int* __restrict__ parent, int t, int Hyp) {
  GTOS_PDL_PROLOGUE();
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)(t + 1) * Hyp;
  if (tid >= total) return;
  const int l = (int)(tid / Hyp), h = (int)(tid % Hyp);
  new_anc[(long)l * ld + h] = (l == t) ? h : old_anc[(long)l * ld + (parent ? parent[h] : h)];
}

int beam_ancestry(const int* old_anc, int* new_anc, long ld, const int* parent, int t, int Hyp, cudaStream_t st) {
  if (Hyp == 0) return GTOS_OK;
  GTOS_REQUIRE(new_anc && (t == 0 || old_anc) && old_anc != new_anc, "beam_ancestry: need distinct old / new tables");
  const long total = (long)(t + 1) * Hyp;
  GTOS_KLAUNCH(beam_ancestry_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, old_anc, new_anc, ld, parent, t, Hyp);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int beam_update_c(int B, int K, int t, int Tmin, int Tmax, int end_id, int unk_id, const float* top_val, const int* top_idx,
                  float* score, unsigned char* live, int* n_done, int* steps, int* tok, int* par, float* done_score,
                  int* done_step, int* done_par, int* parent_out, long long* last_tok, cudaStream_t st) {
  BeamUpd a;
  a.B = B; a.K = K; a.t = t; a.Tmin = Tmin; a.Tmax = Tmax; a.end_id = end_id; a.unk_id = unk_id;
  a.top_val = top_val; a.top_idx = top_idx; a.score = score; a.live = live; a.n_done = n_done; a.steps = steps;
  a.tok = tok; a.par = par; a.done_score = done_score; a.done_step = done_step; a.done_par = done_par;
  a.parent_out = parent_out; a.last_tok = last_tok;
  GTOS_REQUIRE(top_val && top_idx && score && live && n_done && steps && tok && par && done_score && done_step && done_par &&
               parent_out && last_tok, "beam_update: null argument");
  return beam_update(a, st);
}

}  // namespace gtos
