// gtos_b200 -- kernels of the widened rows of SURVEY.md §8(f): incremental beam decode (f-1), the work=True
// log-probability table (f-2) and the fused Adam / global-norm clip over flat buffers (f-4).
// All of them are HBM/L2-bound streaming kernels: no tensor-core shapes here (T_q = 1, or one pass over a vector).
#include <math.h>

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
  int Hyp, L, H, hd;
  const float* q; long ldq;
  const __nv_bfloat16* kv; long ld_kv; int v_off; long row_stride;
  const int* slot; long slot_ld;
  const unsigned char* key_pad; long pad_ld;
  float scale;
  float* out; long ldo;
  __nv_bfloat16* out_bf16; long ldob;
  float* probs;
};

__global__ void attn_decode_kernel(const DecodeAttn a) {
  extern __shared__ float sm[];
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long item = (long)blockIdx.x * wpb + w;
  if (item >= (long)a.Hyp * a.H) return;                       // warps are independent: only __syncwarp below
  const int hyp = (int)(item / a.H), head = (int)(item % a.H);
  const int hd = a.hd, L = a.L;
  float* qs = sm + (size_t)w * (hd + 2 * L);
  float* sc = qs + hd;
  int* rows = reinterpret_cast<int*>(sc + L);
  const float* q = a.q + (long)hyp * a.ldq + head * hd;
  for (int d = lane; d < hd; d += 32) qs[d] = q[d] * a.scale;
  __syncwarp();
  const __nv_bfloat16* kbase = a.kv + head * hd;
  float mx = -INFINITY;
  for (int l = lane; l < L; l += 32) {
    const int s = a.slot ? a.slot[(long)l * a.slot_ld + hyp] : hyp;
    const long row = (long)l * a.row_stride + s;
    const bool masked = a.key_pad && a.key_pad[(long)l * a.pad_ld + s];
    const uint4* kr = reinterpret_cast<const uint4*>(kbase + row * a.ld_kv);
    float dot = 0.f;
    for (int c = 0; c < hd / 8; ++c) {
      const uint4 u = kr[c];
      const float* qq = qs + c * 8;
      const __nv_bfloat162 p0 = *reinterpret_cast<const __nv_bfloat162*>(&u.x), p1 = *reinterpret_cast<const __nv_bfloat162*>(&u.y),
                           p2 = *reinterpret_cast<const __nv_bfloat162*>(&u.z), p3 = *reinterpret_cast<const __nv_bfloat162*>(&u.w);
      dot += qq[0] * __low2float(p0) + qq[1] * __high2float(p0) + qq[2] * __low2float(p1) + qq[3] * __high2float(p1) +
             qq[4] * __low2float(p2) + qq[5] * __high2float(p2) + qq[6] * __low2float(p3) + qq[7] * __high2float(p3);
    }
    dot = masked ? -INFINITY : dot;
    sc[l] = dot;
    rows[l] = (int)row;
    mx = fmaxf(mx, dot);
  }
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
  const __nv_bfloat16* vbase = a.kv + a.v_off + head * hd;
  for (int dp = lane; dp < hd / 2; dp += 32) {
    float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;
    int l = 0;
    for (; l + 1 < L; l += 2) {
      const __nv_bfloat162 v0 = *reinterpret_cast<const __nv_bfloat162*>(vbase + (long)rows[l] * a.ld_kv + 2 * dp);
      const __nv_bfloat162 v1 = *reinterpret_cast<const __nv_bfloat162*>(vbase + (long)rows[l + 1] * a.ld_kv + 2 * dp);
      const float p0 = sc[l], p1 = sc[l + 1];
      ax += p0 * __low2float(v0); ay += p0 * __high2float(v0);
      bx += p1 * __low2float(v1); by += p1 * __high2float(v1);
    }
    if (l < L) {
      const __nv_bfloat162 v0 = *reinterpret_cast<const __nv_bfloat162*>(vbase + (long)rows[l] * a.ld_kv + 2 * dp);
      ax += sc[l] * __low2float(v0); ay += sc[l] * __high2float(v0);
    }
    ax += bx; ay += by;
    if (a.out) *reinterpret_cast<float2*>(a.out + (long)hyp * a.ldo + head * hd + 2 * dp) = make_float2(ax, ay);
    if (a.out_bf16)
      *reinterpret_cast<uint32_t*>(a.out_bf16 + (long)hyp * a.ldob + head * hd + 2 * dp) = pack_bf16x2(ax, ay);
  }
}

int attn_decode(int Hyp, int L, int H, int hd, const float* q, long ldq, const void* kv, long ld_kv, int v_off,
                long row_stride, const int* slot, long slot_ld, const unsigned char* key_pad, long pad_ld, float scale,
                float* out, long ldo, void* out_bf16, long ldob, float* probs, cudaStream_t st) {
  if (Hyp == 0 || H == 0) return GTOS_OK;
  GTOS_REQUIRE(L > 0 && hd > 0 && hd % 8 == 0, "attn_decode: need L > 0 and head_dim %% 8 == 0 (L=%d, hd=%d)", L, hd);
  GTOS_REQUIRE(ld_kv % 8 == 0 && v_off % 8 == 0, "attn_decode: cache row stride / value offset must be multiples of 8");
  GTOS_REQUIRE(q && kv && (out || out_bf16), "attn_decode: null argument");
  GTOS_REQUIRE(ldo % 2 == 0 && ldob % 2 == 0, "attn_decode: output strides must be even");
  const size_t per_warp = (size_t)(hd + 2 * L) * sizeof(float);
  int wpb = 4;
  while (wpb > 1 && per_warp * wpb > 200 * 1024) wpb >>= 1;
  GTOS_REQUIRE(per_warp * wpb <= 200 * 1024, "attn_decode: L=%d keys do not fit in shared memory", L);
  const size_t smem = per_warp * wpb;
  if (smem > 48 * 1024)
    GTOS_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  DecodeAttn a;
  a.Hyp = Hyp; a.L = L; a.H = H; a.hd = hd; a.q = q; a.ldq = ldq;
  a.kv = reinterpret_cast<const __nv_bfloat16*>(kv); a.ld_kv = ld_kv; a.v_off = v_off; a.row_stride = row_stride;
  a.slot = slot; a.slot_ld = slot_ld; a.key_pad = key_pad; a.pad_ld = pad_ld; a.scale = scale;
  a.out = out; a.ldo = ldo; a.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); a.ldob = ldob; a.probs = probs;
  const long items = (long)Hyp * H;
  attn_decode_kernel<<<(unsigned)((items + wpb - 1) / wpb), wpb * 32, smem, st>>>(a);
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
  token_logprob_kernel<<<(unsigned)rows, 256, 0, st>>>(logits, ldl, V, gate_logits, align, S, copy_seq, Bsrc, src_index, B,
                                                       table, ldt, W);
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
  sumsq_partial_kernel<<<kSumsqBlocks, 256, 0, st>>>(g, n, workspace);
  GTOS_LAUNCH_CHECK();
  sumsq_final_kernel<<<1, 256, 0, st>>>(workspace, kSumsqBlocks, out);
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
  adam_step_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, n_decay, lr_ptr, b1, b2, eps, wd, norm_sq, max_norm);
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
  beam_ancestry_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(old_anc, new_anc, ld, parent, t, Hyp);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

}  // namespace gtos
