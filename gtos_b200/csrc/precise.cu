// gtos_b200 -- fp32 mode ("1e-3 against the fp32 reference", BASELINE north star): the pieces around the GEMMs.
//
// The reference computes in fp32 end to end (generator/graph_transformer.py:122-133, transformer.py:111-162,
// encoder.py:90-119).  The default path of this library feeds the tensor cores bf16 operands (1e-2 tolerance).  In fp32
// mode every matrix product still runs on tcgen05, but on SPLIT operands: x = x_hi + x_lo with x_hi = bf16(x),
// x_lo = bf16(x - x_hi) (16 significand bits), and
//     A B^T  ~=  A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T            (dropped term A_lo B_lo^T: 2^-18 relative)
// is ONE call of the unchanged GEMM kernels over a K-tripled operand pair:
//     [A_hi | A_lo | A_hi] (role 0)   x   [B_hi | B_hi | B_lo] (role 1),   fp32 accumulation in TMEM.
// split3_kernel below writes those tripled operands.  Because the three thirds are contiguous, the same buffer viewed as
// [3 rows, kp] (row 3r+s = third s of row r) is the row-stacked operand of the weight-gradient GEMM (gtos_gemm_nn sums over
// rows), so one staging pass per tensor serves the forward GEMM and both backward GEMMs.
//
// Everything that is not a GEMM keeps fp32 values between kernels in this mode: the per-pair relation scores and their
// gradient (rel_score_f32 / rel_grad_f32 / rel_dqk_f32: the relation projection is a K-tripled GEMM into an fp32 [P, 2D]
// tensor, these kernels do the per-pair q/k adds and per-head dots that the bf16 path fuses into its tcgen05 epilogue),
// the GRU gate math with fp32 saved gates (gru_gate_fwd_f32 / gru_gate_bwd_f32), the FFN's ReLU/dropout backward on the
// fp32 activation.  The attention core runs its three-pass variant (attention.cu, AttnArgs::precise).
#include "elementwise.cuh"

namespace gtos {

// ---------------------------------------------------------------------------------------
// split3: src fp32 (element (r, c) at src[r * ld_r + c * ld_c], so a transposed view costs nothing) ->
// dst bf16 [rows, 3 * kp] (row stride ldd): thirds (hi, lo, hi) for role 0, (hi, hi, lo) for role 1; columns
// cols..kp-1 of every third are zero (kp % 8 == 0: TMA rows are 16-byte multiples).
// ---------------------------------------------------------------------------------------
__global__ void split3_kernel(const float* __restrict__ src, long ld_r, long ld_c, long rows, int cols,
                              __nv_bfloat16* __restrict__ dst, long ldd, int kp, int role) {
  GTOS_PDL_PROLOGUE();
  const long total = rows * kp;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / kp;
    const int c = (int)(i - r * kp);
    const float v = c < cols ? src[r * ld_r + c * ld_c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
    __nv_bfloat16* d = dst + r * ldd + c;
    d[0] = hi;
    d[kp] = role == 0 ? lo : hi;
    d[2 * kp] = role == 0 ? hi : lo;
  }
}

// contiguous rows (ld_c == 1), 4 columns per thread: 16-byte loads, 8-byte stores
__global__ void split3_vec_kernel(const float* __restrict__ src, long ld_r, long rows, int cols,
                                  __nv_bfloat16* __restrict__ dst, long ldd, int kp, int role) {
  GTOS_PDL_PROLOGUE();
  const int q = kp >> 2;
  const long total = rows * q;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / q;
    const int c = (int)(i - r * q) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c + 3 < cols) {
      v = *reinterpret_cast<const float4*>(src + r * ld_r + c);
    } else {
      const float* p = src + r * ld_r + c;
      if (c < cols) v.x = p[0];
      if (c + 1 < cols) v.y = p[1];
      if (c + 2 < cols) v.z = p[2];
    }
    const uint2 hi = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hi);
    const float2 h01 = __bfloat1622float2(h2[0]), h23 = __bfloat1622float2(h2[1]);
    const uint2 lo = make_uint2(pack_bf16x2(v.x - h01.x, v.y - h01.y), pack_bf16x2(v.z - h23.x, v.w - h23.y));
    __nv_bfloat16* d = dst + r * ldd + c;
    *reinterpret_cast<uint2*>(d) = hi;
    *reinterpret_cast<uint2*>(d + kp) = role == 0 ? lo : hi;
    *reinterpret_cast<uint2*>(d + 2 * kp) = role == 0 ? hi : lo;
  }
}

int split3(const float* src, long ld_r, long ld_c, long rows, int cols, void* dst, long ldd, int kp, int role,
           cudaStream_t st) {
  if (rows == 0) return GTOS_OK;
  GTOS_REQUIRE(kp % 8 == 0 && kp >= cols && ldd >= 3L * kp && ldd % 8 == 0 && (role == 0 || role == 1),
               "split3: kp must be a multiple of 8 >= cols, ldd >= 3 kp (cols=%d kp=%d ldd=%ld)", cols, kp, ldd);
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
  const bool vec = ld_c == 1 && (ld_r & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
  long work = vec ? rows * (kp / 4) : rows * (long)kp;
  long blocks = (work + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (vec)
    GTOS_KLAUNCH(split3_vec_kernel, dim3((unsigned)blocks), dim3(256), 0, st, src, ld_r, rows, cols, d, ldd, kp, role);
  else
    GTOS_KLAUNCH(split3_kernel, dim3((unsigned)blocks), dim3(256), 0, st, src, ld_r, ld_c, rows, cols, d, ldd, kp, role);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// relation scores in fp32 (graph_transformer.py:122-133).  PR fp32 [P, 2D] = relation_in_proj(relation) in the reference's
// column order [ra (D) | rb (D)], row p = (j * N + i) * B + b  (relation[j][i][b]: query i, key j - the `.transpose(0, 1)`
// of :124-125).  qkv rows (n * B + b): q at column 0, k at column D (row stride ldqk).
//   scores[b,h,j,i] = hd^-1/2 * sum_{d in head h} (q[i,b,d] + ra[p,d]) * (k[j,b,d] + rb[p,d])
// One warp per pair; a head's hd features are spread over the lanes, one shuffle reduction per head.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rel_score_f32_kernel(const float* __restrict__ PR, long ldpr,
                                                            const float* __restrict__ q, const float* __restrict__ k,
                                                            long ldqk, float* __restrict__ scores, int N, int B, int D,
                                                            int H, float scale) {
  GTOS_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const long P = (long)N * N * B;
  const int hd = D / H;
  for (long p = warp; p < P; p += nwarps) {
    const int b = (int)(p % B);
    const long ji = p / B;
    const int i = (int)(ji % N), j = (int)(ji / N);
    const float* ra = PR + p * ldpr;
    const float* rb = ra + D;
    const float* qi = q + ((long)i * B + b) * ldqk;
    const float* kj = k + ((long)j * B + b) * ldqk;
    for (int h = 0; h < H; ++h) {
      float acc = 0.f;
      for (int d = h * hd + lane; d < (h + 1) * hd; d += 32) acc = fmaf(qi[d] + ra[d], kj[d] + rb[d], acc);
      acc = warp_sum(acc);
      if (lane == 0) scores[(((long)b * H + h) * N + j) * N + i] = acc * scale;
    }
  }
}

int rel_score_f32(const float* PR, long ldpr, const float* q, const float* k, long ldqk, float* scores, int N, int B,
                  int D, int H, cudaStream_t st) {
  if (N == 0 || B == 0) return GTOS_OK;
  GTOS_REQUIRE(H > 0 && D % H == 0 && ldpr >= 2L * D, "rel_score_f32: bad shape (D=%d H=%d ldpr=%ld)", D, H, ldpr);
  const long P = (long)N * N * B;
  long blocks = (P + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  GTOS_KLAUNCH(rel_score_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, st, PR, ldpr, q, k, ldqk, scores, N, B, D, H,
               1.0f / sqrtf((float)(D / H)));
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// gradient of the scores w.r.t. the per-pair projections, fp32:
//   G[p, d]     = d ra = hd^-1/2 * dscores[b,h(d),j,i] * (k[j,b,d] + rb[p,d])      (also the pair's share of dq[i,b,d])
//   G[p, D + d] = d rb = hd^-1/2 * dscores[b,h(d),j,i] * (q[i,b,d] + ra[p,d])      (also its share of dk[j,b,d])
__global__ void __launch_bounds__(256) rel_grad_f32_kernel(const float* __restrict__ PR, long ldpr,
                                                           const float* __restrict__ q, const float* __restrict__ k,
                                                           long ldqk, const float* __restrict__ dscores,
                                                           float* __restrict__ G, long ldg, int N, int B, int D, int H,
                                                           float scale) {
  GTOS_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const long P = (long)N * N * B;
  const int hd = D / H;
  for (long p = warp; p < P; p += nwarps) {
    const int b = (int)(p % B);
    const long ji = p / B;
    const int i = (int)(ji % N), j = (int)(ji / N);
    const float* ra = PR + p * ldpr;
    const float* rb = ra + D;
    const float* qi = q + ((long)i * B + b) * ldqk;
    const float* kj = k + ((long)j * B + b) * ldqk;
    float* g = G + p * ldg;
    for (int d = lane; d < D; d += 32) {
      const int h = d / hd;
      const float ds = dscores[(((long)b * H + h) * N + j) * N + i] * scale;
      g[d] = ds * (kj[d] + rb[d]);
      g[D + d] = ds * (qi[d] + ra[d]);
    }
  }
}

int rel_grad_f32(const float* PR, long ldpr, const float* q, const float* k, long ldqk, const float* dscores, float* G,
                 long ldg, int N, int B, int D, int H, cudaStream_t st) {
  if (N == 0 || B == 0) return GTOS_OK;
  GTOS_REQUIRE(H > 0 && D % H == 0 && ldpr >= 2L * D && ldg >= 2L * D, "rel_grad_f32: bad shape");
  const long P = (long)N * N * B;
  long blocks = (P + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  GTOS_KLAUNCH(rel_grad_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, st, PR, ldpr, q, k, ldqk, dscores, G, ldg, N, B,
               D, H, 1.0f / sqrtf((float)(D / H)));
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// dq[i,b,:] = sum_j G[(j,i,b), 0:D]   (blockIdx.y == 0)      dk[j,b,:] = sum_i G[(j,i,b), D:2D]   (blockIdx.y == 1)
// fixed summation order (no atomics): one block per (node, graph), one thread per feature
__global__ void __launch_bounds__(256) rel_dqk_f32_kernel(const float* __restrict__ G, long ldg, float* __restrict__ dq,
                                                          float* __restrict__ dk, long ld, int N, int B, int D) {
  GTOS_PDL_PROLOGUE();
  const int n = blockIdx.x / B, b = blockIdx.x % B;
  const bool key_side = blockIdx.y == 1;
  float* out = (key_side ? dk : dq) + ((long)n * B + b) * ld;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
#pragma unroll 8
    for (int m = 0; m < N; ++m) {
      const long p = key_side ? ((long)n * N + m) * B + b : ((long)m * N + n) * B + b;
      acc += G[p * ldg + (key_side ? D : 0) + d];
    }
    out[d] = acc;
  }
}

int rel_dqk_f32(const float* G, long ldg, float* dq, float* dk, long ld, int N, int B, int D, cudaStream_t st) {
  if (N == 0 || B == 0) return GTOS_OK;
  GTOS_REQUIRE(ldg >= 2L * D && ld >= D, "rel_dqk_f32: bad strides");
  GTOS_KLAUNCH(rel_dqk_f32_kernel, dim3((unsigned)(N * B), 2), dim3(256), 0, st, G, ldg, dq, dk, ld, N, B, D);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// FFN backward through dropout(relu(.)) on the fp32 activation the forward kept (post-dropout: zero where either killed it)
__global__ void relu_drop_bwd_f32_kernel(const float* __restrict__ dh_in, const float* __restrict__ act,
                                         float* __restrict__ dh_out, long n, float p) {
  GTOS_PDL_PROLOGUE();
  const float ks = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dh_out[i] = act[i] > 0.f ? dh_in[i] * ks : 0.f;
}

int relu_drop_bwd_f32(const float* dh_in, const float* act, float* dh_out, long n, float p, cudaStream_t st) {
  if (n == 0) return GTOS_OK;
  long blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  GTOS_KLAUNCH(relu_drop_bwd_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, st, dh_in, act, dh_out, n, p);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// GRU cell in fp32 (nn.GRU, gate order r, z, n; encoder.py:105-106), packed-sequence masking by `lengths`:
//   gi = x_t W_ih^T + b_ih  [R, 3H]   gh = h W_hh^T + b_hh  [R, 3H]    (K-tripled GEMMs)
//   r = sigmoid(gi_r + gh_r)   z = sigmoid(gi_z + gh_z)   n = tanh(gi_n + r * gh_n)   h' = (1 - z) n + z h
// rows with lengths[row] <= t keep h and emit a zero layer output.  gates fp32 [R, 4H] = [r | z | n | gh_n] for backward.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void gru_gate_fwd_f32_kernel(const float* __restrict__ gi, long ldgi, const float* __restrict__ gh, long ldgh,
                                        const float* __restrict__ h_prev, const long long* __restrict__ lengths, int t,
                                        float* __restrict__ h_new, float* __restrict__ out_t, long ldout,
                                        float* __restrict__ gates, long R, int Hh) {
  GTOS_PDL_PROLOGUE();
  const long total = R * Hh;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long r = e / Hh;
    const int c = (int)(e - r * Hh);
    const float hp = h_prev[e];
    const bool live = lengths[r] > t;
    float hn = hp, o = 0.f, gr = 0.f, gz = 0.f, gn = 0.f, ghn = 0.f;
    if (live) {
      const float* a = gi + r * ldgi;
      const float* bb = gh + r * ldgh;
      gr = sigmoid_exact(a[c] + bb[c]);
      gz = sigmoid_exact(a[Hh + c] + bb[Hh + c]);
      ghn = bb[2 * Hh + c];
      gn = tanhf(a[2 * Hh + c] + gr * ghn);
      hn = (1.f - gz) * gn + gz * hp;
      o = hn;
    }
    h_new[e] = hn;
    if (out_t) out_t[r * ldout + c] = o;
    if (gates) {
      float* g = gates + r * 4L * Hh;
      g[c] = gr; g[Hh + c] = gz; g[2 * Hh + c] = gn; g[3 * Hh + c] = ghn;
    }
  }
}

int gru_gate_fwd_f32(const float* gi, long ldgi, const float* gh, long ldgh, const float* h_prev, const long long* lengths,
                     int t, float* h_new, float* out_t, long ldout, float* gates, long R, int Hh, cudaStream_t st) {
  if (R == 0) return GTOS_OK;
  long blocks = (R * Hh + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  GTOS_KLAUNCH(gru_gate_fwd_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, st, gi, ldgi, gh, ldgh, h_prev, lengths, t,
               h_new, out_t, ldout, gates, R, Hh);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// backward of one step: dh_tot = dh + dout_t; dgi = [d a_r | d a_z | d a_n], dgh = [d a_r | d a_z | d a_n * r] (fp32 [R, 3H]);
// dh_part = dh_tot * z (the W_hh^T dgh term is added by the following GEMM); finished rows pass dh through and emit zeros.
__global__ void gru_gate_bwd_f32_kernel(const float* __restrict__ dh, const float* __restrict__ dout_t, long lddout,
                                        const float* __restrict__ gates, const float* __restrict__ h_prev,
                                        const long long* __restrict__ lengths, int t, float* __restrict__ dh_part,
                                        float* __restrict__ dgi, long lddgi, float* __restrict__ dgh, long lddgh, long R,
                                        int Hh) {
  GTOS_PDL_PROLOGUE();
  const long total = R * Hh;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long r = e / Hh;
    const int c = (int)(e - r * Hh);
    float d = dh ? dh[e] : 0.f;
    float dar = 0.f, daz = 0.f, dan = 0.f, dhn = 0.f, dprev = d;
    if (lengths[r] > t) {
      if (dout_t) d += dout_t[r * lddout + c];
      const float* g = gates + r * 4L * Hh;
      const float gr = g[c], gz = g[Hh + c], gn = g[2 * Hh + c], ghn = g[3 * Hh + c];
      const float dn = d * (1.f - gz);
      const float dz = d * (h_prev[e] - gn);
      dan = dn * (1.f - gn * gn);
      daz = dz * gz * (1.f - gz);
      dar = dan * ghn * gr * (1.f - gr);
      dhn = dan * gr;
      dprev = d * gz;
    }
    dh_part[e] = dprev;
    float* a = dgi + r * lddgi;
    float* bb = dgh + r * lddgh;
    a[c] = dar; a[Hh + c] = daz; a[2 * Hh + c] = dan;
    bb[c] = dar; bb[Hh + c] = daz; bb[2 * Hh + c] = dhn;
  }
}

int gru_gate_bwd_f32(const float* dh, const float* dout_t, long lddout, const float* gates, const float* h_prev,
                     const long long* lengths, int t, float* dh_part, float* dgi, long lddgi, float* dgh, long lddgh,
                     long R, int Hh, cudaStream_t st) {
  if (R == 0) return GTOS_OK;
  long blocks = (R * Hh + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  GTOS_KLAUNCH(gru_gate_bwd_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, st, dh, dout_t, lddout, gates, h_prev,
               lengths, t, dh_part, dgi, lddgi, dgh, lddgh, R, Hh);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

}  // namespace gtos
