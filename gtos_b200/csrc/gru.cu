// gtos_b200 -- RelationEncoder helpers: embedding gather / scatter-add and the GRU gate math
// (forward and backward through time).  The matrix products of the GRU run on the tcgen05 GEMM
// (gemm.cu); these kernels are the per-timestep elementwise epilogues with packed-sequence
// masking.  Reference: generator/encoder.py:90-119 (nn.GRU, gate order r,z,n).
#include "elementwise.cuh"

namespace gtos {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void gru_gate_fwd_kernel(const float* __restrict__ gi, long ldgi, const float* __restrict__ gh, long ldgh,
                                    const float* __restrict__ h_prev, const long long* __restrict__ lengths, int t,
                                    float* __restrict__ h_new, __nv_bfloat16* __restrict__ h_new_bf16,
                                    float* __restrict__ out_t, long ldout, __nv_bfloat16* __restrict__ out_t_bf16,
                                    long ldoutb, float* __restrict__ gates, long R, int Hh) {
  const long total = R * Hh;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long r = idx / Hh;
    const int c = (int)(idx % Hh);
    const bool live = lengths[r] > t;
    const float hp = h_prev[idx];
    float hn = hp, o = 0.f, gr = 0.f, gz = 0.f, gn = 0.f;
    if (live) {
      const float* a = gi + r * ldgi;
      const float* bq = gh + r * ldgh;
      gr = sigmoidf_(a[c] + bq[c]);
      gz = sigmoidf_(a[Hh + c] + bq[Hh + c]);
      gn = tanhf(a[2 * Hh + c] + gr * bq[2 * Hh + c]);
      hn = (1.f - gz) * gn + gz * hp;
      o = hn;
    }
    h_new[idx] = hn;
    if (h_new_bf16) h_new_bf16[idx] = __float2bfloat16(hn);
    if (out_t) out_t[r * ldout + c] = o;
    if (out_t_bf16) out_t_bf16[r * ldoutb + c] = __float2bfloat16(o);
    if (gates) {
      gates[r * 3 * Hh + c] = gr;
      gates[r * 3 * Hh + Hh + c] = gz;
      gates[r * 3 * Hh + 2 * Hh + c] = gn;
    }
  }
}

int gru_gate_fwd(const float* gi, long ldgi, const float* gh, long ldgh, const float* h_prev, const long long* lengths,
                 int t, float* h_new, void* h_new_bf16, float* out_t, long ldout, void* out_t_bf16, long ldoutb,
                 float* gates, long R, int Hh, cudaStream_t st) {
  if (R == 0) return GTOS_OK;
  long blocks = (R * Hh + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gru_gate_fwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(gi, ldgi, gh, ldgh, h_prev, lengths, t, h_new,
                                                        reinterpret_cast<__nv_bfloat16*>(h_new_bf16), out_t, ldout,
                                                        reinterpret_cast<__nv_bfloat16*>(out_t_bf16), ldoutb, gates, R,
                                                        Hh);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// dh_tot = dh + dout_t ;  n,z,r chain rule ; dh_prev = dh_tot * z (the W_hh^T dgh term is added by a
// following GEMM with accumulate) ; inactive rows pass dh through untouched and emit zero gate grads.
__global__ void gru_gate_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ dout_t, long lddout,
                                    const float* __restrict__ gates, const float* __restrict__ gh, long ldgh,
                                    const float* __restrict__ h_prev, const long long* __restrict__ lengths, int t,
                                    float* __restrict__ dh_prev, __nv_bfloat16* __restrict__ dgi, long lddgi,
                                    __nv_bfloat16* __restrict__ dgh, long lddgh, long R, int Hh) {
  const long total = R * Hh;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long r = idx / Hh;
    const int c = (int)(idx % Hh);
    const bool live = lengths[r] > t;
    float d = dh ? dh[idx] : 0.f;
    float dar = 0.f, daz = 0.f, dan = 0.f, dhn = 0.f, dprev = d;
    if (live) {
      if (dout_t) d += dout_t[r * lddout + c];
      const float gr = gates[r * 3 * Hh + c], gz = gates[r * 3 * Hh + Hh + c], gn = gates[r * 3 * Hh + 2 * Hh + c];
      const float hnn = gh[r * ldgh + 2 * Hh + c];  // W_hn h + b_hn
      const float hp = h_prev[idx];
      const float dn = d * (1.f - gz);
      const float dz = d * (hp - gn);
      dan = dn * (1.f - gn * gn);
      daz = dz * gz * (1.f - gz);
      dar = dan * hnn * gr * (1.f - gr);
      dhn = dan * gr;
      dprev = d * gz;
    }
    dh_prev[idx] = dprev;
    dgi[r * lddgi + c] = __float2bfloat16(dar);
    dgi[r * lddgi + Hh + c] = __float2bfloat16(daz);
    dgi[r * lddgi + 2 * Hh + c] = __float2bfloat16(dan);
    dgh[r * lddgh + c] = __float2bfloat16(dar);
    dgh[r * lddgh + Hh + c] = __float2bfloat16(daz);
    dgh[r * lddgh + 2 * Hh + c] = __float2bfloat16(dhn);
  }
}

int gru_gate_bwd(const float* dh, const float* dout_t, long lddout, const float* gates, const float* gh, long ldgh,
                 const float* h_prev, const long long* lengths, int t, float* dh_prev, void* dgi_bf16, long lddgi,
                 void* dgh_bf16, long lddgh, long R, int Hh, cudaStream_t st) {
  if (R == 0) return GTOS_OK;
  long blocks = (R * Hh + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gru_gate_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(dh, dout_t, lddout, gates, gh, ldgh, h_prev, lengths, t, dh_prev,
                                                        reinterpret_cast<__nv_bfloat16*>(dgi_bf16), lddgi,
                                                        reinterpret_cast<__nv_bfloat16*>(dgh_bf16), lddgh, R, Hh);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// out[n, :] = dropout(table[idx[n], :]) ; bf16 copy is zero-padded to ldb columns
__global__ void embed_gather_kernel(const float* __restrict__ table, const long long* __restrict__ idx, long n, int dim,
                                    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, long ldb,
                                    float p, const unsigned long long* seed_ptr, unsigned long long seed_off) {
  const unsigned long long seed = p > 0.f ? seed_ptr[0] + seed_off : 0ull;
  const float ks = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const long total = n * ldb;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / ldb;
    const int c = (int)(i % ldb);
    float v = 0.f;
    if (c < dim) {
      v = table[idx[r] * dim + c];
      if (p > 0.f) v = (rng_uniform(seed, (unsigned long long)(r * dim + c)) >= p) ? v * ks : 0.f;
      if (out_f32) out_f32[r * dim + c] = v;
    }
    if (out_bf16) out_bf16[i] = __float2bfloat16(v);
  }
}

int embed_gather(const float* table, const long long* idx, long n, int dim, float* out_f32, void* out_bf16, long ldb,
                 float p_drop, const void* seed_ptr, unsigned long long seed_off, cudaStream_t st) {
  if (n == 0) return GTOS_OK;
  GTOS_REQUIRE(ldb >= dim, "embed_gather: ldb < dim");
  long blocks = (n * ldb + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  embed_gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(table, idx, n, dim, out_f32,
                                                        reinterpret_cast<__nv_bfloat16*>(out_bf16), ldb, p_drop,
                                                        reinterpret_cast<const unsigned long long*>(seed_ptr), seed_off);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

__global__ void embed_scatter_kernel(const float* __restrict__ dx, long lddx, const long long* __restrict__ idx, long n,
                                     int dim, float* __restrict__ dtable, float p, const unsigned long long* seed_ptr,
                                     unsigned long long seed_off) {
  const unsigned long long seed = p > 0.f ? seed_ptr[0] + seed_off : 0ull;
  const float ks = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const long total = n * dim;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / dim;
    const int c = (int)(i % dim);
    float v = dx[r * lddx + c];
    if (p > 0.f) v = (rng_uniform(seed, (unsigned long long)(r * dim + c)) >= p) ? v * ks : 0.f;
    if (v != 0.f) atomicAdd(&dtable[idx[r] * dim + c], v);
  }
}

int embed_scatter_add(const float* dx, const long long* idx, long n, int dim, float* dtable, float p_drop,
                      const void* seed_ptr, unsigned long long seed_off, cudaStream_t st) {
  if (n == 0) return GTOS_OK;
  long blocks = (n * dim + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  embed_scatter_kernel<<<(unsigned)blocks, 256, 0, st>>>(dx, dim, idx, n, dim, dtable, p_drop,
                                                         reinterpret_cast<const unsigned long long*>(seed_ptr), seed_off);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

}  // namespace gtos
