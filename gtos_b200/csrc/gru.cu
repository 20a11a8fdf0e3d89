// gtos_b200 -- RelationEncoder helpers: embedding gather / scatter-add and the GRU gate math
// (forward and backward through time).  The matrix products of the GRU run on the tcgen05 GEMM
// (gemm.cu); these kernels are the per-timestep elementwise epilogues with packed-sequence
// masking.  Reference: generator/encoder.py:90-119 (nn.GRU, gate order r,z,n).
#include "elementwise.cuh"

namespace gtos {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// Gate-interleaved GRU weights for the fused step kernel (gemm.cu, MODE_GRU).
// Wcat[4H, Kx + H8] rows, per block of UB hidden units: [r | z | n_input | n_hidden]; columns [x part (Kx) | h part].
__global__ void gru_weight_prep_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                       const float* __restrict__ b_ih, const float* __restrict__ b_hh, int Kin, int H,
                                       int Kx, int UB, __nv_bfloat16* __restrict__ Wcat, long ldw,
                                       float* __restrict__ bcat) {
  const int pr = blockIdx.x;                 // permuted row in [0, 4H)
  const int u = pr / (4 * UB), g = (pr / UB) & 3, cc = pr % UB;
  const int c = u * UB + cc;                 // hidden unit
  const int gate = g < 2 ? g : 2;            // r, z, n
  const float* wi = w_ih + (long)(gate * H + c) * Kin;
  const float* wh = w_hh + (long)(gate * H + c) * H;
  for (int k = threadIdx.x; k < ldw; k += blockDim.x) {
    float v = 0.f;
    if (k < Kx) {
      if (k < Kin && g != 3) v = wi[k];
    } else if (k - Kx < H && g != 2) {
      v = wh[k - Kx];
    }
    Wcat[(long)pr * ldw + k] = __float2bfloat16(v);
  }
  if (threadIdx.x == 0) {
    float b = 0.f;
    if (g != 3) b += b_ih[gate * H + c];
    if (g != 2) b += b_hh[gate * H + c];
    bcat[pr] = b;
  }
}

int gru_weight_prep(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int Kin, int H, int Kx,
                    void* Wcat, long ldw, float* bcat, cudaStream_t st) {
  GTOS_REQUIRE(H % 16 == 0 && Kx % 64 == 0 && Kx >= Kin && ldw >= Kx + H, "gru_weight_prep: bad shape");
  const int UB = (H % 64 == 0) ? 64 : 16;
  gru_weight_prep_kernel<<<4 * H, 128, 0, st>>>(w_ih, w_hh, b_ih, b_hh, Kin, H, Kx, UB,
                                                reinterpret_cast<__nv_bfloat16*>(Wcat), ldw, bcat);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// dh_tot = dh + dout_t ;  n,z,r chain rule ; dh_prev = dh_tot * z (the W_hh^T dgh term is added by a
// following GEMM with accumulate) ; finished rows pass dh through untouched and emit zero gate grads.
// gates: bf16 [R, 4H] as saved by the fused step kernel (blocks of UB units: [r | z | n | W_hn h + b_hn]).
__global__ void gru_gate_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ dout_t, long lddout,
                                    const __nv_bfloat16* __restrict__ gates, const float* __restrict__ h_prev,
                                    const long long* __restrict__ lengths, int t, float* __restrict__ dh_prev,
                                    __nv_bfloat16* __restrict__ dgi, long lddgi, __nv_bfloat16* __restrict__ dgh,
                                    long lddgh, float* __restrict__ db_ih, float* __restrict__ db_hh, long R, int Hh,
                                    int UB, int rows_per_block) {
  // thread = hidden unit c (coalesced across c), loop over this block's rows; the bias gradients (column sums of the
  // gate gradients) are accumulated in registers and flushed with one atomic per (block, column)
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Hh) return;
  const long r0 = (long)blockIdx.y * rows_per_block;
  const long r1 = r0 + rows_per_block < R ? r0 + rows_per_block : R;
  float s_r = 0.f, s_z = 0.f, s_n = 0.f, s_h = 0.f;
  const int goff = (c / UB) * 4 * UB + (c % UB);
  for (long r = r0; r < r1; ++r) {
    const long idx = r * Hh + c;
    const bool live = lengths[r] > t;
    float d = dh ? dh[idx] : 0.f;
    float dar = 0.f, daz = 0.f, dan = 0.f, dhn = 0.f, dprev = d;
    if (live) {
      if (dout_t) d += dout_t[r * lddout + c];
      const __nv_bfloat16* gp = gates + r * 4L * Hh + goff;
      const float gr = __bfloat162float(gp[0]), gz = __bfloat162float(gp[UB]), gn = __bfloat162float(gp[2 * UB]);
      const float hnn = __bfloat162float(gp[3 * UB]);
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
    s_r += dar; s_z += daz; s_n += dan; s_h += dhn;
  }
  if (db_ih) {
    atomicAdd(&db_ih[c], s_r); atomicAdd(&db_ih[Hh + c], s_z); atomicAdd(&db_ih[2 * Hh + c], s_n);
  }
  if (db_hh) {
    atomicAdd(&db_hh[c], s_r); atomicAdd(&db_hh[Hh + c], s_z); atomicAdd(&db_hh[2 * Hh + c], s_h);
  }
}

int gru_gate_bwd(const float* dh, const float* dout_t, long lddout, const void* gates, const float* h_prev,
                 const long long* lengths, int t, float* dh_prev, void* dgi_bf16, long lddgi, void* dgh_bf16,
                 long lddgh, float* db_ih, float* db_hh, long R, int Hh, cudaStream_t st) {
  if (R == 0) return GTOS_OK;
  GTOS_REQUIRE(Hh % 16 == 0, "gru_gate_bwd: hidden size must be a multiple of 16");
  const int UB = (Hh % 64 == 0) ? 64 : 16;
  const int thr = Hh < 256 ? ((Hh + 31) / 32 * 32) : 256;
  const int rpb = 32;
  dim3 grid((Hh + thr - 1) / thr, (unsigned)((R + rpb - 1) / rpb));
  gru_gate_bwd_kernel<<<grid, thr, 0, st>>>(dh, dout_t, lddout, reinterpret_cast<const __nv_bfloat16*>(gates), h_prev,
                                            lengths, t, dh_prev, reinterpret_cast<__nv_bfloat16*>(dgi_bf16), lddgi,
                                            reinterpret_cast<__nv_bfloat16*>(dgh_bf16), lddgh, db_ih, db_hh, R, Hh, UB, rpb);
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
