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
  GTOS_PDL_PROLOGUE();
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
  GTOS_KLAUNCH(gru_weight_prep_kernel, dim3(4 * H), dim3(128), 0, st, w_ih, w_hh, b_ih, b_hh, Kin, H, Kx, UB,
                                                reinterpret_cast<__nv_bfloat16*>(Wcat), ldw, bcat);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// dh_tot = dh + dout_t ;  n,z,r chain rule ; dh_prev = dh_tot * z (the W_hh^T dgh term is added by a
// following GEMM with accumulate) ; finished rows pass dh through untouched and emit zero gate grads.
// gates: bf16 [R, 4H] as saved by the fused step kernel (blocks of UB units: [r | z | n | W_hn h + b_hn]).
__global__ void __launch_bounds__(256) gru_gate_bwd_kernel(
    const float* __restrict__ dh, const float* __restrict__ dout_t, long lddout, const __nv_bfloat16* __restrict__ gates,
    const float* __restrict__ h_prev, const long long* __restrict__ lengths, int t, float* __restrict__ dh_prev,
    __nv_bfloat16* __restrict__ dgi, long lddgi, __nv_bfloat16* __restrict__ dgh, long lddgh, float* __restrict__ db_ih,
    float* __restrict__ db_hh, long R, int Hh, int UB, int rows_per_block) {
  GTOS_PDL_PROLOGUE();
  // lane = 8 consecutive hidden units (16-byte bf16 / 32-byte fp32 accesses), warp = one row at a time, 8 rows of the
  // block in flight; the bias gradients (column sums of the gate gradients) are accumulated in registers, combined
  // across the block's warps in shared memory and flushed with one atomic per (block, column)
  __shared__ float s_sum[8][4][256 + 8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + lane) * 8;               // first hidden unit of this lane
  const bool on = c0 < Hh;
  const long r0 = (long)blockIdx.y * rows_per_block;
  const long r1 = r0 + rows_per_block < R ? r0 + rows_per_block : R;
  float s_r[8], s_z[8], s_n[8], s_h[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) s_r[u] = s_z[u] = s_n[u] = s_h[u] = 0.f;
  const int goff = on ? (c0 / UB) * 4 * UB + (c0 % UB) : 0;
  if (on) {
    for (long r = r0 + wib; r < r1; r += 8) {
      const long idx = r * Hh + c0;
      const bool live = lengths[r] > t;
      float d[8], dar[8], daz[8], dan[8], dhn[8], dprev[8];
      {
        const float4 a = dh ? *reinterpret_cast<const float4*>(dh + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 b = dh ? *reinterpret_cast<const float4*>(dh + idx + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { dar[u] = daz[u] = dan[u] = dhn[u] = 0.f; dprev[u] = d[u]; }
      if (live) {
        if (dout_t) {
          const float4 a = *reinterpret_cast<const float4*>(dout_t + r * lddout + c0);
          const float4 b = *reinterpret_cast<const float4*>(dout_t + r * lddout + c0 + 4);
          d[0] += a.x; d[1] += a.y; d[2] += a.z; d[3] += a.w; d[4] += b.x; d[5] += b.y; d[6] += b.z; d[7] += b.w;
        }
        const __nv_bfloat16* gp = gates + r * 4L * Hh + goff;
        const uint4 vr = *reinterpret_cast<const uint4*>(gp), vz = *reinterpret_cast<const uint4*>(gp + UB);
        const uint4 vn = *reinterpret_cast<const uint4*>(gp + 2 * UB), vh = *reinterpret_cast<const uint4*>(gp + 3 * UB);
        const float4 ha = *reinterpret_cast<const float4*>(h_prev + idx);
        const float4 hb = *reinterpret_cast<const float4*>(h_prev + idx + 4);
        const float hp[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        const __nv_bfloat162* pr = reinterpret_cast<const __nv_bfloat162*>(&vr);
        const __nv_bfloat162* pz = reinterpret_cast<const __nv_bfloat162*>(&vz);
        const __nv_bfloat162* pn = reinterpret_cast<const __nv_bfloat162*>(&vn);
        const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&vh);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fr = __bfloat1622float2(pr[q]), fz = __bfloat1622float2(pz[q]);
          const float2 fn = __bfloat1622float2(pn[q]), fh = __bfloat1622float2(ph[q]);
          const float gr2[2] = {fr.x, fr.y}, gz2[2] = {fz.x, fz.y}, gn2[2] = {fn.x, fn.y}, hn2[2] = {fh.x, fh.y};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int u = 2 * q + e;
            const float gr = gr2[e], gz = gz2[e], gn = gn2[e], hnn = hn2[e];
            const float dn = d[u] * (1.f - gz);
            const float dz = d[u] * (hp[u] - gn);
            dan[u] = dn * (1.f - gn * gn);
            daz[u] = dz * gz * (1.f - gz);
            dar[u] = dan[u] * hnn * gr * (1.f - gr);
            dhn[u] = dan[u] * gr;
            dprev[u] = d[u] * gz;
          }
        }
      }
      *reinterpret_cast<float4*>(dh_prev + idx) = make_float4(dprev[0], dprev[1], dprev[2], dprev[3]);
      *reinterpret_cast<float4*>(dh_prev + idx + 4) = make_float4(dprev[4], dprev[5], dprev[6], dprev[7]);
#define GTOS_PK8(v) make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]))
      const uint4 ur = GTOS_PK8(dar), uz = GTOS_PK8(daz), un = GTOS_PK8(dan), uh = GTOS_PK8(dhn);
#undef GTOS_PK8
      *reinterpret_cast<uint4*>(dgi + r * lddgi + c0) = ur;
      *reinterpret_cast<uint4*>(dgi + r * lddgi + Hh + c0) = uz;
      *reinterpret_cast<uint4*>(dgi + r * lddgi + 2 * Hh + c0) = un;
      *reinterpret_cast<uint4*>(dgh + r * lddgh + c0) = ur;
      *reinterpret_cast<uint4*>(dgh + r * lddgh + Hh + c0) = uz;
      *reinterpret_cast<uint4*>(dgh + r * lddgh + 2 * Hh + c0) = uh;
#pragma unroll
      for (int u = 0; u < 8; ++u) { s_r[u] += dar[u]; s_z[u] += daz[u]; s_n[u] += dan[u]; s_h[u] += dhn[u]; }
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    s_sum[wib][0][lane * 8 + u] = s_r[u];
    s_sum[wib][1][lane * 8 + u] = s_z[u];
    s_sum[wib][2][lane * 8 + u] = s_n[u];
    s_sum[wib][3][lane * 8 + u] = s_h[u];
  }
  __syncthreads();
  // thread = one of the block's 256 columns
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < Hh) {
    float tot[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += s_sum[w][k][threadIdx.x];
      tot[k] = v;
    }
    if (db_ih) {
      atomicAdd(&db_ih[c], tot[0]); atomicAdd(&db_ih[Hh + c], tot[1]); atomicAdd(&db_ih[2 * Hh + c], tot[2]);
    }
    if (db_hh) {
      atomicAdd(&db_hh[c], tot[0]); atomicAdd(&db_hh[Hh + c], tot[1]); atomicAdd(&db_hh[2 * Hh + c], tot[3]);
    }
  }
}

int gru_gate_bwd(const float* dh, const float* dout_t, long lddout, const void* gates, const float* h_prev,
                 const long long* lengths, int t, float* dh_prev, void* dgi_bf16, long lddgi, void* dgh_bf16,
                 long lddgh, float* db_ih, float* db_hh, long R, int Hh, cudaStream_t st) {
  if (R == 0) return GTOS_OK;
  GTOS_REQUIRE(Hh % 16 == 0 && lddgi % 8 == 0 && lddgh % 8 == 0 && (!dout_t || lddout % 4 == 0),
               "gru_gate_bwd: hidden size must be a multiple of 16 and the row strides 16-byte aligned");
  const int UB = (Hh % 64 == 0) ? 64 : 16;
  const int rpb = 64;
  dim3 grid((Hh + 255) / 256, (unsigned)((R + rpb - 1) / rpb));
  GTOS_KLAUNCH(gru_gate_bwd_kernel, dim3(grid), dim3(256), 0, st, dh, dout_t, lddout, reinterpret_cast<const __nv_bfloat16*>(gates), h_prev,
                                            lengths, t, dh_prev, reinterpret_cast<__nv_bfloat16*>(dgi_bf16), lddgi,
                                            reinterpret_cast<__nv_bfloat16*>(dgh_bf16), lddgh, db_ih, db_hh, R, Hh, UB, rpb);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// out[n, :] = dropout(table[idx[n], :]) ; bf16 copy is zero-padded to ldb columns
__global__ void embed_gather_kernel(const float* __restrict__ table, const long long* __restrict__ idx, long n, int dim,
                                    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, long ldb,
                                    float p, const unsigned long long* seed_ptr, unsigned long long seed_off) {
  GTOS_PDL_PROLOGUE();
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
  GTOS_KLAUNCH(embed_gather_kernel, dim3((unsigned)blocks), dim3(256), 0, st, table, idx, n, dim, out_f32,
                                                        reinterpret_cast<__nv_bfloat16*>(out_bf16), ldb, p_drop,
                                                        reinterpret_cast<const unsigned long long*>(seed_ptr), seed_off);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

__global__ void embed_scatter_kernel(const float* __restrict__ dx, long lddx, const long long* __restrict__ idx, long n,
                                     int dim, float* __restrict__ dtable, float p, const unsigned long long* seed_ptr,
                                     unsigned long long seed_off) {
  GTOS_PDL_PROLOGUE();
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
  GTOS_KLAUNCH(embed_scatter_kernel, dim3((unsigned)blocks), dim3(256), 0, st, dx, dim, idx, n, dim, dtable, p_drop,
                                                         reinterpret_cast<const unsigned long long*>(seed_ptr), seed_off);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

}  // namespace gtos
