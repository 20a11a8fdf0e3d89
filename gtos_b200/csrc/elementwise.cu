// gtos_b200 -- memory-bound helper kernels: casts, weight preparation, residual+LayerNorm
// (forward / backward), column sums, ReLU / dropout backward, relation dq/dk reduction.
// All are HBM/L2-bound streaming kernels: vectorised, coalesced, one pass.
#include "elementwise.cuh"

namespace gtos {

// ---------------------------------------------------------------------------------------
// fp32 -> bf16 with optional row padding (dst row stride ldd >= cols, pad columns zeroed)
// ---------------------------------------------------------------------------------------
__global__ void cast_pad_kernel(const float* __restrict__ src, long lds, __nv_bfloat16* __restrict__ dst, long ldd,
                                long rows, int cols) {
  GTOS_PDL_PROLOGUE();
  const long total = rows * (ldd / 2);
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long r = idx / (ldd / 2);
    const int c = (int)(idx % (ldd / 2)) * 2;
    const float* s = src + r * lds + c;
    float a = c < cols ? s[0] : 0.f;
    float b = c + 1 < cols ? s[1] : 0.f;
    *reinterpret_cast<uint32_t*>(dst + r * ldd + c) = pack_bf16x2(a, b);
  }
}

__global__ void cast_vec_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long n4) {
  GTOS_PDL_PROLOGUE();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 v = src[i];
    dst[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

int cast_f32_bf16(const float* src, long lds, void* dst, long ldd, long rows, int cols, cudaStream_t st) {
  if (rows == 0 || cols == 0) return GTOS_OK;
  GTOS_REQUIRE(ldd % 2 == 0 && ldd >= cols, "cast: destination row stride must be even and >= cols");
  if (lds == cols && ldd == cols && cols % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    long n4 = rows * cols / 4;
    int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    GTOS_KLAUNCH(cast_vec_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), n4);
  } else {
    long total = rows * (ldd / 2);
    int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    GTOS_KLAUNCH(cast_pad_kernel, dim3(blocks), dim3(256), 0, st, src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols);
  }
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// fp32 [rows, cols] -> bf16 copy AND column sums in the same pass (every Linear backward needs both:
// the bf16 operand of the two gradient GEMMs and the bias gradient)
__global__ void cast_colsum_kernel(const float* __restrict__ src, long lds, __nv_bfloat16* __restrict__ dst, long ldd,
                                   float* __restrict__ sums, long rows, int cols, int rows_per_block) {
  GTOS_PDL_PROLOGUE();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ldd) return;
  const long r0 = (long)blockIdx.y * rows_per_block;
  const long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f;
  if (c < cols) {
    for (long r = r0; r < r1; ++r) {
      const float v = src[r * lds + c];
      dst[r * ldd + c] = __float2bfloat16(v);
      s += v;
    }
    atomicAdd(&sums[c], s);
  } else {
    for (long r = r0; r < r1; ++r) dst[r * ldd + c] = __float2bfloat16(0.f);
  }
}

// 4 columns per thread (16-byte loads, 8-byte stores), 4 rows in flight per thread; blockDim.y row groups per CTA whose
// partial sums meet in shared memory, so a CTA issues ONE atomic per column whatever its depth (more CTAs instead -
// 8 rows each - measured 0.4 ms per step SLOWER: 480 atomics per column address serialise in L2)
__global__ void cast_colsum_vec4_kernel(const float* __restrict__ src, long lds, __nv_bfloat16* __restrict__ dst, long ldd,
                                        float* __restrict__ sums, long rows, int cols, int rows_per_block) {
  GTOS_PDL_PROLOGUE();
  __shared__ float4 red[3][128];
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int ty = threadIdx.y, ny = blockDim.y;
  const int rpg = rows_per_block / ny;                       // rows per row group
  const long rb = (long)blockIdx.y * rows_per_block;
  const long r0 = rb + (long)ty * rpg;
  const long r1 = r0 + rpg < rows ? r0 + rpg : rows;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) {     // cols % 4 == 0: a thread's 4 columns are all inside or all outside
    long r = r0;
    for (; r + 4 <= r1; r += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(src + (r + u) * lds + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        *reinterpret_cast<uint2*>(dst + (r + u) * ldd + c) = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
        s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
      }
    }
    for (; r < r1; ++r) {
      const float4 v = *reinterpret_cast<const float4*>(src + r * lds + c);
      *reinterpret_cast<uint2*>(dst + r * ldd + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  } else if (c < ldd) {
    for (long r = r0; r < r1; ++r) *reinterpret_cast<uint2*>(dst + r * ldd + c) = make_uint2(0u, 0u);
  }
  if (ny > 1) {
    if (ty > 0) red[ty - 1][threadIdx.x] = s;
    __syncthreads();
    if (ty == 0)
      for (int g = 0; g < ny - 1; ++g) {
        const float4 o = red[g][threadIdx.x];
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
      }
  }
  if (ty == 0 && c < cols) {
    atomicAdd(&sums[c], s.x); atomicAdd(&sums[c + 1], s.y); atomicAdd(&sums[c + 2], s.z); atomicAdd(&sums[c + 3], s.w);
  }
}

int cast_colsum(const float* src, long lds, void* dst, long ldd, float* sums, long rows, int cols, cudaStream_t st) {
  GTOS_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * cols, st));
  if (rows == 0 || cols == 0) return GTOS_OK;
  GTOS_REQUIRE(ldd >= cols, "cast_colsum: destination row stride must be >= cols");
  if (cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    const int rpb = 32;
    static const int ny_env = getenv("GTOS_CC_NY") ? atoi(getenv("GTOS_CC_NY")) : 4;   // 1, 2 or 4 row groups per CTA
    const int ny = (ny_env == 1 || ny_env == 2) ? ny_env : 4;                          // anything else: the default
    dim3 grid((unsigned)((ldd / 4 + 127) / 128), (unsigned)((rows + rpb - 1) / rpb));
    GTOS_KLAUNCH(cast_colsum_vec4_kernel, dim3(grid), dim3(128, ny), 0, st, src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, sums, rows, cols, rpb);
    GTOS_LAUNCH_CHECK();
    return GTOS_OK;
  }
  const int rpb = 32;
  dim3 grid((unsigned)((ldd + 127) / 128), (unsigned)((rows + rpb - 1) / rpb));
  GTOS_KLAUNCH(cast_colsum_kernel, dim3(grid), dim3(128), 0, st, src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, sums, rows, cols, rpb);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// weight prep: W f32 [R,C] -> Wb bf16 [R,ldw] (optional) and Wt bf16 [C,ldt] (optional)
// optional row permutation for relation_in_proj (perm_D > 0): output row pr <- input row orig(pr)
// ---------------------------------------------------------------------------------------
__global__ void weight_prep_kernel(const float* __restrict__ W, int R, int C, __nv_bfloat16* __restrict__ Wb, long ldw,
                                   __nv_bfloat16* __restrict__ Wt, long ldt, int perm_D, int perm_hd) {
  GTOS_PDL_PROLOGUE();
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    int pr = r0 + y, c = c0 + threadIdx.x;
    float v = 0.f;
    if (pr < R && c < C) {
      int ro = perm_D ? rel_perm_to_orig(pr, perm_D, perm_hd) : pr;
      v = W[(long)ro * C + c];
    }
    tile[y][threadIdx.x] = v;
    if (Wb && pr < R && c < ldw) Wb[(long)pr * ldw + c] = __float2bfloat16(v);
  }
  __syncthreads();
  if (Wt) {
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
      int c = c0 + y, pr = r0 + threadIdx.x;
      if (c < C && pr < ldt) Wt[(long)c * ldt + pr] = __float2bfloat16(pr < R ? tile[threadIdx.x][y] : 0.f);
    }
  }
}

int weight_prep(const float* W, int R, int C, void* Wb, long ldw, void* Wt, long ldt, int perm_D, int perm_hd,
                cudaStream_t st) {
  // grid covers the padded extents so pad columns/rows are written as zeros
  int cx = (int)(((Wb ? (ldw > C ? ldw : C) : C) + 31) / 32);
  int ry = (int)(((Wt ? (ldt > R ? ldt : R) : R) + 31) / 32);
  dim3 grid(cx, ry), block(32, 8);
  GTOS_KLAUNCH(weight_prep_kernel, dim3(grid), dim3(block), 0, st, W, R, C, reinterpret_cast<__nv_bfloat16*>(Wb), ldw,
                                             reinterpret_cast<__nv_bfloat16*>(Wt), ldt, perm_D, perm_hd);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// z = res + dropout(x);  y = LayerNorm(z) * gamma + beta       (one warp per row, D <= 1024)
// reference: generator/graph_transformer.py:57-58,64-65; transformer.py:56-57,63,70-71
// ---------------------------------------------------------------------------------------
constexpr int LN_MAX_PER_LANE = 32;

__global__ void add_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ z_out,
                                  float* __restrict__ mean_out, float* __restrict__ rstd_out, long rows, int D,
                                  float p_drop, const unsigned long long* __restrict__ seed_ptr,
                                  unsigned long long seed_off, float eps) {
  GTOS_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const unsigned long long seed = (p_drop > 0.f) ? (seed_ptr[0] + seed_off) : 0ull;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float v[LN_MAX_PER_LANE];
  float s = 0.f;
  const int per = (D + 31) / 32;
#pragma unroll
  for (int t = 0; t < LN_MAX_PER_LANE; ++t) {
    if (t < per) {
      int c = t * 32 + lane;
      float a = 0.f;
      if (c < D) {
        a = x[row * D + c];
        if (p_drop > 0.f) a = (rng_uniform(seed, (unsigned long long)(row * D + c)) >= p_drop) ? a * keep_scale : 0.f;
        if (res) a += res[row * D + c];
      }
      v[t] = a;
      s += a;
    }
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int t = 0; t < LN_MAX_PER_LANE; ++t) {
    if (t < per) {
      int c = t * 32 + lane;
      float d = c < D ? v[t] - mean : 0.f;
      q += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
  for (int t = 0; t < LN_MAX_PER_LANE; ++t) {
    if (t < per) {
      int c = t * 32 + lane;
      if (c < D) {
        float o = (v[t] - mean) * rstd * gamma[c] + beta[c];
        y[row * D + c] = o;
        if (y_bf16) y_bf16[row * D + c] = __float2bfloat16(o);
        if (z_out) z_out[row * D + c] = v[t];
      }
    }
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// D % 128 == 0: every lane owns NV float4 groups of its row (16-byte accesses; the scalar kernel above issues 4x as many
// memory instructions and runs at < 1.5 TB/s).  Dropout uses the same per-element counter as the scalar path.
template <int NV>
__global__ void __launch_bounds__(128) add_ln_fwd_vec_kernel(
    const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ z_out,
    float* __restrict__ mean_out, float* __restrict__ rstd_out, long rows, float p_drop,
    const unsigned long long* __restrict__ seed_ptr, unsigned long long seed_off, float eps) {
  GTOS_PDL_PROLOGUE();
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const unsigned long long seed = (p_drop > 0.f) ? (seed_ptr[0] + seed_off) : 0ull;
  const float ks = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const float4* x4 = reinterpret_cast<const float4*>(x + row * D);
  const float4* r4 = res ? reinterpret_cast<const float4*>(res + row * D) : nullptr;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int c4 = t * 32 + lane;
    float4 a = x4[c4];
    if (p_drop > 0.f) {
      const unsigned long long e = (unsigned long long)(row * D + c4 * 4);
      a.x = (rng_uniform(seed, e) >= p_drop) ? a.x * ks : 0.f;
      a.y = (rng_uniform(seed, e + 1) >= p_drop) ? a.y * ks : 0.f;
      a.z = (rng_uniform(seed, e + 2) >= p_drop) ? a.z * ks : 0.f;
      a.w = (rng_uniform(seed, e + 3) >= p_drop) ? a.w * ks : 0.f;
    }
    if (r4) {
      const float4 r = r4[c4];
      a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
    }
    v[t] = a;
    s += (a.x + a.y) + (a.z + a.w);
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const float d0 = v[t].x - mean, d1 = v[t].y - mean, d2 = v[t].z - mean, d3 = v[t].w - mean;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int c4 = t * 32 + lane;
    const float4 g = g4[c4], bb = b4[c4];
    float4 o;
    o.x = (v[t].x - mean) * rstd * g.x + bb.x;
    o.y = (v[t].y - mean) * rstd * g.y + bb.y;
    o.z = (v[t].z - mean) * rstd * g.z + bb.z;
    o.w = (v[t].w - mean) * rstd * g.w + bb.w;
    reinterpret_cast<float4*>(y + row * D)[c4] = o;
    if (y_bf16) reinterpret_cast<uint2*>(y_bf16 + row * D)[c4] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    if (z_out) reinterpret_cast<float4*>(z_out + row * D)[c4] = v[t];
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int add_ln_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y, void* y_bf16,
               float* z, float* mean, float* rstd, long rows, int D, float p_drop, const void* seed_ptr,
               unsigned long long seed_off, cudaStream_t st) {
  GTOS_REQUIRE(D <= 32 * LN_MAX_PER_LANE, "LayerNorm width %d > %d unsupported", D, 32 * LN_MAX_PER_LANE);
  GTOS_REQUIRE(p_drop == 0.f || seed_ptr, "dropout needs a device seed pointer");
  if (rows == 0) return GTOS_OK;
  if (D % 128 == 0 && al16(x) && al16(res) && al16(gamma) && al16(beta) && al16(y) && al16(y_bf16) && al16(z)) {
    const unsigned blocks = (unsigned)((rows + 3) / 4);
    __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(y_bf16);
    const unsigned long long* sp = reinterpret_cast<const unsigned long long*>(seed_ptr);
#define GTOS_LN_FWD(NV)                                                                                               \
  GTOS_KLAUNCH(add_ln_fwd_vec_kernel<NV>, dim3(blocks), dim3(128), 0, st, x, res, gamma, beta, y, yb, z, mean, rstd, rows, p_drop, sp,      \
                                                    seed_off, 1e-5f)
    switch (D / 128) {
      case 1: GTOS_LN_FWD(1); break;
      case 2: GTOS_LN_FWD(2); break;
      case 3: GTOS_LN_FWD(3); break;
      case 4: GTOS_LN_FWD(4); break;
      case 5: GTOS_LN_FWD(5); break;
      case 6: GTOS_LN_FWD(6); break;
      case 7: GTOS_LN_FWD(7); break;
      default: GTOS_LN_FWD(8); break;
    }
#undef GTOS_LN_FWD
    GTOS_LAUNCH_CHECK();
    return GTOS_OK;
  }
  const int wpb = 4;
  GTOS_KLAUNCH(add_ln_fwd_kernel, dim3((unsigned)((rows + wpb - 1) / wpb)), dim3(wpb * 32), 0, st, 
      x, res, gamma, beta, y, reinterpret_cast<__nv_bfloat16*>(y_bf16), z, mean, rstd, rows, D, p_drop,
      reinterpret_cast<const unsigned long long*>(seed_ptr), seed_off, 1e-5f);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// backward: given dy, z (pre-norm), mean, rstd, gamma:
//   xhat = (z-mean)*rstd ; g = dy*gamma ; dz = rstd*(g - mean(g) - xhat*mean(g*xhat))
//   dres = dz ; dx = dz * dropmask/(1-p) ; dgamma += dy*xhat ; dbeta += dy   (atomics into zeroed [D])
__global__ void add_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                  const float* __restrict__ gamma, float* __restrict__ dres, float* __restrict__ dx,
                                  __nv_bfloat16* __restrict__ dx_bf16, long rows, int D, float p_drop,
                                  const unsigned long long* __restrict__ seed_ptr, unsigned long long seed_off) {
  GTOS_PDL_PROLOGUE();
  // one warp per row (many warps in flight: the kernel is latency-bound on three dependent passes per row)
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const unsigned long long seed = (p_drop > 0.f) ? (seed_ptr[0] + seed_off) : 0ull;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const int per = (D + 31) / 32;
  const float mu = mean[row], rs = rstd[row];
  float g[LN_MAX_PER_LANE], xh[LN_MAX_PER_LANE];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int t = 0; t < LN_MAX_PER_LANE; ++t) {
    if (t < per) {
      int c = t * 32 + lane;
      float gg = 0.f, xx = 0.f;
      if (c < D) {
        xx = (z[row * D + c] - mu) * rs;
        gg = dy[row * D + c] * gamma[c];
      }
      g[t] = gg;
      xh[t] = xx;
      s1 += gg;
      s2 += gg * xx;
    }
  }
  s1 = warp_sum(s1) / D;
  s2 = warp_sum(s2) / D;
#pragma unroll
  for (int t = 0; t < LN_MAX_PER_LANE; ++t) {
    if (t < per) {
      int c = t * 32 + lane;
      if (c < D) {
        float dz = rs * (g[t] - s1 - xh[t] * s2);
        if (dres) dres[row * D + c] = dz;
        float dxx = dz;
        if (p_drop > 0.f) dxx = (rng_uniform(seed, (unsigned long long)(row * D + c)) >= p_drop) ? dz * keep_scale : 0.f;
        if (dx) dx[row * D + c] = dxx;
        if (dx_bf16) dx_bf16[row * D + c] = __float2bfloat16(dxx);
      }
    }
  }
}

// dgamma[c] = sum_rows dy * xhat, dbeta[c] = sum_rows dy : coalesced column reduction, one atomic per (column, row block)
__global__ void ln_param_grad_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, long rows, int D,
                                     int rows_per_block) {
  GTOS_PDL_PROLOGUE();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  long r0 = (long)blockIdx.y * rows_per_block;
  long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float ag = 0.f, ab = 0.f;
#pragma unroll 8
  for (long r = r0; r < r1; ++r) {
    const float d = dy[r * D + c];
    ag += d * (z[r * D + c] - mean[r]) * rstd[r];
    ab += d;
  }
  atomicAdd(&dgamma[c], ag);
  atomicAdd(&dbeta[c], ab);
}

template <int NV>
__global__ void __launch_bounds__(128) add_ln_bwd_vec_kernel(
    const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, float* __restrict__ dres, float* __restrict__ dx,
    __nv_bfloat16* __restrict__ dx_bf16, long rows, float p_drop, const unsigned long long* __restrict__ seed_ptr,
    unsigned long long seed_off) {
  GTOS_PDL_PROLOGUE();
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const unsigned long long seed = (p_drop > 0.f) ? (seed_ptr[0] + seed_off) : 0ull;
  const float ks = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const float mu = mean[row], rs = rstd[row];
  const float4* z4 = reinterpret_cast<const float4*>(z + row * D);
  const float4* d4 = reinterpret_cast<const float4*>(dy + row * D);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4 g[NV], xh[NV];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int c4 = t * 32 + lane;
    const float4 zz = z4[c4], dd = d4[c4], gm = g4[c4];
    xh[t] = make_float4((zz.x - mu) * rs, (zz.y - mu) * rs, (zz.z - mu) * rs, (zz.w - mu) * rs);
    g[t] = make_float4(dd.x * gm.x, dd.y * gm.y, dd.z * gm.z, dd.w * gm.w);
    s1 += (g[t].x + g[t].y) + (g[t].z + g[t].w);
    s2 += (g[t].x * xh[t].x + g[t].y * xh[t].y) + (g[t].z * xh[t].z + g[t].w * xh[t].w);
  }
  s1 = warp_sum(s1) / D;
  s2 = warp_sum(s2) / D;
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    const int c4 = t * 32 + lane;
    float4 dz;
    dz.x = rs * (g[t].x - s1 - xh[t].x * s2);
    dz.y = rs * (g[t].y - s1 - xh[t].y * s2);
    dz.z = rs * (g[t].z - s1 - xh[t].z * s2);
    dz.w = rs * (g[t].w - s1 - xh[t].w * s2);
    if (dres) reinterpret_cast<float4*>(dres + row * D)[c4] = dz;
    float4 dd = dz;
    if (p_drop > 0.f) {
      const unsigned long long e = (unsigned long long)(row * D + c4 * 4);
      dd.x = (rng_uniform(seed, e) >= p_drop) ? dz.x * ks : 0.f;
      dd.y = (rng_uniform(seed, e + 1) >= p_drop) ? dz.y * ks : 0.f;
      dd.z = (rng_uniform(seed, e + 2) >= p_drop) ? dz.z * ks : 0.f;
      dd.w = (rng_uniform(seed, e + 3) >= p_drop) ? dz.w * ks : 0.f;
    }
    if (dx) reinterpret_cast<float4*>(dx + row * D)[c4] = dd;
    if (dx_bf16) reinterpret_cast<uint2*>(dx_bf16 + row * D)[c4] = make_uint2(pack_bf16x2(dd.x, dd.y), pack_bf16x2(dd.z, dd.w));
  }
}

// dgamma / dbeta over all rows (one thread per column, rows_per_block rows per CTA, atomics into the zeroed outputs).  A
// separate entry point so the host can run it on a second stream: nothing on the critical path of a backward pass reads it.
int ln_param_grad(const float* dy, const float* z, const float* mean, const float* rstd, float* dgamma, float* dbeta,
                  long rows, int D, cudaStream_t st) {
  GTOS_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * D, st));
  GTOS_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * D, st));
  if (rows == 0) return GTOS_OK;
  const int rpb = 32;   // thread per column: 4 x more threads in flight than a float4-per-thread variant, which measured slower
  dim3 grid((D + 127) / 128, (unsigned)((rows + rpb - 1) / rpb));
  GTOS_KLAUNCH(ln_param_grad_kernel, dim3(grid), dim3(128), 0, st, dy, z, mean, rstd, dgamma, dbeta, rows, D, rpb);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// dgamma == nullptr: input gradients only (the caller computes the parameter gradients with ln_param_grad)
int add_ln_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma, float* dres,
               float* dx, void* dx_bf16, float* dgamma, float* dbeta, long rows, int D, float p_drop,
               const void* seed_ptr, unsigned long long seed_off, cudaStream_t st) {
  GTOS_REQUIRE(D <= 32 * LN_MAX_PER_LANE, "LayerNorm width %d unsupported", D);
  GTOS_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "add_ln_bwd: dgamma and dbeta go together");
  if (rows > 0) {
    if (D % 128 == 0 && al16(dy) && al16(z) && al16(gamma) && al16(dres) && al16(dx) && al16(dx_bf16)) {
      const unsigned blocks = (unsigned)((rows + 3) / 4);
      __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
      const unsigned long long* sp = reinterpret_cast<const unsigned long long*>(seed_ptr);
#define GTOS_LN_BWD(NV) GTOS_KLAUNCH(add_ln_bwd_vec_kernel<NV>, dim3(blocks), dim3(128), 0, st, dy, z, mean, rstd, gamma, dres, dx, xb, rows, p_drop, sp, seed_off)
      switch (D / 128) {
        case 1: GTOS_LN_BWD(1); break;
        case 2: GTOS_LN_BWD(2); break;
        case 3: GTOS_LN_BWD(3); break;
        case 4: GTOS_LN_BWD(4); break;
        case 5: GTOS_LN_BWD(5); break;
        case 6: GTOS_LN_BWD(6); break;
        case 7: GTOS_LN_BWD(7); break;
        default: GTOS_LN_BWD(8); break;
      }
#undef GTOS_LN_BWD
      GTOS_LAUNCH_CHECK();
    } else {
      const int wpb = 4;
      GTOS_KLAUNCH(add_ln_bwd_kernel, dim3((unsigned)((rows + wpb - 1) / wpb)), dim3(wpb * 32), 0, st,
                   dy, z, mean, rstd, gamma, dres, dx, reinterpret_cast<__nv_bfloat16*>(dx_bf16), rows, D, p_drop,
                   reinterpret_cast<const unsigned long long*>(seed_ptr), seed_off);
      GTOS_LAUNCH_CHECK();
    }
  }
  if (dgamma) return ln_param_grad(dy, z, mean, rstd, dgamma, dbeta, rows, D, st);
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// column sums: out[n] = sum_m x[m, n]  (bias gradients)
// ---------------------------------------------------------------------------------------
__global__ void colsum_kernel(const float* __restrict__ x, long ld, float* __restrict__ out, long rows, int cols,
                              int rows_per_block) {
  GTOS_PDL_PROLOGUE();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  long r0 = (long)blockIdx.y * rows_per_block;
  long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f;
#pragma unroll 8
  for (long r = r0; r < r1; ++r) s += x[r * ld + c];
  atomicAdd(&out[c], s);
}

__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long ld, float* __restrict__ out, long rows,
                                   int cols, int rows_per_block) {
  GTOS_PDL_PROLOGUE();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  long r0 = (long)blockIdx.y * rows_per_block;
  long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f;
#pragma unroll 8
  for (long r = r0; r < r1; ++r) s += __bfloat162float(x[r * ld + c]);
  atomicAdd(&out[c], s);
}

int colsum_bf16(const void* x, long ld, float* out, long rows, int cols, cudaStream_t st) {
  GTOS_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  if (rows == 0) return GTOS_OK;
  int rpb = 64;
  dim3 grid((cols + 127) / 128, (unsigned)((rows + rpb - 1) / rpb));
  GTOS_KLAUNCH(colsum_bf16_kernel, dim3(grid), dim3(128), 0, st, reinterpret_cast<const __nv_bfloat16*>(x), ld, out, rows, cols, rpb);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int colsum(const float* x, long ld, float* out, long rows, int cols, cudaStream_t st) {
  GTOS_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  if (rows == 0) return GTOS_OK;
  int rpb = 64;
  dim3 grid((cols + 127) / 128, (unsigned)((rows + rpb - 1) / rpb));
  GTOS_KLAUNCH(colsum_kernel, dim3(grid), dim3(128), 0, st, x, ld, out, rows, cols, rpb);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// zero up to ZERO_MAX_REGIONS byte ranges in ONE launch (the GRU's length-sorted schedule clears ~75 small row ranges per
// step - rows no kernel writes but a later GEMM reads; as separate fills they were 75 launches of ~2.3 us each, 30 of them
// in front of the first GRU step).  blockIdx.y = region, the blocks of a region stride over it with 16-byte stores.
// ---------------------------------------------------------------------------------------
__global__ void zero_regions_kernel(const ZeroRegions z) {
  GTOS_PDL_PROLOGUE();
  uint8_t* p = reinterpret_cast<uint8_t*>(z.ptr[blockIdx.y]);
  const long n = z.bytes[blockIdx.y];
  if (n <= 0) return;
  const long head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;          // bytes up to the first 16-byte boundary
  const long h = head < n ? head : n;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = tid; i < h; i += stride) p[i] = 0;
  const long n16 = (n - h) >> 4;
  uint4* q = reinterpret_cast<uint4*>(p + h);
  for (long i = tid; i < n16; i += stride) q[i] = make_uint4(0u, 0u, 0u, 0u);
  for (long i = h + (n16 << 4) + tid; i < n; i += stride) p[i] = 0;
}

int zero_regions(const ZeroRegions& z, int n, cudaStream_t st) {
  if (n <= 0) return GTOS_OK;
  GTOS_REQUIRE(n <= ZERO_MAX_REGIONS, "zero_regions: at most %d regions per call", ZERO_MAX_REGIONS);
  long mx = 0;
  for (int i = 0; i < n; ++i) mx = z.bytes[i] > mx ? z.bytes[i] : mx;
  if (mx <= 0) return GTOS_OK;
  long bx = (mx / 16 + 255) / 256;
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  GTOS_KLAUNCH(zero_regions_kernel, dim3((unsigned)bx, (unsigned)n), dim3(256), 0, st, z);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// elementwise dropout forward on a bf16 activation (FFN hidden), in place; and
// dh = dh_in * (h > 0) * dropmask/(1-p) -> bf16 (+ optional fp32)      [ReLU/dropout backward]
// ---------------------------------------------------------------------------------------
__global__ void dropout_bf16_kernel(__nv_bfloat16* __restrict__ h, long n, float p, const unsigned long long* seed_ptr,
                                    unsigned long long seed_off) {
  GTOS_PDL_PROLOGUE();
  const unsigned long long seed = seed_ptr[0] + seed_off;
  const float ks = 1.f / (1.f - p);
  const long stride = (long)gridDim.x * blockDim.x;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n8 = ((reinterpret_cast<uintptr_t>(h) & 15) == 0) ? n / 8 : 0;   // 16-byte groups
  for (long g = tid; g < n8; g += stride) {
    uint4 v = reinterpret_cast<uint4*>(h)[g];
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 f = __bfloat1622float2(h2[t]);
      const unsigned long long e = (unsigned long long)(g * 8 + 2 * t);
      f.x = rng_uniform(seed, e) >= p ? f.x * ks : 0.f;
      f.y = rng_uniform(seed, e + 1) >= p ? f.y * ks : 0.f;
      h2[t] = __floats2bfloat162_rn(f.x, f.y);
    }
    reinterpret_cast<uint4*>(h)[g] = v;
  }
  for (long i = n8 * 8 + tid; i < n; i += stride) {
    float v = __bfloat162float(h[i]);
    h[i] = __float2bfloat16(rng_uniform(seed, (unsigned long long)i) >= p ? v * ks : 0.f);
  }
}

int dropout_bf16(void* h, long n, float p, const void* seed_ptr, unsigned long long seed_off, cudaStream_t st) {
  if (p <= 0.f || n == 0) return GTOS_OK;
  GTOS_REQUIRE(seed_ptr, "dropout needs a device seed pointer");
  long blocks = (n / 8 + 255) / 256 + 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  GTOS_KLAUNCH(dropout_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, st, reinterpret_cast<__nv_bfloat16*>(h), n, p,
                                                        reinterpret_cast<const unsigned long long*>(seed_ptr), seed_off);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

__global__ void dropout_f32_kernel(const float* __restrict__ x, float* __restrict__ out, long n, float p,
                                   const unsigned long long* seed_ptr, unsigned long long seed_off) {
  GTOS_PDL_PROLOGUE();
  const unsigned long long seed = seed_ptr[0] + seed_off;
  const float ks = 1.f / (1.f - p);
  const long stride = (long)gridDim.x * blockDim.x;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n4 = (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) ? n / 4 : 0;
  for (long g = tid; g < n4; g += stride) {
    float4 v = reinterpret_cast<const float4*>(x)[g];
    const unsigned long long e = (unsigned long long)(g * 4);
    v.x = rng_uniform(seed, e) >= p ? v.x * ks : 0.f;
    v.y = rng_uniform(seed, e + 1) >= p ? v.y * ks : 0.f;
    v.z = rng_uniform(seed, e + 2) >= p ? v.z * ks : 0.f;
    v.w = rng_uniform(seed, e + 3) >= p ? v.w * ks : 0.f;
    reinterpret_cast<float4*>(out)[g] = v;
  }
  for (long i = n4 * 4 + tid; i < n; i += stride) out[i] = rng_uniform(seed, (unsigned long long)i) >= p ? x[i] * ks : 0.f;
}

int dropout_f32(const float* x, float* out, long n, float p, const void* seed_ptr, unsigned long long seed_off,
                cudaStream_t st) {
  if (n == 0) return GTOS_OK;
  GTOS_REQUIRE(p > 0.f && p < 1.f && seed_ptr, "dropout_f32: need 0 < p < 1 and a device seed pointer");
  long blocks = (n / 4 + 255) / 256 + 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  GTOS_KLAUNCH(dropout_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, st, x, out, n, p,
                                                       reinterpret_cast<const unsigned long long*>(seed_ptr), seed_off);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// act = post-dropout hidden activation actually fed to fc2 (bf16): zero where relu OR dropout killed it
__global__ void relu_drop_bwd_kernel(const float* __restrict__ dh_in, const __nv_bfloat16* __restrict__ act,
                                     float* __restrict__ dh_f32, __nv_bfloat16* __restrict__ dh_bf16, long n, float p) {
  GTOS_PDL_PROLOGUE();
  const float ks = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const long stride = (long)gridDim.x * blockDim.x;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool al = ((reinterpret_cast<uintptr_t>(dh_in) | reinterpret_cast<uintptr_t>(dh_f32)) & 15) == 0 &&
                  ((reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(dh_bf16)) & 7) == 0;
  const long n4 = al ? n / 4 : 0;
  for (long g = tid; g < n4; g += stride) {
    const float4 d = reinterpret_cast<const float4*>(dh_in)[g];
    const uint2 a = reinterpret_cast<const uint2*>(act)[g];
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const float2 a01 = __bfloat1622float2(a2[0]), a23 = __bfloat1622float2(a2[1]);
    float4 v;
    v.x = a01.x > 0.f ? d.x * ks : 0.f;
    v.y = a01.y > 0.f ? d.y * ks : 0.f;
    v.z = a23.x > 0.f ? d.z * ks : 0.f;
    v.w = a23.y > 0.f ? d.w * ks : 0.f;
    if (dh_f32) reinterpret_cast<float4*>(dh_f32)[g] = v;
    if (dh_bf16) reinterpret_cast<uint2*>(dh_bf16)[g] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  for (long i = n4 * 4 + tid; i < n; i += stride) {
    float v = (__bfloat162float(act[i]) > 0.f) ? dh_in[i] * ks : 0.f;
    if (dh_f32) dh_f32[i] = v;
    if (dh_bf16) dh_bf16[i] = __float2bfloat16(v);
  }
}

int relu_drop_bwd(const float* dh_in, const void* act, float* dh_f32, void* dh_bf16, long n, float p,
                  cudaStream_t st) {
  if (n == 0) return GTOS_OK;
  long blocks = (n / 4 + 255) / 256 + 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  GTOS_KLAUNCH(relu_drop_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, st, dh_in, reinterpret_cast<const __nv_bfloat16*>(act), dh_f32,
                                                         reinterpret_cast<__nv_bfloat16*>(dh_bf16), n, p);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// relation backward: dq[i,b,:] = sum_j G_x[(j,i,b)], dk[j,b,:] = sum_i G_y[(j,i,b)]
// G is bf16 [tiles*128, 2D] in tile-major row order and head-interleaved column order
// (per head: [d(q+ra) (hd) | d(k+rb) (hd)]).  Outputs are written into the [N*B, ld] grad buffer
// of the fused QKV projection (dq at column 0, dk at column D).
// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------
// TokenGenerator training tail (generator/decoder.py:42-64) in one pass over the vocabulary logits:
//   p[target] = gen * softmax(logits)[target] + copy * sum_s align[s] * [copy_seq[s] == target]
//   loss_row  = -log(p + 1e-12), 0 where target == pad
// replaces softmax / zero-extension cat / scatter_add / log / gather and their autograd (8 passes over [T*B, V]).
// One CTA per (t, b) row.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float r = (lane < nw) ? sh[lane] : (is_max ? -INFINITY : 0.f);   // every warp reduces the per-warp partials itself
  r = is_max ? warp_max(r) : warp_sum(r);
  return r;
}

__global__ void token_nll_fwd_kernel(const float* __restrict__ logits, long ldl, int V, const float* __restrict__ gate_logits,
                                     const float* __restrict__ align, int S, const long long* __restrict__ copy_seq,
                                     const long long* __restrict__ target, int B, long long pad_idx,
                                     float* __restrict__ loss_row, float* __restrict__ stats) {
  GTOS_PDL_PROLOGUE();
  __shared__ float sh[32];
  const long row = blockIdx.x;
  const int b = (int)(row % B);
  const float* lr = logits + row * ldl;
  float mx = -INFINITY;
  float se = 0.f;
  // the row is read ONCE, 16 bytes per load, and kept in registers between the max and the sum of exponentials (<= 12
  // float4 per thread: V <= 12288); the two scalar passes this replaces took 112 us for the [3840, 10000] logits of config 2,
  // 4.7x the HBM time of one read
  constexpr int NV = 12;
  const int n4 = V >> 2;
  if ((V & 3) == 0 && (ldl & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && n4 <= NV * (int)blockDim.x) {
    float4 reg[NV];
    const float4* l4 = reinterpret_cast<const float4*>(lr);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int q = threadIdx.x + i * blockDim.x;
      reg[i] = q < n4 ? l4[q] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      mx = fmaxf(mx, fmaxf(fmaxf(reg[i].x, reg[i].y), fmaxf(reg[i].z, reg[i].w)));
    }
    mx = block_reduce(mx, sh, true);
#pragma unroll
    for (int i = 0; i < NV; ++i)      // exp(-inf - mx) = 0 for the padding lanes
      se += __expf(reg[i].x - mx) + __expf(reg[i].y - mx) + __expf(reg[i].z - mx) + __expf(reg[i].w - mx);
    se = block_reduce(se, sh, false);
  } else {
    for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, lr[v]);
    mx = block_reduce(mx, sh, true);
    for (int v = threadIdx.x; v < V; v += blockDim.x) se += __expf(lr[v] - mx);
    se = block_reduce(se, sh, false);
  }
  const long long tgt = target[row];
  float cp = 0.f;
  for (int s = threadIdx.x; s < S; s += blockDim.x)
    if (copy_seq[(long)s * B + b] == tgt) cp += align[row * S + s];
  cp = block_reduce(cp, sh, false);
  if (threadIdx.x == 0) {
    const float g0 = gate_logits[row * 2], g1 = gate_logits[row * 2 + 1];
    const float gm = fmaxf(g0, g1);
    const float e0 = __expf(g0 - gm), e1 = __expf(g1 - gm);
    const float gen = e0 / (e0 + e1), cpy = e1 / (e0 + e1);
    const float sm_t = (tgt >= 0 && tgt < V) ? __expf(lr[tgt] - mx) / se : 0.f;
    const float p = gen * sm_t + cpy * cp;
    loss_row[row] = (tgt == pad_idx) ? 0.f : -logf(p + 1e-12f);
    stats[row * 6 + 0] = mx; stats[row * 6 + 1] = se; stats[row * 6 + 2] = gen; stats[row * 6 + 3] = cpy;
    stats[row * 6 + 4] = sm_t; stats[row * 6 + 5] = cp;
  }
}

__global__ void token_nll_bwd_kernel(const float* __restrict__ dloss_row, const float* __restrict__ logits, long ldl, int V,
                                     const float* __restrict__ align, int S, const long long* __restrict__ copy_seq,
                                     const long long* __restrict__ target, int B, long long pad_idx,
                                     const float* __restrict__ stats, float* __restrict__ dlogits, long lddl,
                                     float* __restrict__ dgate_logits, float* __restrict__ dalign,
                                     __nv_bfloat16* __restrict__ dlogits_b, long lddb) {
  GTOS_PDL_PROLOGUE();
  const long row = blockIdx.x;
  const int b = (int)(row % B);
  const long long tgt = target[row];
  const float mx = stats[row * 6], se = stats[row * 6 + 1], gen = stats[row * 6 + 2], cpy = stats[row * 6 + 3];
  const float sm_t = stats[row * 6 + 4], cp = stats[row * 6 + 5];
  const float p = gen * sm_t + cpy * cp;
  const float dp = (tgt == pad_idx) ? 0.f : -dloss_row[row] / (p + 1e-12f);   // dL/dp
  const float* lr = logits + row * ldl;
  float* dl = dlogits + row * lddl;
  const float coef = dp * gen;
  const float inv = 1.f / se;
  if (dlogits_b && (V & 1) == 0) {
    // two columns per thread: the bf16 operand copy of d logits for the vocabulary projection's backward GEMMs is made
    // here, so that layer does not re-read the fp32 [T*B, V] tensor just to cast it
    __nv_bfloat16* db = dlogits_b + row * lddb;
    for (int v = 2 * threadIdx.x; v < V; v += 2 * blockDim.x) {
      const float2 l2 = *reinterpret_cast<const float2*>(lr + v);
      const float s0 = __expf(l2.x - mx) * inv, s1 = __expf(l2.y - mx) * inv;
      const float g0 = coef * s0 * ((v == tgt ? 1.f : 0.f) - sm_t), g1 = coef * s1 * ((v + 1 == tgt ? 1.f : 0.f) - sm_t);
      *reinterpret_cast<float2*>(dl + v) = make_float2(g0, g1);
      *reinterpret_cast<uint32_t*>(db + v) = pack_bf16x2(g0, g1);
    }
  } else {
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float sm = __expf(lr[v] - mx) * inv;
      const float gv = coef * sm * ((v == tgt ? 1.f : 0.f) - sm_t);
      dl[v] = gv;
      if (dlogits_b) dlogits_b[row * lddb + v] = __float2bfloat16(gv);
    }
  }
  for (int s = threadIdx.x; s < S; s += blockDim.x)
    dalign[row * S + s] = (copy_seq[(long)s * B + b] == tgt) ? dp * cpy : 0.f;
  if (threadIdx.x == 0) {
    const float dgen = dp * sm_t, dcpy = dp * cp;
    const float dot = gen * dgen + cpy * dcpy;
    dgate_logits[row * 2] = gen * (dgen - dot);
    dgate_logits[row * 2 + 1] = cpy * (dcpy - dot);
  }
}

int token_nll_fwd(const float* logits, long ldl, int V, const float* gate_logits, const float* align, int S,
                  const long long* copy_seq, const long long* target, long rows, int B, long long pad_idx,
                  float* loss_row, float* stats, cudaStream_t st) {
  if (rows == 0) return GTOS_OK;
  GTOS_KLAUNCH(token_nll_fwd_kernel, dim3((unsigned)rows), dim3(256), 0, st, logits, ldl, V, gate_logits, align, S, copy_seq, target, B, pad_idx,
                                                       loss_row, stats);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int token_nll_bwd(const float* dloss_row, const float* logits, long ldl, int V, const float* align, int S,
                  const long long* copy_seq, const long long* target, long rows, int B, long long pad_idx,
                  const float* stats, float* dlogits, long lddl, float* dgate_logits, float* dalign, void* dlogits_bf16,
                  long lddb, cudaStream_t st) {
  if (rows == 0) return GTOS_OK;
  GTOS_REQUIRE(!dlogits_bf16 || ((V & 1) || (ldl % 2 == 0 && lddl % 2 == 0 && lddb % 2 == 0)),
               "token_nll_bwd: even row strides are needed for the bf16 gradient copy");
  GTOS_KLAUNCH(token_nll_bwd_kernel, dim3((unsigned)rows), dim3(256), 0, st, dloss_row, logits, ldl, V, align, S, copy_seq, target, B, pad_idx,
                                                       stats, dlogits, lddl, dgate_logits, dalign,
                                                       reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), lddb);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// bank -> dense relation gather (generator/generator.py:79) and its backward scatter-add.
// forward writes the fp32 tensor the caller's contract needs AND the bf16 copy the tensor-core kernels read.
// ---------------------------------------------------------------------------------------
__global__ void bank_gather_kernel(const float* __restrict__ bank, const long long* __restrict__ idx, long P, int D,
                                   float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
  GTOS_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long p = warp; p < P; p += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(bank + idx[p] * D);
    for (int c = lane; c < D / 4; c += 32) {
      float4 v = src[c];
      if (out_f32) reinterpret_cast<float4*>(out_f32 + p * D)[c] = v;
      if (out_bf16) reinterpret_cast<uint2*>(out_bf16 + p * D)[c] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
  }
}

__global__ void bank_scatter_add_kernel(const float* __restrict__ d_rel, const long long* __restrict__ idx, long P, int D,
                                        float* __restrict__ d_bank) {
  GTOS_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long p = warp; p < P; p += nwarps) {
    float* dst = d_bank + idx[p] * D;
    const float4* src = reinterpret_cast<const float4*>(d_rel + p * D);
    for (int c = lane; c < D / 4; c += 32) {
      float4 v = src[c];
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * c), "f"(v.x), "f"(v.y), "f"(v.z),
                   "f"(v.w)
                   : "memory");
    }
  }
}

// backward of the gather, sorted: d_bank[r] = sum of d_rel rows whose index is r.  The plain scatter above issues one vector
// reduction per pair and chunk; at config 2 the <TL> row alone receives 41 % of the pairs and the L2 serialises them (417 us
// for 220 MB).  Here the pairs arrive sorted by bank row (`order` = pair indices, `keys` = their rows, one radix sort per
// batch made during the forward pass): a warp walks 32 consecutive sorted pairs, keeps the running sum of the current row
// in registers and issues ONE reduction per (row, window) - 32x fewer reductions on the hot rows, every d_rel byte read once.
template <int U>
__global__ void __launch_bounds__(256) bank_segsum_f32_kernel(const float* __restrict__ d_rel, const long long* __restrict__ order,
                                                              const long long* __restrict__ keys, long P, int D,
                                                              float* __restrict__ d_bank) {
  GTOS_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long p0 = warp * 32;
  if (p0 >= P) return;
  const int cnt = (int)((P - p0) < 32 ? (P - p0) : 32);
  const long long my_row = lane < cnt ? order[p0 + lane] : 0;
  const long long my_key = lane < cnt ? keys[p0 + lane] : -1;
  const int nch = D / 4;
  float4 acc[U];
#pragma unroll
  for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  long long cur = __shfl_sync(0xffffffffu, my_key, 0);
  auto flush = [&](long long r) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ch = lane + 32 * u;
      if (ch < nch) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d_bank + r * D + 4 * ch), "f"(acc[u].x),
                     "f"(acc[u].y), "f"(acc[u].z), "f"(acc[u].w) : "memory");
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  constexpr int RIF = 4;                       // rows in flight per lane
  for (int q0 = 0; q0 < cnt; q0 += RIF) {
    float4 v[RIF][U];
#pragma unroll
    for (int t = 0; t < RIF; ++t) {
      const long long row = __shfl_sync(0xffffffffu, my_row, (q0 + t) & 31);
      const float4* src = reinterpret_cast<const float4*>(d_rel + row * D);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ch = lane + 32 * u;
        v[t][u] = (q0 + t < cnt && ch < nch) ? src[ch] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int t = 0; t < RIF; ++t) {
      if (q0 + t < cnt) {
        const long long key = __shfl_sync(0xffffffffu, my_key, (q0 + t) & 31);
        if (key != cur) {
          flush(cur);
          cur = key;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc[u].x += v[t][u].x; acc[u].y += v[t][u].y; acc[u].z += v[t][u].z; acc[u].w += v[t][u].w; }
      }
    }
  }
  flush(cur);
}

int bank_segsum_f32(const float* d_rel, const long long* order, const long long* keys, long P, int D, float* d_bank, long R,
                    cudaStream_t st) {
  GTOS_REQUIRE(D % 4 == 0 && D <= 1024, "bank_segsum: D must be a multiple of 4 and <= 1024 (got %d)", D);
  GTOS_CHECK_CUDA(cudaMemsetAsync(d_bank, 0, sizeof(float) * (size_t)R * D, st));
  if (P == 0) return GTOS_OK;
  const long warps = (P + 31) / 32;
  const unsigned blocks = (unsigned)((warps + 7) / 8);
  const int U = (D / 4 + 31) / 32;
  if (U <= 1) GTOS_KLAUNCH(bank_segsum_f32_kernel<1>, dim3(blocks), dim3(256), 0, st, d_rel, order, keys, P, D, d_bank);
  else if (U <= 2) GTOS_KLAUNCH(bank_segsum_f32_kernel<2>, dim3(blocks), dim3(256), 0, st, d_rel, order, keys, P, D, d_bank);
  else if (U <= 4) GTOS_KLAUNCH(bank_segsum_f32_kernel<4>, dim3(blocks), dim3(256), 0, st, d_rel, order, keys, P, D, d_bank);
  else GTOS_KLAUNCH(bank_segsum_f32_kernel<8>, dim3(blocks), dim3(256), 0, st, d_rel, order, keys, P, D, d_bank);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// evaluation batches (generator.py:83-88): relation[p] = mean over the pair's shortest paths of bank rows, where index 0
// (<PAD>) marks an empty slot: sum_k [idx[p][k] != 0] bank[idx[p][k]] / max(1, #{k: idx[p][k] != 0}).  One warp per pair.
__global__ void bank_gather_mean_kernel(const float* __restrict__ bank, const long long* __restrict__ idx, long P, int K, int D,
                                        float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
  GTOS_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long p = warp; p < P; p += nwarps) {
    const long long* ip = idx + p * K;
    int cnt = 0;
    for (int k = 0; k < K; ++k) cnt += (ip[k] != 0);
    const float inv = 1.f / (float)(cnt > 0 ? cnt : 1);
    for (int c = lane; c < D / 4; c += 32) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < K; ++k) {
        const long long r = ip[k];
        if (r != 0) {
          const float4 v = reinterpret_cast<const float4*>(bank + r * D)[c];
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
      }
      a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
      if (out_f32) reinterpret_cast<float4*>(out_f32 + p * D)[c] = a;
      if (out_bf16) reinterpret_cast<uint2*>(out_bf16 + p * D)[c] = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
    }
  }
}

int bank_gather_mean(const float* bank, const long long* idx, long P, int K, int D, float* out_f32, void* out_bf16,
                     cudaStream_t st) {
  GTOS_REQUIRE(D % 4 == 0 && K >= 1, "bank_gather_mean: D must be a multiple of 4 and K >= 1");
  if (P == 0) return GTOS_OK;
  long blocks = (P + 7) / 8;
  if (blocks > 148L * 16) blocks = 148L * 16;
  GTOS_KLAUNCH(bank_gather_mean_kernel, dim3((unsigned)blocks), dim3(256), 0, st, bank, idx, P, K, D, out_f32,
               reinterpret_cast<__nv_bfloat16*>(out_bf16));
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int bank_gather(const float* bank, const long long* idx, long P, int D, float* out_f32, void* out_bf16, cudaStream_t st) {
  GTOS_REQUIRE(D % 4 == 0, "bank_gather: D must be a multiple of 4");
  if (P == 0) return GTOS_OK;
  long blocks = (P * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  GTOS_KLAUNCH(bank_gather_kernel, dim3((unsigned)blocks), dim3(256), 0, st, bank, idx, P, D, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int bank_scatter_add(const float* d_rel, const long long* idx, long P, int D, float* d_bank, long R, cudaStream_t st) {
  GTOS_REQUIRE(D % 4 == 0, "bank_scatter_add: D must be a multiple of 4");
  GTOS_CHECK_CUDA(cudaMemsetAsync(d_bank, 0, sizeof(float) * R * D, st));
  if (P == 0) return GTOS_OK;
  long blocks = (P * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  GTOS_KLAUNCH(bank_scatter_add_kernel, dim3((unsigned)blocks), dim3(256), 0, st, d_rel, idx, P, D, d_bank);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

__global__ void rel_dqk_kernel(const __nv_bfloat16* __restrict__ G, RelTiling rt, float* __restrict__ dq,
                               float* __restrict__ dk, long ld, __nv_bfloat16* __restrict__ dq_b,
                               __nv_bfloat16* __restrict__ dk_b) {
  GTOS_PDL_PROLOGUE();
  // block = (node n, batch b); thread = one 8-column (16-byte) chunk of the 2D-wide G row.
  // chunks inside the d(q+ra) half of a head sum over keys j (-> dq[n]); chunks inside the d(k+rb) half sum
  // over queries i (-> dk[n]).  Every G element is read exactly once, with 16-byte loads.
  const int n = blockIdx.x, b = blockIdx.y;
  const int D = rt.D, hd = rt.hd;
  const int c = threadIdx.x;                 // chunk index, 2D/8 chunks per row
  const int col = c * 8;
  const int h = col / (2 * hd), w = col % (2 * hd);
  const bool is_x = w < hd;
  float acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.f;
  const int nb = is_x ? n / rt.bi : n / rt.bj;       // block index of the fixed coordinate
  const int nr = is_x ? n % rt.bi : n % rt.bj;
  // G row of (running coordinate o = ob * blk + oo):  base + ob * blk_stride + oo * in_stride.  Same loop for both
  // halves (lanes of one warp hold both kinds of chunk) and no division inside it: the first version spent most of
  // its issue slots on four integer divisions per 16-byte load.
  //   fixed query i = n, running key j:   row = ((b * nj_blk + jb) * ni_blk + nb) * 128 + jj * bi + nr
  //   fixed key j = n, running query i:   row = ((b * nj_blk + nb) * ni_blk + ib) * 128 + nr * bi + ii
  const int blk = is_x ? rt.bj : rt.bi;
  const long in_stride = is_x ? rt.bi : 1;
  const long blk_stride = is_x ? (long)rt.ni_blk * 128 : 128;
  const long base = is_x ? ((long)b * rt.nj_blk * rt.ni_blk + nb) * 128 + nr
                         : (((long)b * rt.nj_blk + nb) * rt.ni_blk) * 128 + (long)nr * rt.bi;
  const long wrap = blk_stride - (long)blk * in_stride;
  const __nv_bfloat16* gp = G + col;
  long row = base;
  int oo = 0;
#pragma unroll 4
  for (int o = 0; o < rt.N; ++o) {
    const uint4 v = *reinterpret_cast<const uint4*>(gp + row * (2L * D));
    row += in_stride;
    if (++oo == blk) {
      oo = 0;
      row += wrap;
    }
    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 f = __bfloat1622float2(p2[t]);
      acc[2 * t] += f.x;
      acc[2 * t + 1] += f.y;
    }
  }
  const long off = ((long)n * rt.B + b) * ld + h * hd + (is_x ? w : w - hd);
  float* dst = (is_x ? dq : dk) + off;
#pragma unroll
  for (int t = 0; t < 8; ++t) dst[t] = acc[t];
  __nv_bfloat16* dst_b = is_x ? dq_b : dk_b;      // optional operand copy for the in_proj backward GEMMs (same layout)
  if (dst_b)
    *reinterpret_cast<uint4*>(dst_b + off) = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                        pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
}

int rel_dqk(const void* G, const RelTiling& rt, float* dq, float* dk, long ld, void* dq_bf16, void* dk_bf16,
            cudaStream_t st) {
  dim3 grid(rt.N, rt.B);
  const int thr = 2 * rt.D / 8;
  GTOS_REQUIRE(thr <= 1024 && rt.hd % 8 == 0, "rel_dqk: unsupported D=%d hd=%d", rt.D, rt.hd);
  GTOS_REQUIRE((dq_bf16 == nullptr) == (dk_bf16 == nullptr), "rel_dqk: the bf16 copies go together");
  GTOS_REQUIRE(!dq_bf16 || (ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dq_bf16) & 15) == 0 &&
                            (reinterpret_cast<uintptr_t>(dk_bf16) & 15) == 0),
               "rel_dqk: bf16 copies need 16-byte aligned rows");
  GTOS_KLAUNCH(rel_dqk_kernel, dim3(grid), dim3(thr), 0, st, reinterpret_cast<const __nv_bfloat16*>(G), rt, dq, dk, ld,
               reinterpret_cast<__nv_bfloat16*>(dq_bf16), reinterpret_cast<__nv_bfloat16*>(dk_bf16));
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// bank-factorised backward of the relation terms (caller generator.py:76-79).
// relation = bank[idx] and relation_in_proj has no bias (graph_transformer.py:80), so
//   d relation_in_proj.weight = sum_pairs G_p^T bank[idx_p] = (sum over bank rows r of S_r^T bank_r),
//   d bank[r]                 = S_r * W,          S_r = sum_{p : idx_p = r} G_p
// i.e. ONE segmented sum of the per-pair gradient rows G (bf16, written by gtos_rel_grad) followed by two GEMMs
// over R bank rows instead of two GEMMs over P pair rows + a [N,N,B,D] fp32 d_relation + an atomic scatter.
// ---------------------------------------------------------------------------------------
__global__ void rel_pair_keys_kernel(const long long* __restrict__ idx, RelTiling rt, int R, int* __restrict__ keys) {
  GTOS_PDL_PROLOGUE();
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;     // G row = tile * 128 + r
  if (g >= (long)rt.tiles * 128) return;
  const int tile = (int)(g >> 7), r = (int)(g & 127);
  const int ib = tile % rt.ni_blk, t2 = tile / rt.ni_blk;
  const int jb = t2 % rt.nj_blk, b = t2 / rt.nj_blk;
  const int jj = r / rt.bi, ii = r - jj * rt.bi;
  const int j = jb * rt.bj + jj, i = ib * rt.bi + ii;
  const bool valid = jj < rt.bj && i < rt.N && j < rt.N && b < rt.B;
  int key = R;                                                    // padding rows of a tile sort to the end
  if (valid) {
    const long long v = idx[((long)j * rt.N + i) * rt.B + b];
    key = (v >= 0 && v < R) ? (int)v : R;
  }
  keys[g] = key;
}

int rel_pair_keys(const long long* idx, const RelTiling& rt, int R, int* keys, cudaStream_t st) {
  const long rows = (long)rt.tiles * 128;
  if (rows == 0) return GTOS_OK;
  GTOS_KLAUNCH(rel_pair_keys_kernel, dim3((unsigned)((rows + 255) / 256)), dim3(256), 0, st, idx, rt, R, keys);
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

static constexpr int SEG_CH = 32;    // sorted pairs per warp
static constexpr int SEG_COLS = 512; // columns per warp (grid.y splits wider rows): keeps SEG_RIF rows in flight per lane at 2 blocks per SM
static constexpr int SEG_RIF = 6;

// one warp per 32 consecutive positions of the row-sorted pair list and per 512-column slice of the C-wide row; lane l
// owns the 8-column (16-byte) chunks l, l+32 of the slice.  A bank row whose pairs all lie inside the warp's window
// is written once (bf16); a row that crosses a window boundary is reduced into the fp32 `spill` row with vector
// atomics and converted by rel_segsum_span_kernel afterwards.
template <int NCH>
__global__ void __launch_bounds__(256, 2) rel_segsum_kernel(const __nv_bfloat16* __restrict__ G, const int* __restrict__ order,
                                                            const int* __restrict__ keys, long n, int C,
                                                            __nv_bfloat16* __restrict__ out, long ldo,
                                                            float* __restrict__ spill) {
  GTOS_PDL_PROLOGUE();
  __shared__ int s_row[8], s_ok[8];
  __shared__ float s_part[8 * NCH * 256];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long warp = (long)blockIdx.x * 8 + wib;
  const int col0 = blockIdx.y * SEG_COLS;                      // first column of this slice
  const int ncols = (C - col0) < SEG_COLS ? (C - col0) : SEG_COLS;
  const long p0 = warp * SEG_CH;
  const bool active = p0 < n;
  const int cnt = active ? (int)((n - p0) < SEG_CH ? (n - p0) : SEG_CH) : 0;
  const int my_key = lane < cnt ? keys[p0 + lane] : -1;
  const int my_row = lane < cnt ? order[p0 + lane] : 0;
  const int nchunks = ncols >> 3;
  float acc[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[c][t] = 0.f;
  int cur = __shfl_sync(0xffffffffu, my_key, 0);
  bool open_left = active && p0 > 0 && keys[p0 - 1] == cur;
  bool single = true;                           // the whole window belongs to one bank row

  auto flush = [&](int r, bool spanning) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        if (spanning) {
          float* d = spill + (long)r * C + col0 + ch * 8;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(acc[c][0]), "f"(acc[c][1]),
                       "f"(acc[c][2]), "f"(acc[c][3]) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(acc[c][4]), "f"(acc[c][5]),
                       "f"(acc[c][6]), "f"(acc[c][7]) : "memory");
        } else {
          *reinterpret_cast<uint4*>(out + (long)r * ldo + col0 + ch * 8) =
              make_uint4(pack_bf16x2(acc[c][0], acc[c][1]), pack_bf16x2(acc[c][2], acc[c][3]),
                         pack_bf16x2(acc[c][4], acc[c][5]), pack_bf16x2(acc[c][6], acc[c][7]));
        }
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[c][t] = 0.f;
    }
  };

  for (int q0 = 0; q0 < cnt; q0 += SEG_RIF) {
    // SEG_RIF rows in flight per lane
    uint4 v[SEG_RIF][NCH];
#pragma unroll
    for (int u = 0; u < SEG_RIF; ++u) {
      const int row = __shfl_sync(0xffffffffu, my_row, (q0 + u) & 31);
      const uint4* src = reinterpret_cast<const uint4*>(G + (long)row * C + col0);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int ch = lane + 32 * c;
        v[u][c] = (q0 + u < cnt && ch < nchunks) ? src[ch] : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < SEG_RIF; ++u) {
      if (q0 + u < cnt) {
        const int key = __shfl_sync(0xffffffffu, my_key, (q0 + u) & 31);
        if (key != cur) {
          flush(cur, open_left);
          open_left = false;
          single = false;
          cur = key;
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u][c]);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = __bfloat1622float2(h[t]);
            acc[c][2 * t] += f.x;
            acc[c][2 * t + 1] += f.y;
          }
        }
      }
    }
  }
  const bool open_right = active && (p0 + cnt < n) && keys[p0 + cnt] == cur;
  const bool spanning = open_left || open_right;
  // a bank row with thousands of pairs (e.g. the <TL> path, data.py:151-154) fills whole blocks: combine the eight
  // windows in shared memory and issue ONE reduction per block instead of eight contended ones
  if (lane == 0) {
    s_row[wib] = cur;
    s_ok[wib] = active && single && spanning;
  }
  __syncthreads();
  bool all = true;
#pragma unroll
  for (int w = 0; w < 8; ++w) all = all && s_ok[w] && (s_row[w] == s_row[0]);
  if (all) {
    float* mine = s_part + wib * (NCH * 256);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        *reinterpret_cast<float4*>(mine + ch * 8) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
        *reinterpret_cast<float4*>(mine + ch * 8 + 4) = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
      }
    }
    __syncthreads();
    for (int c4 = threadIdx.x; c4 < ncols / 4; c4 += 256) {
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float4 t4 = *reinterpret_cast<const float4*>(s_part + w * (NCH * 256) + c4 * 4);
        sum.x += t4.x; sum.y += t4.y; sum.z += t4.z; sum.w += t4.w;
      }
      float* d = spill + (long)s_row[0] * C + col0 + c4 * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(sum.x), "f"(sum.y), "f"(sum.z),
                   "f"(sum.w) : "memory");
    }
  } else if (active) {
    flush(cur, spanning);
  }
}

// mode 0: zero the spill rows of bank rows that cross a window boundary; mode 1: convert them to bf16 (one warp per
// boundary; the first boundary inside a row does the conversion)
__global__ void rel_segsum_span_kernel(const int* __restrict__ keys, long n, int C, float* __restrict__ spill,
                                       __nv_bfloat16* __restrict__ out, long ldo, int mode) {
  GTOS_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long b = (warp + 1) * SEG_CH;
  if (b >= n) return;
  const int r = keys[b];
  if (keys[b - 1] != r) return;
  if (mode == 0) {
    float4* d = reinterpret_cast<float4*>(spill + (long)r * C);
    for (int c = lane; c < C / 4; c += 32) d[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    if (b - SEG_CH >= 1 && keys[b - SEG_CH - 1] == r) return;   // an earlier boundary of the same row converts it
    const float4* s4 = reinterpret_cast<const float4*>(spill + (long)r * C);
    uint2* d = reinterpret_cast<uint2*>(out + (long)r * ldo);
    for (int c = lane; c < C / 4; c += 32) {
      const float4 v = s4[c];
      d[c] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
  }
}

int rel_segsum(const void* G, const int* order, const int* keys, long n, int C, void* out_bf16, long ldo, float* spill,
               cudaStream_t st) {
  GTOS_REQUIRE(C % 8 == 0 && C <= 2048 && ldo % 8 == 0 && ldo >= C,
               "rel_segsum: row width %d must be a multiple of 8 and <= 2048 (ldo %ld)", C, ldo);
  if (n == 0) return GTOS_OK;
  const long warps = (n + SEG_CH - 1) / SEG_CH;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(G);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (warps > 1) {
    GTOS_KLAUNCH(rel_segsum_span_kernel, dim3(blocks), dim3(256), 0, st, keys, n, C, spill, o, ldo, 0);
    GTOS_LAUNCH_CHECK();
  }
  const dim3 grid((unsigned)((warps + 7) / 8), (unsigned)((C + SEG_COLS - 1) / SEG_COLS));
  if (C <= 256) GTOS_KLAUNCH(rel_segsum_kernel<1>, dim3(grid), dim3(256), 0, st, g, order, keys, n, C, o, ldo, spill);
  else GTOS_KLAUNCH(rel_segsum_kernel<2>, dim3(grid), dim3(256), 0, st, g, order, keys, n, C, o, ldo, spill);
  GTOS_LAUNCH_CHECK();
  if (warps > 1) {
    GTOS_KLAUNCH(rel_segsum_span_kernel, dim3(blocks), dim3(256), 0, st, keys, n, C, spill, o, ldo, 1);
    GTOS_LAUNCH_CHECK();
  }
  return GTOS_OK;
}


}  // namespace gtos
