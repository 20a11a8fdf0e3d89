// gtos_b200 -- extern "C" surface of libgtos_b200.so (declared in include/gtos_b200.h).
#include <stdarg.h>
#include <string.h>

#include "../../include/gtos_b200.h"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "graph_paths_core.h"

namespace gtos {
static thread_local char g_err[512] = "";
unsigned long long g_kernel_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace gtos

using namespace gtos;
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

const char* gtos_last_error(void) { return g_err; }
int gtos_abi_version(void) { return 5; }
int gtos_set_sm_reserve(int32_t n) { return set_sm_reserve(n); }
uint64_t gtos_launch_count(void) { return __atomic_load_n(&g_kernel_launches, __ATOMIC_RELAXED); }

int gtos_device_check(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device");
    return GTOS_ERR_NO_DEVICE;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    set_error("gtos_b200 kernels are built for sm_100a only (device is sm_%d0)", major);
    return GTOS_ERR_NO_DEVICE;
  }
  CUtensorMap tm;
  static __nv_bfloat16* probe = nullptr;
  if (!probe) GTOS_CHECK_CUDA(cudaMalloc(&probe, 128 * 64 * 2));
  return make_tmap_2d_bf16(&tm, probe, 128, 64, 64, 128);
}

int gtos_cast_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int32_t cols, void* stream) {
  return cast_f32_bf16(src, lds, dst, ldd, rows, cols, S(stream));
}

int gtos_cast_colsum(const float* src, int64_t lds, void* dst, int64_t ldd, float* sums, int64_t rows, int32_t cols,
                     void* stream) {
  return cast_colsum(src, lds, dst, ldd, sums, rows, cols, S(stream));
}

int gtos_weight_prep(const float* W, int32_t R, int32_t C, void* Wb, int64_t ldw, void* Wt, int64_t ldt,
                     int32_t rel_heads, void* stream) {
  int perm_D = 0, perm_hd = 0;
  if (rel_heads > 0) {
    GTOS_REQUIRE(R == 2 * C && C % rel_heads == 0, "weight_prep: relation weight must be [2D, D] with D %% H == 0");
    perm_D = C;
    perm_hd = C / rel_heads;
  }
  return weight_prep(W, R, C, Wb, ldw, Wt, ldt, perm_D, perm_hd, S(stream));
}

int gtos_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, float* out_f32, int64_t ldo,
                 void* out_bf16, int64_t ldob, int32_t M, int32_t N, int32_t K, int32_t relu, int32_t accumulate,
                 void* stream) {
  if (M == 0 || N == 0) return GTOS_OK;
  GemmTnArgs a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.lda = lda; a.Bm = B; a.ldb = ldb; a.M = M; a.N = N; a.K = K;
  a.bias = bias; a.out_f32 = out_f32; a.ldo = ldo; a.out_bf16 = out_bf16; a.ldob = ldob;
  a.relu = relu; a.accumulate = accumulate;
  return launch_gemm_tn(MODE_PLAIN, a, S(stream));
}

int gtos_gemm_tn_add(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, const float* addend,
                     int64_t ldadd, float* out_f32, int64_t ldo, int32_t M, int32_t N, int32_t K, void* stream) {
  if (M == 0 || N == 0) return GTOS_OK;
  GemmTnArgs a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.lda = lda; a.Bm = B; a.ldb = ldb; a.M = M; a.N = N; a.K = K;
  a.bias = bias; a.addend = addend; a.ldadd = ldadd; a.out_f32 = out_f32; a.ldo = ldo;
  return launch_gemm_tn(MODE_PLAIN, a, S(stream));
}

int64_t gtos_gemm_nn_workspace(int32_t M, int32_t N, int32_t Kd) { return gemm_nn_workspace_elems(M, N, Kd, 0); }

int gtos_gemm_nn(const void* A, int64_t lda, const void* B, int64_t ldb, float* out, int64_t ldo, int32_t M, int32_t N,
                 int32_t Kd, float* workspace, int64_t workspace_elems, void* stream) {
  if (M == 0 || N == 0) return GTOS_OK;
  GemmNnArgs a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.lda = lda; a.Bm = B; a.ldb = ldb; a.M = M; a.N = N; a.Kd = Kd; a.rel = 0;
  a.out = out; a.ldo = ldo; a.workspace = workspace; a.workspace_elems = workspace_elems;
  return launch_gemm_nn(a, S(stream));
}

int gtos_rel_tiling(int32_t N, int32_t B, int32_t D, int32_t H, int32_t* out5) {
  RelTiling rt;
  int e = choose_rel_tiling(&rt, N, B, D, H);
  if (e) return e;
  out5[0] = rt.bi; out5[1] = rt.bj; out5[2] = rt.ni_blk; out5[3] = rt.nj_blk; out5[4] = rt.tiles;
  return GTOS_OK;
}

static int rel_args(GemmTnArgs* a, int32_t N, int32_t B, int32_t D, int32_t H) {
  memset(a, 0, sizeof(*a));
  int e = choose_rel_tiling(&a->rt, N, B, D, H);
  if (e) return e;
  a->N = 2 * D;
  a->K = D;
  a->M = a->rt.tiles * 128;
  return GTOS_OK;
}

int gtos_rel_score(const void* relb, const void* Wperm, const void* q, const void* k, int64_t ldqk, float* scores,
                   int32_t N, int32_t B, int32_t D, int32_t H, void* stream) {
  GemmTnArgs a;
  int e = rel_args(&a, N, B, D, H);
  if (e) return e;
  GTOS_REQUIRE(ldqk % 8 == 0, "rel_score: q/k row stride must be a multiple of 8 bf16 elements");
  a.A = relb; a.lda = D; a.Bm = Wperm; a.ldb = D; a.q = q; a.k = k; a.ldqk = ldqk; a.scores = scores;
  return launch_gemm_tn(MODE_SCORE, a, S(stream));
}

int gtos_rel_attn_fusable(int32_t N, int32_t B, int32_t D, int32_t H) {
  RelTiling rt;
  if (choose_rel_tiling(&rt, N, B, D, H)) return 0;
  return rel_attn_fusable(rt) ? 1 : 0;
}

int gtos_rel_attn_fwd(const void* relb, const void* Wperm, const void* q, const void* k, int64_t ldqk, const void* v,
                      int64_t ldv, const uint8_t* key_pad, float p_drop, const void* seed_ptr, uint64_t seed_off,
                      float* probs, float* probs_dropped, float* att, int64_t ldatt, void* att_bf16, int32_t N, int32_t B,
                      int32_t D, int32_t H, void* stream) {
  GemmTnArgs a;
  int e = rel_args(&a, N, B, D, H);
  if (e) return e;
  GTOS_REQUIRE(ldqk % 8 == 0, "rel_attn_fwd: q/k row stride must be a multiple of 8 bf16 elements");
  if (!rel_attn_fusable(a.rt)) {
    set_error("rel_attn_fwd: N=%d D=%d H=%d is not fusable (needs all keys of a query in one 128-pair tile and head_dim 64); "
              "use gtos_rel_score + gtos_attn_fwd", N, D, H);
    return GTOS_ERR_UNSUPPORTED;
  }
  a.A = relb; a.lda = D; a.Bm = Wperm; a.ldb = D; a.q = q; a.k = k; a.ldqk = ldqk;
  a.fuse = 1; a.v = v; a.ldv = ldv; a.key_pad = key_pad; a.p_drop = p_drop; a.seed_ptr = seed_ptr; a.seed_off = seed_off;
  a.probs = probs; a.probs_dropped = probs_dropped; a.att = att; a.ldatt = ldatt; a.att_bf16 = att_bf16;
  return launch_gemm_tn(MODE_SCORE, a, S(stream));
}

int gtos_rel_grad(const void* relb, const void* Wperm, const void* q, const void* k, int64_t ldqk,
                  const float* dscores, void* G, int32_t N, int32_t B, int32_t D, int32_t H, void* stream) {
  GemmTnArgs a;
  int e = rel_args(&a, N, B, D, H);
  if (e) return e;
  GTOS_REQUIRE(ldqk % 8 == 0, "rel_grad: q/k row stride must be a multiple of 8 bf16 elements");
  a.A = relb; a.lda = D; a.Bm = Wperm; a.ldb = D; a.q = q; a.k = k; a.ldqk = ldqk; a.dscores = dscores; a.G = G;
  return launch_gemm_tn(MODE_GRAD, a, S(stream));
}

int gtos_rel_drel(const void* G, const void* WpermT, float* d_relation, int32_t accumulate, int32_t N, int32_t B,
                  int32_t D, int32_t H, void* stream) {
  GemmTnArgs a;
  int e = rel_args(&a, N, B, D, H);
  if (e) return e;
  a.A = G; a.lda = 2 * D; a.Bm = WpermT; a.ldb = 2 * D;
  a.N = D; a.K = 2 * D;
  a.out_f32 = d_relation; a.ldo = D; a.accumulate = accumulate;
  return launch_gemm_tn(MODE_DREL, a, S(stream));
}

int64_t gtos_rel_dw_workspace(int32_t N, int32_t B, int32_t D, int32_t H) {
  RelTiling rt;
  if (choose_rel_tiling(&rt, N, B, D, H)) return -1;
  return gemm_nn_workspace_elems(2 * D, D, rt.tiles * 128, 1);
}

int gtos_rel_dw(const void* G, const void* relb, float* dW, float* workspace, int64_t workspace_elems, int32_t N,
                int32_t B, int32_t D, int32_t H, void* stream) {
  GemmNnArgs a;
  memset(&a, 0, sizeof(a));
  int e = choose_rel_tiling(&a.rt, N, B, D, H);
  if (e) return e;
  a.A = G; a.lda = 2 * D; a.Bm = relb; a.ldb = D; a.M = 2 * D; a.N = D; a.Kd = a.rt.tiles * 128; a.rel = 1;
  a.out = dW; a.ldo = D; a.workspace = workspace; a.workspace_elems = workspace_elems;
  return launch_gemm_nn(a, S(stream));
}

int gtos_rel_pair_keys(const int64_t* idx, int32_t N, int32_t B, int32_t D, int32_t H, int32_t R, int32_t* keys,
                       void* stream) {
  RelTiling rt;
  int e = choose_rel_tiling(&rt, N, B, D, H);
  if (e) return e;
  return rel_pair_keys(reinterpret_cast<const long long*>(idx), rt, R, keys, S(stream));
}

int gtos_rel_segsum(const void* G, const int32_t* order, const int32_t* keys, int64_t n, int32_t C, void* out_bf16,
                    int64_t ldo, float* spill, void* stream) {
  return rel_segsum(G, order, keys, n, C, out_bf16, ldo, spill, S(stream));
}

int gtos_rel_dw_bank(const void* Sb, int64_t lds, const void* bankb, float* dW, int32_t R, int32_t D, int32_t H,
                     void* stream) {
  GTOS_REQUIRE(H > 0 && D % H == 0, "rel_dw_bank: D=%d not divisible by H=%d", D, H);
  if (R == 0) return GTOS_OK;
  GemmNnArgs a;
  memset(&a, 0, sizeof(a));
  a.A = Sb; a.lda = lds; a.Bm = bankb; a.ldb = D; a.M = 2 * D; a.N = D; a.Kd = R; a.rel = 0;
  a.perm_D = D; a.perm_hd = D / H;
  a.out = dW; a.ldo = D;
  return launch_gemm_nn(a, S(stream));
}

int gtos_rel_dqk(const void* G, float* dq, float* dk, int64_t ld, void* dq_bf16, void* dk_bf16, int32_t N, int32_t B,
                 int32_t D, int32_t H, void* stream) {
  RelTiling rt;
  int e = choose_rel_tiling(&rt, N, B, D, H);
  if (e) return e;
  return rel_dqk(G, rt, dq, dk, ld, dq_bf16, dk_bf16, S(stream));
}

int gtos_rel_attn_banked_fwd(const void* PB, int64_t ldpb, const int64_t* idx, const void* q, const void* k, int64_t ldqk,
                             const float* v, int64_t ldv, const uint8_t* key_pad, const uint8_t* attn_mask, float p_drop,
                             const void* seed_ptr, uint64_t seed_off, float* probs, float* probs_dropped, float* out,
                             int64_t ldo, void* out_bf16, int32_t N, int32_t B, int32_t D, int32_t H, int32_t R,
                             void* stream) {
  RelBankedArgs a;
  memset(&a, 0, sizeof(a));
  a.PB = PB; a.ldpb = ldpb; a.idx = reinterpret_cast<const long long*>(idx); a.q = q; a.k = k; a.ldqk = ldqk;
  a.v = v; a.ldv = ldv; a.key_pad = key_pad; a.attn_mask = attn_mask; a.p_drop = p_drop; a.seed_ptr = seed_ptr;
  a.seed_off = seed_off; a.probs = probs; a.probs_dropped = probs_dropped; a.out = out; a.ldo = ldo; a.out_bf16 = out_bf16;
  a.N = N; a.B = B; a.D = D; a.H = H; a.R = R;
  return rel_attn_banked_fwd(a, S(stream));
}

int gtos_rel_grad_banked(const void* PB, int64_t ldpb, const int64_t* idx, const void* q, const void* k, int64_t ldqk,
                         const float* dscores, void* G, int32_t N, int32_t B, int32_t D, int32_t H, int32_t R,
                         void* stream) {
  RelBankedArgs a;
  memset(&a, 0, sizeof(a));
  a.PB = PB; a.ldpb = ldpb; a.idx = reinterpret_cast<const long long*>(idx); a.q = q; a.k = k; a.ldqk = ldqk;
  a.dscores = dscores; a.G = G; a.N = N; a.B = B; a.D = D; a.H = H; a.R = R;
  return rel_grad_banked(a, S(stream));
}

static void fill_attn(const gtos_attn_desc* d, AttnArgs* a) {
  a->T = d->T; a->S = d->S; a->B = d->B; a->H = d->H; a->hd = d->hd;
  a->q = d->q; a->ldq = d->ldq; a->k = d->k; a->ldk = d->ldk; a->v = d->v; a->ldv = d->ldv;
  a->scale = d->scale; a->scores_jt = d->scores_jt; a->key_pad = d->key_pad; a->attn_mask = d->attn_mask;
  a->p_drop = d->p_drop; a->seed_ptr = d->seed_ptr; a->seed_off = d->seed_off;
  a->probs = d->probs; a->probs_dropped = d->probs_dropped; a->out = d->out; a->ldo = d->ldo; a->out_bf16 = d->out_bf16;
  a->precise = d->precise;
}

int gtos_attn_fwd(const gtos_attn_desc* d, void* stream) {
  GTOS_REQUIRE(d && d->v && d->probs && d->out, "attn_fwd: null argument");
  GTOS_REQUIRE(d->scores_jt || (d->q && d->k), "attn_fwd: need either scores or q/k");
  AttnArgs a;
  fill_attn(d, &a);
  return attn_fwd(a, S(stream));
}

int gtos_attn_bwd_dk_on_query_side(const gtos_attn_desc* d) {
  if (!d) return 0;
  AttnArgs a;
  fill_attn(d, &a);
  return attn_bwd_dk_on_query_side(a, d->q != nullptr && d->k != nullptr);
}

int gtos_attn_bwd(const gtos_attn_desc* d, void* stream) {
  GTOS_REQUIRE(d && d->v && d->probs && d->dout && d->dscores_ts && d->dv, "attn_bwd: null argument");
  AttnBwdArgs g;
  memset(&g, 0, sizeof(g));
  fill_attn(d, &g.f);
  g.dout = d->dout; g.lddo = d->lddo; g.dprobs_extra = d->dprobs_extra;
  g.dscores_jt = d->dscores_jt; g.dscores_ts = d->dscores_ts;
  g.dq = d->dq; g.lddq = d->lddq; g.dk = d->dk; g.lddk = d->lddk; g.dv = d->dv; g.lddv = d->lddv;
  g.dq_bf16 = d->dq_bf16; g.dk_bf16 = d->dk_bf16; g.dv_bf16 = d->dv_bf16;
  GTOS_REQUIRE(!g.dq || (d->q && d->k && g.dk), "attn_bwd: decoder mode needs q, k, dq and dk");
  GTOS_REQUIRE(d->bwd_part >= 0 && d->bwd_part <= 2, "attn_bwd: bwd_part must be 0, 1 or 2");
  return attn_bwd(g, d->bwd_part, S(stream));
}

int gtos_zero_regions(void* const* ptrs, const int64_t* bytes, int32_t n, void* stream) {
  GTOS_REQUIRE(n >= 0 && (n == 0 || (ptrs && bytes)), "zero_regions: null argument");
  for (int32_t i0 = 0; i0 < n; i0 += ZERO_MAX_REGIONS) {              // HOST arrays: chunks of one launch each
    ZeroRegions z;
    memset(&z, 0, sizeof(z));
    const int m = n - i0 < ZERO_MAX_REGIONS ? n - i0 : ZERO_MAX_REGIONS;
    for (int i = 0; i < m; ++i) {
      GTOS_REQUIRE(bytes[i0 + i] >= 0 && (bytes[i0 + i] == 0 || ptrs[i0 + i]), "zero_regions: bad region %d", i0 + i);
      z.ptr[i] = ptrs[i0 + i];
      z.bytes[i] = bytes[i0 + i];
    }
    int e = zero_regions(z, m, S(stream));
    if (e) return e;
  }
  return GTOS_OK;
}

int gtos_split3(const float* src, int64_t ld_r, int64_t ld_c, int64_t rows, int32_t cols, void* dst, int64_t ldd,
                int32_t kp, int32_t role, void* stream) {
  return split3(src, ld_r, ld_c, rows, cols, dst, ldd, kp, role, S(stream));
}
int gtos_rel_score_f32(const float* PR, int64_t ldpr, const float* q, const float* k, int64_t ldqk, float* scores,
                       int32_t N, int32_t B, int32_t D, int32_t H, void* stream) {
  return rel_score_f32(PR, ldpr, q, k, ldqk, scores, N, B, D, H, S(stream));
}
int gtos_rel_grad_f32(const float* PR, int64_t ldpr, const float* q, const float* k, int64_t ldqk, const float* dscores,
                      float* G, int64_t ldg, int32_t N, int32_t B, int32_t D, int32_t H, void* stream) {
  return rel_grad_f32(PR, ldpr, q, k, ldqk, dscores, G, ldg, N, B, D, H, S(stream));
}
int gtos_rel_dqk_f32(const float* G, int64_t ldg, float* dq, float* dk, int64_t ld, int32_t N, int32_t B, int32_t D,
                     void* stream) {
  return rel_dqk_f32(G, ldg, dq, dk, ld, N, B, D, S(stream));
}
int gtos_relu_drop_bwd_f32(const float* dh_in, const float* act, float* dh_out, int64_t n, float p, void* stream) {
  return relu_drop_bwd_f32(dh_in, act, dh_out, n, p, S(stream));
}
int gtos_gru_gate_fwd_f32(const float* gi, int64_t ldgi, const float* gh, int64_t ldgh, const float* h_prev,
                          const int64_t* lengths, int32_t t, float* h_new, float* out_t, int64_t ldout, float* gates,
                          int64_t R, int32_t H, void* stream) {
  return gru_gate_fwd_f32(gi, ldgi, gh, ldgh, h_prev, reinterpret_cast<const long long*>(lengths), t, h_new, out_t, ldout,
                          gates, R, H, S(stream));
}
int gtos_gru_gate_bwd_f32(const float* dh, const float* dout_t, int64_t lddout, const float* gates, const float* h_prev,
                          const int64_t* lengths, int32_t t, float* dh_part, float* dgi, int64_t lddgi, float* dgh,
                          int64_t lddgh, int64_t R, int32_t H, void* stream) {
  return gru_gate_bwd_f32(dh, dout_t, lddout, gates, h_prev, reinterpret_cast<const long long*>(lengths), t, dh_part, dgi,
                          lddgi, dgh, lddgh, R, H, S(stream));
}

int gtos_add_ln_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y, void* y_bf16,
                    float* z, float* mean, float* rstd, int64_t rows, int32_t D, float p_drop, const void* seed_ptr,
                    uint64_t seed_off, void* stream) {
  return add_ln_fwd(x, res, gamma, beta, y, y_bf16, z, mean, rstd, rows, D, p_drop, seed_ptr, seed_off, S(stream));
}
int gtos_add_ln_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                    float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta, int64_t rows, int32_t D,
                    float p_drop, const void* seed_ptr, uint64_t seed_off, void* stream) {
  return add_ln_bwd(dy, z, mean, rstd, gamma, dres, dx, dx_bf16, dgamma, dbeta, rows, D, p_drop, seed_ptr, seed_off,
                    S(stream));
}
int gtos_ln_param_grad(const float* dy, const float* z, const float* mean, const float* rstd, float* dgamma, float* dbeta,
                       int64_t rows, int32_t D, void* stream) {
  return ln_param_grad(dy, z, mean, rstd, dgamma, dbeta, rows, D, S(stream));
}
int gtos_colsum(const float* x, int64_t ld, float* out, int64_t rows, int32_t cols, void* stream) {
  return colsum(x, ld, out, rows, cols, S(stream));
}
int gtos_colsum_bf16(const void* x, int64_t ld, float* out, int64_t rows, int32_t cols, void* stream) {
  return colsum_bf16(x, ld, out, rows, cols, S(stream));
}
int gtos_dropout_bf16(void* h, int64_t n, float p, const void* seed_ptr, uint64_t seed_off, void* stream) {
  return dropout_bf16(h, n, p, seed_ptr, seed_off, S(stream));
}
int gtos_dropout_f32(const float* x, float* out, int64_t n, float p, const void* seed_ptr, uint64_t seed_off,
                     void* stream) {
  return dropout_f32(x, out, n, p, seed_ptr, seed_off, S(stream));
}
int gtos_relu_drop_bwd(const float* dh_in, const void* act_bf16, float* dh_f32, void* dh_bf16, int64_t n, float p,
                       void* stream) {
  return relu_drop_bwd(dh_in, act_bf16, dh_f32, dh_bf16, n, p, S(stream));
}
int gtos_token_nll_fwd(const float* logits, int64_t ldl, int32_t V, const float* gate_logits, const float* align,
                       int32_t S, const int64_t* copy_seq, const int64_t* target, int64_t rows, int32_t B, int64_t pad_idx,
                       float* loss_row, float* stats, void* stream) {
  return token_nll_fwd(logits, ldl, V, gate_logits, align, S, reinterpret_cast<const long long*>(copy_seq),
                       reinterpret_cast<const long long*>(target), rows, B, pad_idx, loss_row, stats, S_(stream));
}
int gtos_token_nll_bwd(const float* dloss_row, const float* logits, int64_t ldl, int32_t V, const float* align, int32_t S,
                       const int64_t* copy_seq, const int64_t* target, int64_t rows, int32_t B, int64_t pad_idx,
                       const float* stats, float* dlogits, int64_t lddl, float* dgate_logits, float* dalign,
                       void* dlogits_bf16, int64_t lddb, void* stream) {
  return token_nll_bwd(dloss_row, logits, ldl, V, align, S, reinterpret_cast<const long long*>(copy_seq),
                       reinterpret_cast<const long long*>(target), rows, B, pad_idx, stats, dlogits, lddl, dgate_logits,
                       dalign, dlogits_bf16, lddb, S_(stream));
}
int gtos_bank_gather(const float* bank, const int64_t* idx, int64_t P, int32_t D, float* out_f32, void* out_bf16,
                     void* stream) {
  return bank_gather(bank, reinterpret_cast<const long long*>(idx), P, D, out_f32, out_bf16, S(stream));
}
int gtos_bank_segsum(const float* d_rel, const int64_t* order, const int64_t* keys, int64_t P, int32_t D, float* d_bank,
                     int64_t R, void* stream) {
  return bank_segsum_f32(d_rel, reinterpret_cast<const long long*>(order), reinterpret_cast<const long long*>(keys), P, D,
                         d_bank, R, S(stream));
}
int gtos_bank_gather_mean(const float* bank, const int64_t* idx, int64_t P, int32_t K, int32_t D, float* out_f32,
                          void* out_bf16, void* stream) {
  return bank_gather_mean(bank, reinterpret_cast<const long long*>(idx), P, K, D, out_f32, out_bf16, S(stream));
}
int gtos_bank_scatter_add(const float* d_rel, const int64_t* idx, int64_t P, int32_t D, float* d_bank, int64_t R,
                          void* stream) {
  return bank_scatter_add(d_rel, reinterpret_cast<const long long*>(idx), P, D, d_bank, R, S(stream));
}
int gtos_embed_gather(const float* table, const int64_t* idx, int64_t n, int32_t dim, float* out_f32, void* out_bf16,
                      int64_t ldb, float p_drop, const void* seed_ptr, uint64_t seed_off, void* stream) {
  return embed_gather(table, reinterpret_cast<const long long*>(idx), n, dim, out_f32, out_bf16, ldb, p_drop, seed_ptr,
                      seed_off, S(stream));
}
int gtos_embed_scatter_add(const float* dx, const int64_t* idx, int64_t n, int32_t dim, float* dtable, float p_drop,
                           const void* seed_ptr, uint64_t seed_off, void* stream) {
  return embed_scatter_add(dx, reinterpret_cast<const long long*>(idx), n, dim, dtable, p_drop, seed_ptr, seed_off,
                           S(stream));
}
int gtos_gru_weight_prep(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int32_t Kin,
                         int32_t H, int32_t Kx, void* Wcat, int64_t ldw, float* bcat, void* stream) {
  return gru_weight_prep(w_ih, w_hh, b_ih, b_hh, Kin, H, Kx, Wcat, ldw, bcat, S(stream));
}
int gtos_gru_step_fwd(const void* x, int64_t ldx, int32_t Kin, const void* hb, int64_t ldhb, const float* h_prev,
                      const void* Wcat, int64_t ldw, int32_t Kx, const float* bcat, const int64_t* lengths, int32_t t,
                      float* h_new, void* hb_new, int64_t ldhbn, void* out_t, int64_t ldout, void* gates, int64_t ldg,
                      int64_t R, int32_t H, void* stream) {
  GruStepArgs a;
  a.x = x; a.ldx = ldx; a.Kin = Kin; a.hb = hb; a.ldhb = ldhb; a.h_prev = h_prev; a.Wcat = Wcat; a.ldw = ldw; a.Kx = Kx;
  a.bcat = bcat; a.lengths = reinterpret_cast<const long long*>(lengths); a.t = t; a.h_new = h_new; a.hb_new = hb_new;
  a.ldhbn = ldhbn; a.out_t = out_t; a.ldout = ldout; a.gates = gates; a.ldg = ldg; a.R = (int)R; a.H = H;
  return launch_gru_step(a, S(stream));
}
int gtos_gru_gate_bwd(const float* dh, const float* dout_t, int64_t lddout, const void* gates, const float* h_prev,
                      const int64_t* lengths, int32_t t, float* dh_prev, void* dgi_bf16, int64_t lddgi, void* dgh_bf16,
                      int64_t lddgh, float* db_ih, float* db_hh, int64_t R, int32_t Hh, void* stream) {
  return gru_gate_bwd(dh, dout_t, lddout, gates, h_prev, reinterpret_cast<const long long*>(lengths), t, dh_prev, dgi_bf16,
                      lddgi, dgh_bf16, lddgh, db_ih, db_hh, R, Hh, S(stream));
}

int gtos_debug_attn_trace(uint64_t* host_out, int32_t enable) {
  return attn_debug_read_trace(reinterpret_cast<unsigned long long*>(host_out), enable);
}
int gtos_debug_read_trace(uint64_t* host_out, int32_t n) {
  return debug_read_trace(reinterpret_cast<unsigned long long*>(host_out), n);
}
int gtos_attn_decode(int32_t Hyp, int32_t L, int32_t H, int32_t hd, const float* q, int64_t ldq, const void* kv,
                     int64_t ld_kv, int32_t v_off, int64_t row_stride, const int32_t* slot, int64_t slot_ld,
                     const uint8_t* key_pad, int64_t pad_ld, float scale, float* out, int64_t ldo, void* out_bf16,
                     int64_t ldob, float* probs, void* stream) {
  return attn_decode(Hyp, L, H, hd, q, ldq, kv, ld_kv, v_off, row_stride, slot, slot_ld, key_pad, pad_ld, scale, out, ldo,
                     out_bf16, ldob, probs, S(stream));
}
int gtos_beam_ancestry(const int32_t* old_anc, int32_t* new_anc, int64_t ld, const int32_t* parent, int32_t t, int32_t Hyp,
                       void* stream) {
  return beam_ancestry(old_anc, new_anc, ld, parent, t, Hyp, S(stream));
}
int gtos_token_logprob(const float* logits, int64_t ldl, int32_t V, const float* gate_logits, const float* align, int32_t S,
                       const int64_t* copy_seq, int32_t Bsrc, const int32_t* src_index, int64_t rows, int32_t B,
                       float* table, int64_t ldt, int32_t W, void* stream) {
  return token_logprob(logits, ldl, V, gate_logits, align, S, reinterpret_cast<const long long*>(copy_seq), Bsrc, src_index,
                       rows, B, table, ldt, W, S_(stream));
}
int gtos_token_topk(const float* logits, int64_t ldl, int32_t V, const float* gate_logits, const float* align, int32_t S,
                    const int64_t* copy_seq, int32_t Bsrc, const int32_t* src_index, int64_t rows, int32_t B, int32_t W,
                    int32_t K, float* top_val, int32_t* top_idx, float* table, int64_t ldt, void* stream) {
  return token_topk(logits, ldl, V, gate_logits, align, S, reinterpret_cast<const long long*>(copy_seq), Bsrc, src_index, rows,
                    B, W, K, top_val, top_idx, table, ldt, S_(stream));
}
int gtos_beam_update(int32_t B, int32_t K, int32_t t, int32_t Tmin, int32_t Tmax, int32_t end_id, int32_t unk_id,
                     const float* top_val, const int32_t* top_idx, float* score, uint8_t* live, int32_t* n_done,
                     int32_t* steps, int32_t* tok, int32_t* par, float* done_score, int32_t* done_step, int32_t* done_par,
                     int32_t* parent_out, int64_t* last_tok, void* stream) {
  return beam_update_c(B, K, t, Tmin, Tmax, end_id, unk_id, top_val, top_idx, score, live, n_done, steps, tok, par, done_score,
                       done_step, done_par, parent_out, reinterpret_cast<long long*>(last_tok), S(stream));
}
int64_t gtos_grad_sumsq_workspace(void) { return grad_sumsq_workspace(); }
int gtos_grad_sumsq(const float* g, int64_t n, float* out, float* workspace, void* stream) {
  return grad_sumsq(g, n, out, workspace, S(stream));
}
int gtos_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int64_t n_decay, const float* lr_ptr, float beta1,
                   float beta2, float eps, float weight_decay, const float* norm_sq, float max_norm, void* stream) {
  return adam_step(p, g, m, v, n, n_decay, lr_ptr, beta1, beta2, eps, weight_decay, norm_sq, max_norm, S(stream));
}


int gtos_graph_paths(const int32_t* n_nodes, const int32_t* deg, const int32_t* nbr, const int32_t* lab, int32_t B,
                     int32_t n_max, int32_t deg_max, int32_t max_len, int32_t self_id, int32_t tl_id, const void* seed_ptr,
                     uint64_t seed_off, int32_t* paths, int32_t* plen, void* stream) {
  GraphPathsArgs a;
  a.n_nodes = n_nodes; a.deg = deg; a.nbr = nbr; a.lab = lab;
  a.B = B; a.n_max = n_max; a.deg_max = deg_max; a.max_len = max_len;
  a.self_id = self_id; a.tl_id = tl_id; a.seed = 0;
  a.paths = paths; a.plen = plen;
  return graph_paths(a, seed_ptr, seed_off, S(stream));
}

int gtos_graph_all_paths(const int32_t* n_nodes, const int32_t* deg, const int32_t* nbr, const int32_t* lab, int32_t B,
                         int32_t n_max, int32_t deg_max, int32_t max_len, int32_t K, int32_t self_id, int32_t tl_id,
                         int32_t* all_paths, int32_t* pcount, void* stream) {
  GraphAllPathsArgs a;
  a.n_nodes = n_nodes; a.deg = deg; a.nbr = nbr; a.lab = lab;
  a.B = B; a.n_max = n_max; a.deg_max = deg_max; a.max_len = max_len; a.K = K;
  a.self_id = self_id; a.tl_id = tl_id;
  a.all_paths = all_paths; a.pcount = pcount;
  return graph_all_paths(a, S(stream));
}

int gtos_graph_bfs(const int32_t* n_nodes, const int32_t* deg, const int32_t* nbr, const int32_t* root, int32_t B, int32_t n_max,
                   int32_t deg_max, int32_t* order, int32_t* depth, int32_t* pos, int32_t* reached, void* stream) {
  GraphBfsArgs a;
  a.n_nodes = n_nodes; a.deg = deg; a.nbr = nbr; a.root = root;
  a.B = B; a.n_max = n_max; a.deg_max = deg_max;
  a.order = order; a.depth = depth; a.pos = pos; a.reached = reached;
  return graph_bfs(a, S(stream));
}

}  // extern "C"
