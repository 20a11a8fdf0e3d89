// gtos_b200 -- tcgen05 GEMM kernels (declarations shared between gemm.cu and api.cu)
#pragma once
#include "common.cuh"

namespace gtos {

// How the N x N x B pair grid of one encoder layer is cut into 128-row MMA tiles.
// A tile is (bj keys j) x (bi queries i) of ONE batch element b, fetched with a single 4-D
// TMA box from relation[j][i][b][:], so a tile touches only bi query rows and bj key rows
// of q / k (they are staged in shared memory for the epilogue).
struct RelTiling {
  int N, B, D, H, hd;
  int bi, bj;          // tile = bj x bi pairs, bi*bj <= 128
  int ni_blk, nj_blk;  // ceil(N/bi), ceil(N/bj)
  int tiles;           // B * nj_blk * ni_blk
  float scale;         // hd^-1/2
};
int choose_rel_tiling(RelTiling* t, int N, int B, int D, int H);

enum GemmMode : int {
  MODE_PLAIN = 0,  // C = A * B^T (+bias, relu), A,B 2-D K-major
  MODE_DREL = 1,   // like PLAIN, rows of A are tile-major pair rows, output rows scattered to relation layout
  MODE_SCORE = 2,  // A = relation tile (4-D box), epilogue: s = scale * <q + ra, k + rb>
  MODE_GRAD = 3,   // A = relation tile, epilogue: G = scale*ds * [k + rb | q + ra]  (bf16, tile-major)
  MODE_GRU = 4,    // one GRU time step: A = [x_t | h_prev], B = gate-interleaved weights, epilogue = gate math
};

struct GemmTnArgs {
  // problem
  const void* A;       // bf16; PLAIN/DREL: [M,K] row-major (lda); SCORE/GRAD: relation bf16 [N,N,B,D]
  long lda;
  const void* Bm;      // bf16 [N,K] row-major (ldb)
  long ldb;
  int M, N, K;
  // plain epilogue
  const float* bias;   // [N] or null
  float* out_f32;      // [M,N] (ldo) or null
  long ldo;
  void* out_bf16;      // [M,N] (ldob) or null
  long ldob;
  int relu;
  int accumulate;      // out_f32 += result
  const float* addend; // optional fp32 [M,N] (ldadd) added in the epilogue: out = A*B^T + bias + addend
  long ldadd;
  // relation modes
  RelTiling rt;
  const void* q;       // bf16 [N,B,D] (bias included, unscaled), row stride ldqk elements
  const void* k;       // bf16 [N,B,D]
  long ldqk;
  float* scores;       // SCORE out: [B,H,N(j),N(i)]  (null when the attention tail is fused)
  // SCORE with the attention tail fused into the epilogue (fuse != 0; see rel_attn_fusable)
  int fuse;
  const void* v;       // bf16 [N,B,D] projected values, row stride ldv elements
  long ldv;
  const uint8_t* key_pad;                // [N,B] or null
  float p_drop; const void* seed_ptr; unsigned long long seed_off;
  float* probs;        // [B,H,N(i),N(j)] softmax (pre-dropout), saved for the backward
  float* probs_dropped;                  // optional: the weights after dropout (need_weights)
  float* att;          // [N*B, D] attention output, row stride ldatt
  long ldatt;
  void* att_bf16;      // optional bf16 copy (same row stride)
  const float* dscores;  // GRAD in : [B,H,N(j),N(i)]
  void* G;             // GRAD out: bf16 [tiles*128, 2D] (permuted feature order)
};
int launch_gemm_tn(int mode, const GemmTnArgs& a, cudaStream_t stream);
// the fused relation-attention forward needs a tile that holds every key of its queries and one head per epilogue warpgroup
bool rel_attn_fusable(const RelTiling& rt);
int debug_read_trace(unsigned long long* host_out, int n);   // GTOS_DBG=2 timestamps, 16 slots per CTA

// One packed-sequence GRU step (generator/encoder.py:105-106, nn.GRU cell) as a GEMM with a fused gate epilogue.
// Wcat rows are interleaved per block of UB = BN/4 hidden units: [r | z | n_input | n_hidden], K = [x (Kx) | h (H)].
struct GruStepArgs {
  const void* x;  long ldx;  int Kin;     // bf16 [R, ldx] input of this time step (Kin valid columns)
  const void* hb; long ldhb;              // bf16 [R, H] previous hidden state (MMA operand)
  const float* h_prev;                    // fp32 [R, H]
  const void* Wcat; long ldw; int Kx;     // bf16 [4H, Kx + H8], Kx = 64*ceil(Kin/64)
  const float* bcat;                      // fp32 [4H] interleaved like Wcat rows
  const long long* lengths; int t;        // packed-sequence masking: row live iff lengths[row] > t
  float* h_new; void* hb_new; long ldhbn; // fp32 / bf16 [R, H]
  void* out_t; long ldout;                // bf16 layer output at time t (zero for finished rows), may be null
  void* gates; long ldg;                  // bf16 [R, 4H] saved (r, z, n, W_hn h + b_hn) for backward
  int R, H;
};
int launch_gru_step(const GruStepArgs& a, cudaStream_t stream);

// out[M,N] = sum_k A[k,m] * B[k,n]   (both operands MN-major), split-K with fp32 partials
struct GemmNnArgs {
  const void* A;       // bf16 [Kd, M] row-major (lda)
  long lda;
  const void* Bm;      // bf16 [Kd, N] row-major (ldb), or relation bf16 [N,N,B,D] when rel != 0
  long ldb;
  int M, N, Kd;
  int rel;             // B rows come from 4-D relation tiles (Kd = tiles*128)
  int perm_D, perm_hd; // != 0 (or rel): accumulator row m is a permuted relation_in_proj row -> written to its reference row
  RelTiling rt;
  float* out;          // [M,N] fp32 (ldo)  (rel: rows un-permuted to the reference [2D,D] layout)
  long ldo;
  float* workspace;    // >= splits*M*N floats
  long workspace_elems;
};
// bank-factorised relation attention (rel_banked.cu)
struct RelBankedArgs {
  const void* PB; long ldpb;                 // bf16 [R, 2D] = bank * Wperm^T (head-interleaved [ra_h | rb_h] columns)
  const long long* idx;                      // int64 [N,N,B]: idx[j][i][b] = bank row of the pair (query i, key j)
  const void* q; const void* k; long ldqk;   // bf16 [N*B, .] projected queries / keys
  const float* v; long ldv;                  // fp32 [N*B, D] projected values
  const uint8_t* key_pad; const uint8_t* attn_mask;
  float p_drop; const void* seed_ptr; unsigned long long seed_off;
  float* probs; float* probs_dropped;        // [B,H,N,N]
  float* out; long ldo; void* out_bf16;      // [N*B, D]
  const float* dscores; void* G;             // backward: d(score) [B,H,N(j),N(i)] -> G [tiles*128, 2D] bf16
  int N, B, D, H, R;
};
int rel_attn_banked_fwd(const RelBankedArgs& a, cudaStream_t st);
int rel_grad_banked(const RelBankedArgs& a, cudaStream_t st);
long gemm_nn_workspace_elems(int M, int N, int Kd, int rel);
int set_sm_reserve(int n);   // persistent GEMM grids use (SM count - n) CTAs
int launch_gemm_nn(const GemmNnArgs& a, cudaStream_t stream);

// permuted row order of relation_in_proj.weight: per head h, [ra_h (hd rows) | rb_h (hd rows)]
__host__ __device__ inline int rel_perm_to_orig(int pr, int D, int hd) {
  int h = pr / (2 * hd), w = pr % (2 * hd);
  return w < hd ? h * hd + w : D + h * hd + (w - hd);
}

}  // namespace gtos
