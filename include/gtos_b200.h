/* gtos_b200 -- C ABI of libgtos_b200.so: the B200 (sm_100a) kernels behind the gtos graph-transformer
 * encode/decode hot path.
 *
 * The reference (jcyk/gtos) is pure PyTorch: its "FFI" for this path is the ATen call sites inside
 * four nn.Module.forward methods.  Each entry point below names the reference lines it replaces
 * (paths relative to /root/reference/generator; translator/ holds byte-identical copies).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error (see GTOS_ERR_*); the message is
 *     available from gtos_last_error() (thread-local); nothing throws, nothing allocates;
 *   - all pointers are DEVICE pointers unless noted; outputs and workspaces are caller-allocated;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host sync;
 *   - fp32 tensors are row-major with explicit row strides (`ld*`, in elements);
 *     "bf16" buffers hold __nv_bfloat16; activations are time-major [len, batch, dim] as in the
 *     reference (graph_transformer.py:50, transformer.py:99);
 *   - dropout is counter-based: keep(idx) = hash(*seed_ptr + seed_off, idx) >= p, so backward
 *     regenerates the mask from the same (seed_ptr, seed_off) and CUDA graphs can bump the seed.
 */
#ifndef GTOS_B200_H_
#define GTOS_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTOS_OK 0
#define GTOS_ERR_CUDA 1
#define GTOS_ERR_ARG 2
#define GTOS_ERR_UNSUPPORTED 3
#define GTOS_ERR_NO_DEVICE 4

const char* gtos_last_error(void);
int gtos_abi_version(void);
/* number of CUDA kernels this library has enqueued so far in this process (bench.py's gpu_launches) */
uint64_t gtos_launch_count(void);
/* The persistent tcgen05 GEMMs launch one CTA per SM, and each CTA owns its SM (~200 KB of shared memory).  A collective
 * that runs BESIDE them (data-parallel gradient buckets reduced during the backward pass, train.py:74-79) needs SMs of
 * its own, or every GEMM launched meanwhile waits a second wave for the CTAs NCCL displaced.  n SMs are left free
 * (default 0, or the GTOS_SM_RESERVE environment variable). */
int gtos_set_sm_reserve(int32_t n);
/* 0 if the current device is sm_100 and the TMA driver entry point resolves */
int gtos_device_check(void);
/* timing experiments only: with GTOS_DBG=2 in the environment the plain GEMM records 16 clock64 timestamps per CTA
 * (phases of its pipeline); copies the first n of them (148 x 16) to HOST memory */
int gtos_debug_read_trace(uint64_t* host_out, int32_t n);
/* same for the attention core: copies 3 x 16 timestamps (fwd, bwd_q, bwd_kv; CTA 0) to HOST memory, then switches the
 * recording on (enable != 0) or off */
int gtos_debug_attn_trace(uint64_t* host_out, int32_t enable);

/* ---- operand staging -------------------------------------------------------------------------- */
/* fp32 [rows, cols] (lds) -> bf16 [rows, ldd]; columns cols..ldd-1 are zero-filled (TMA needs 16 B rows) */
int gtos_cast_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int32_t cols, void* stream);
/* same cast plus sums[c] = sum_r src[r,c] in one pass (operand copy + bias gradient of a Linear backward) */
int gtos_cast_colsum(const float* src, int64_t lds, void* dst, int64_t ldd, float* sums, int64_t rows, int32_t cols,
                     void* stream);
/* weight W fp32 [R,C] -> Wb bf16 [R, ldw] and/or Wt bf16 [C, ldt] (transpose, for input gradients).
 * rel_heads > 0: rows are reordered head-interleaved, per head [ra_h | rb_h], for
 * relation_in_proj.weight [2D, D] (graph_transformer.py:80,122) */
int gtos_weight_prep(const float* W, int32_t R, int32_t C, void* Wb, int64_t ldw, void* Wt, int64_t ldt,
                     int32_t rel_heads, void* stream);

/* ---- dense projections (F.linear call sites: graph_transformer.py:191-197,165; transformer.py:175-196,162;
 *      fc1/fc2 graph_transformer.py:60-63, transformer.py:66-69; nn.GRU / out_proj encoder.py:106,117) ------ */
/* C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (ReLU);  A,B bf16;  writes fp32 and/or bf16;  accumulate: C_f32 += */
int gtos_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, float* out_f32, int64_t ldo,
                 void* out_bf16, int64_t ldob, int32_t M, int32_t N, int32_t K, int32_t relu, int32_t accumulate,
                 void* stream);
/* C[M,N] (fp32) = A * B^T + bias[N] + addend[M,N]   (e.g. dh_prev = dgh * W_hh + dh * z in the GRU backward) */
int gtos_gemm_tn_add(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, const float* addend,
                     int64_t ldadd, float* out_f32, int64_t ldo, int32_t M, int32_t N, int32_t K, void* stream);
/* C[M,N] = sum_k A[k,m] * B[k,n]  (weight gradients, autograd of the call sites above) */
int64_t gtos_gemm_nn_workspace(int32_t M, int32_t N, int32_t Kd);
int gtos_gemm_nn(const void* A, int64_t lda, const void* B, int64_t ldb, float* out, int64_t ldo, int32_t M, int32_t N,
                 int32_t Kd, float* workspace, int64_t workspace_elems, void* stream);

/* ---- fused relation attention (RelationMultiheadAttention.forward, graph_transformer.py:93-174) ---- */
/* tiling of the N x N x B pair grid into 128-row MMA tiles: out = {bi, bj, ni_blk, nj_blk, tiles} */
int gtos_rel_tiling(int32_t N, int32_t B, int32_t D, int32_t H, int32_t* out5);
/* scores[b,h,j,i] = hd^-1/2 < q[i,b,h] + Wa r[j,i,b] , k[j,b,h] + Wb r[j,i,b] >     (:122-133)
 * relb: relation as bf16 [N,N,B,D]; Wperm: gtos_weight_prep(rel_heads=H) output [2D,D];
 * q,k: the projected queries / keys as bf16 [N,B,D] with row stride ldqk elements (staged by TMA for the epilogue) */
int gtos_rel_score(const void* relb, const void* Wperm, const void* q, const void* k, int64_t ldqk, float* scores,
                   int32_t N, int32_t B, int32_t D, int32_t H, void* stream);
/* backward of the above w.r.t. the per-pair projections: G[tile-major pair, 2D] (bf16) = hd^-1/2 dscores * [k+rb | q+ra] */
/* graph_transformer.py:122-159 as ONE kernel on the dense relation tensor (the north star's fused relation attention):
 * gtos_rel_score's projection GEMM + score epilogue, then - in the same epilogue, on the tile's complete softmax rows - the
 * key-padding mask, softmax over keys, dropout (same counter-based draw as gtos_attn_fwd) and o_i = sum_j w_ij v_j.  Scores
 * never reach HBM.  v: bf16 [N*B, .] projected values (row stride ldv); probs [B,H,N,N] pre-dropout (saved for
 * gtos_attn_bwd); att fp32 [N*B, D] (+ optional bf16 copy, same stride).  Needs every key of a query inside one 128-pair tile
 * and head_dim 64 (gtos_rel_attn_fusable: configs 2 and 3; not the 257-node stress config, not an attention mask) -
 * otherwise GTOS_ERR_UNSUPPORTED and the caller runs gtos_rel_score + gtos_attn_fwd. */
int gtos_rel_attn_fusable(int32_t N, int32_t B, int32_t D, int32_t H);
int gtos_rel_attn_fwd(const void* relb, const void* Wperm, const void* q, const void* k, int64_t ldqk, const void* v,
                      int64_t ldv, const uint8_t* key_pad, float p_drop, const void* seed_ptr, uint64_t seed_off,
                      float* probs, float* probs_dropped, float* att, int64_t ldatt, void* att_bf16, int32_t N, int32_t B,
                      int32_t D, int32_t H, void* stream);
int gtos_rel_grad(const void* relb, const void* Wperm, const void* q, const void* k, int64_t ldqk,
                  const float* dscores, void* G, int32_t N, int32_t B, int32_t D, int32_t H, void* stream);
/* d_relation[j,i,b,:] (+)= G * Wperm  (WpermT = transposed prep output [D,2D]) */
int gtos_rel_drel(const void* G, const void* WpermT, float* d_relation, int32_t accumulate, int32_t N, int32_t B,
                  int32_t D, int32_t H, void* stream);
/* d relation_in_proj.weight [2D,D] (reference row order) = G^T * relation */
int64_t gtos_rel_dw_workspace(int32_t N, int32_t B, int32_t D, int32_t H);
int gtos_rel_dw(const void* G, const void* relb, float* dW, float* workspace, int64_t workspace_elems, int32_t N,
                int32_t B, int32_t D, int32_t H, void* stream);
/* dq[i,b,:] = sum_j G_x, dk[j,b,:] = sum_i G_y ; written with row stride ld (into the QKV grad buffer).
 * dq_bf16 / dk_bf16: optional bf16 copies with the same element layout (operand of the in_proj backward GEMMs). */
int gtos_rel_dqk(const void* G, float* dq, float* dk, int64_t ld, void* dq_bf16, void* dk_bf16, int32_t N, int32_t B,
                 int32_t D, int32_t H, void* stream);

/* ---- bank-factorised relation attention, forward half (caller generator/generator.py:76-90) ----
 * relation = bank[idx] and relation_in_proj has no bias (graph_transformer.py:80), so [ra_ij | rb_ij] is row idx[j][i][b]
 * of PB = bank * Wperm^T (bf16 [R, 2D], head-interleaved columns; one R-row gtos_gemm_tn per layer).
 * gtos_rel_attn_banked_fwd is graph_transformer.py:122-159 as ONE kernel: gather, per-head scores hd^-1/2 <q_i + ra, k_j + rb>,
 * key-padding [N,B] / attention [N,N] masks, softmax over keys, dropout (same counter-based draw as gtos_attn_fwd, so
 * gtos_attn_bwd replays it), o_i = sum_j w_ij v_j.  probs [B,H,N,N] (pre-dropout, saved for the backward), out fp32
 * [N*B, D] (+ bf16 copy).  gtos_rel_grad_banked writes the per-pair gradient rows G exactly as gtos_rel_grad lays them
 * out (tile-major rows, head-interleaved [d(q+ra) | d(k+rb)] columns) from the same gather instead of a P-row GEMM. */
int gtos_rel_attn_banked_fwd(const void* PB, int64_t ldpb, const int64_t* idx, const void* q, const void* k, int64_t ldqk,
                             const float* v, int64_t ldv, const uint8_t* key_pad, const uint8_t* attn_mask, float p_drop,
                             const void* seed_ptr, uint64_t seed_off, float* probs, float* probs_dropped, float* out,
                             int64_t ldo, void* out_bf16, int32_t N, int32_t B, int32_t D, int32_t H, int32_t R,
                             void* stream);
int gtos_rel_grad_banked(const void* PB, int64_t ldpb, const int64_t* idx, const void* q, const void* k, int64_t ldqk,
                         const float* dscores, void* G, int32_t N, int32_t B, int32_t D, int32_t H, int32_t R,
                         void* stream);

/* ---- bank-factorised backward of the relation terms (caller generator/generator.py:76-79) ----
 * When relation = bank[idx] (bank [R,D], idx [N,N,B] int64, layout idx[j][i][b]) the two P-row GEMMs of the backward
 * collapse to R-row GEMMs after ONE segmented sum of the per-pair gradient rows G (from gtos_rel_grad):
 *   S[r,:] = sum_{pairs p : idx_p = r} G[p,:]      d relation_in_proj.weight = S^T bank      d bank = S * Wperm
 * gtos_rel_pair_keys: keys[g] = bank row of G row g (tile-major pair rows, gtos_rel_tiling), R for tile padding rows.
 *   The caller sorts (keys, g) once per batch -> `keys` ascending, `order` = the matching G rows; n = N*N*B.
 * gtos_rel_segsum: S (bf16 [R,C], row stride ldo >= C = 2D) from the sorted lists; rows without pairs are NOT
 *   written (zero them once); spill = fp32 scratch [R,C] (only rows whose pairs straddle a 32-pair window are touched).
 *   The row stride lets the L layers of an encoder write side by side, so d bank = [S_1|..|S_L] * [Wperm_1;..;Wperm_L]
 *   is ONE gtos_gemm_tn at the end of the backward.
 * gtos_rel_dw_bank: dW [2D,D] fp32 (reference row order) = S^T * bank_bf16 [R,D]; S row stride lds. */
int gtos_rel_pair_keys(const int64_t* idx, int32_t N, int32_t B, int32_t D, int32_t H, int32_t R, int32_t* keys,
                       void* stream);
int gtos_rel_segsum(const void* G, const int32_t* order, const int32_t* keys, int64_t n, int32_t C, void* out_bf16,
                    int64_t ldo, float* spill, void* stream);
int gtos_rel_dw_bank(const void* S, int64_t lds, const void* bankb, float* dW, int32_t R, int32_t D, int32_t H,
                     void* stream);

/* ---- attention core: masks, softmax, dropout, PV (graph_transformer.py:136-159; transformer.py:131-155) ---- */
typedef struct gtos_attn_desc {
  int32_t T, S, B, H, hd;
  int32_t bwd_part;  /* gtos_attn_bwd only: 0 = both kernels; 1 = query side only (dS, dq); 2 = key side only (dV, dK).
                      * The key side's dV needs nothing from the query side (its dK reads dscores_ts), so a caller that does
                      * not ask for dK (relation attention: dk comes from G) may run the two parts on different streams. */
  const float* q; int64_t ldq;          /* element (t,b,h,d) at q[(t*B+b)*ldq + h*hd + d]; NULL in encoder mode */
  const float* k; int64_t ldk;
  const float* v; int64_t ldv;
  float scale; float p_drop;
  const float* scores_jt;               /* encoder mode: scores from gtos_rel_score, [B,H,S,T] */
  const uint8_t* key_pad;               /* [S,B], 1 = padding (may be NULL) */
  const uint8_t* attn_mask;             /* [T,S], 1 = blocked (may be NULL) */
  const void* seed_ptr; uint64_t seed_off;
  float* probs;                         /* out (fwd) / in (bwd): softmax before dropout, [B,H,T,S] */
  float* probs_dropped;                 /* optional out: weights after dropout, [B,H,T,S] */
  float* out; int64_t ldo; void* out_bf16;
  /* backward only */
  const float* dout; int64_t lddo;
  const float* dprobs_extra;            /* optional grad w.r.t. probs_dropped */
  float* dscores_jt;                    /* encoder mode out [B,H,S,T] */
  float* dscores_ts;                    /* out [B,H,T,S] */
  float* dq; int64_t lddq;
  float* dk; int64_t lddk;
  float* dv; int64_t lddv;
  /* optional bf16 copies of dq / dk / dv (same element layout and row strides as the fp32 outputs): the operand of
   * the in_proj backward GEMMs, so no separate cast pass sits between this kernel and them */
  void* dq_bf16; void* dk_bf16; void* dv_bf16;
  /* fp32 mode (north star: 1e-3 against the fp32 reference): non-zero = every MMA operand is staged as a split-bf16 pair
   * (x_hi = bf16(x), x_lo = bf16(x - x_hi)) and every product runs as three tensor-core passes hi*hi + lo*hi + hi*lo
   * with fp32 accumulation; softmax with expf.  The bf16 output copies are still written when asked for. */
  int32_t precise;
} gtos_attn_desc;
int gtos_attn_fwd(const gtos_attn_desc* d, void* stream);
int gtos_attn_bwd(const gtos_attn_desc* d, void* stream);
/* 1 if, for this shape, the query-side kernel of gtos_attn_bwd (bwd_part 0 or 1, decoder mode with dq and dk) also writes
 * dK = scale dS^T q - it does when one CTA holds every query row of its (batch, head).  The key side is then only
 * dV = Pd^T dO, which depends on nothing the query side produces: a caller may launch bwd_part = 2 with dq = dk = NULL on a
 * second stream BESIDE bwd_part = 1 instead of after it.  0: the key side needs the query side's dscores_ts (run part 0). */
int gtos_attn_bwd_dk_on_query_side(const gtos_attn_desc* d);

/* ---- residual + dropout + LayerNorm (graph_transformer.py:57-58,64-65; transformer.py:56-57,63,70-71) ---- */
int gtos_add_ln_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y, void* y_bf16,
                    float* z, float* mean, float* rstd, int64_t rows, int32_t D, float p_drop, const void* seed_ptr,
                    uint64_t seed_off, void* stream);
int gtos_add_ln_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                    float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta, int64_t rows, int32_t D,
                    float p_drop, const void* seed_ptr, uint64_t seed_off, void* stream);
/* LayerNorm parameter gradients alone (the weight / bias gradients torch's LayerNorm backward returns for
 * graph_transformer.py:58,65): dgamma[d] = sum_rows dy * (z - mean) * rstd, dbeta[d] = sum_rows dy.  gtos_add_ln_bwd
 * with dgamma = dbeta = NULL computes the input gradients only, so the host can run this on a second stream. */
int gtos_ln_param_grad(const float* dy, const float* z, const float* mean, const float* rstd, float* dgamma, float* dbeta,
                       int64_t rows, int32_t D, void* stream);
/* out[n] = sum_m x[m,n]  (bias gradients) */
int gtos_colsum(const float* x, int64_t ld, float* out, int64_t rows, int32_t cols, void* stream);
int gtos_colsum_bf16(const void* x, int64_t ld, float* out, int64_t rows, int32_t cols, void* stream);
/* FFN hidden dropout (in place, bf16) and its ReLU/dropout backward (graph_transformer.py:60-61) */
int gtos_dropout_bf16(void* h, int64_t n, float p, const void* seed_ptr, uint64_t seed_off, void* stream);
/* generic fp32 dropout, out may alias x (decoder.py:82, transformer.py:156); same seed in backward */
int gtos_dropout_f32(const float* x, float* out, int64_t n, float p, const void* seed_ptr, uint64_t seed_off,
                     void* stream);
int gtos_relu_drop_bwd(const float* dh_in, const void* act_bf16, float* dh_f32, void* dh_bf16, int64_t n, float p,
                       void* stream);

/* ---- TokenGenerator training tail (decoder.py:42-64): copy/generate mixture NLL of the target token in ONE pass over
 * the vocabulary logits [rows = T*B, V] (row = t*B + b):
 *   p = softmax(gate_logits)[0] * softmax(logits)[target] + softmax(gate_logits)[1] * sum_s align[row,s] * [copy_seq[s,b] == target]
 *   loss_row = -log(p + 1e-12), 0 where target == pad_idx;   stats: 6 floats per row kept for backward.
 * backward writes dlogits [rows, V], dgate_logits [rows, 2], dalign [rows, S]. */
int gtos_token_nll_fwd(const float* logits, int64_t ldl, int32_t V, const float* gate_logits, const float* align,
                       int32_t S, const int64_t* copy_seq, const int64_t* target, int64_t rows, int32_t B, int64_t pad_idx,
                       float* loss_row, float* stats, void* stream);
int gtos_token_nll_bwd(const float* dloss_row, const float* logits, int64_t ldl, int32_t V, const float* align, int32_t S,
                       const int64_t* copy_seq, const int64_t* target, int64_t rows, int32_t B, int64_t pad_idx,
                       const float* stats, float* dlogits, int64_t lddl, float* dgate_logits, float* dalign,
                       void* dlogits_bf16 /* optional bf16 copy of dlogits, row stride lddb */, int64_t lddb, void* stream);

/* ---- bank -> dense relation gather (generator.py:79: relation = bank.index_select(0, idx)) and its backward ----
 * forward emits the fp32 [P,D] tensor of the caller's contract and (optionally) the bf16 copy the fused kernels read;
 * backward zero-fills d_bank [R,D] and scatter-adds the dense gradient with 16-byte vector reductions. */
int gtos_bank_gather(const float* bank, const int64_t* idx, int64_t P, int32_t D, float* out_f32, void* out_bf16,
                     void* stream);
int gtos_bank_scatter_add(const float* d_rel, const int64_t* idx, int64_t P, int32_t D, float* d_bank, int64_t R,
                          void* stream);
/* the same backward from pairs SORTED by bank row (order = pair indices, keys = idx[order], both int64 [P]; one stable sort
 * per batch): running sums in registers, one vector reduction per (bank row, 32-pair window) instead of one per pair - the
 * plain scatter serialises on hot rows (config 2: the <TL> path holds 41 % of the pairs). */
int gtos_bank_segsum(const float* d_rel, const int64_t* order, const int64_t* keys, int64_t P, int32_t D, float* d_bank,
                     int64_t R, void* stream);
/* evaluation batches (generator.py:83-88: `relation[0,:] = 0; relation[idx].sum(3) / count(idx != 0).clamp(min=1)`):
 * idx [P,K] with 0 = empty slot; out[p] = mean of the bank rows of the pair's shortest paths (fp32 and / or bf16). */
int gtos_bank_gather_mean(const float* bank, const int64_t* idx, int64_t P, int32_t K, int32_t D, float* out_f32,
                          void* out_bf16, void* stream);

/* ---- RelationEncoder (encoder.py:90-119): embedding + GRU gate math; GEMMs via gtos_gemm_* ---- */
int gtos_embed_gather(const float* table, const int64_t* idx, int64_t n, int32_t dim, float* out_f32, void* out_bf16,
                      int64_t ldb, float p_drop, const void* seed_ptr, uint64_t seed_off, void* stream);
int gtos_embed_scatter_add(const float* dx, const int64_t* idx, int64_t n, int32_t dim, float* dtable, float p_drop,
                           const void* seed_ptr, uint64_t seed_off, void* stream);
/* gate-interleaved weights for the fused step: Wcat bf16 [4H, Kx + H] (Kx = 64*ceil(Kin/64)), bcat fp32 [4H] */
int gtos_gru_weight_prep(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int32_t Kin,
                         int32_t H, int32_t Kx, void* Wcat, int64_t ldw, float* bcat, void* stream);
/* ONE packed-sequence GRU time step (nn.GRU cell, gates r,z,n; encoder.py:105-106): tcgen05 GEMM
 * [x_t | h_prev] x Wcat^T with the gate math, length masking and all stores fused in the epilogue.
 * gates: bf16 [R,4H] saved for backward (r, z, n, W_hn h + b_hn).  out_t: bf16 layer output (0 for finished rows). */
int gtos_gru_step_fwd(const void* x, int64_t ldx, int32_t Kin, const void* hb, int64_t ldhb, const float* h_prev,
                      const void* Wcat, int64_t ldw, int32_t Kx, const float* bcat, const int64_t* lengths, int32_t t,
                      float* h_new, void* hb_new, int64_t ldhbn, void* out_t, int64_t ldout, void* gates, int64_t ldg,
                      int64_t R, int32_t H, void* stream);
/* db_ih / db_hh (fp32 [3H], may be NULL): bias gradients, ACCUMULATED over the calls of one (layer, direction) */
int gtos_gru_gate_bwd(const float* dh, const float* dout_t, int64_t lddout, const void* gates, const float* h_prev,
                      const int64_t* lengths, int32_t t, float* dh_prev, void* dgi_bf16, int64_t lddgi, void* dgh_bf16,
                      int64_t lddgh, float* db_ih, float* db_hh, int64_t R, int32_t Hh, void* stream);

/* zero n byte ranges of device memory with one launch per 48 ranges.  ptrs / bytes are HOST arrays of n entries (the
 * device pointers and their lengths are passed to the kernel by value).  RelationEncoder's length-sorted schedule
 * (encoder.py:93-99: like pack_padded_sequence, step t runs on the paths longer than t only) clears ~75 short row ranges
 * per training step that no kernel writes but a later GEMM reads. */
int gtos_zero_regions(void* const* ptrs, const int64_t* bytes, int32_t n, void* stream);

/* ---- fp32 mode (BASELINE north star: "within 1e-3 fp32"; the reference computes in fp32 end to end,
 *      graph_transformer.py:122-133, transformer.py:111-162, encoder.py:90-119) ----
 * Every matrix product still runs on the tcgen05 GEMMs above, on SPLIT operands: x = x_hi + x_lo, x_hi = bf16(x),
 * x_lo = bf16(x - x_hi), and A B^T ~= A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T is ONE gtos_gemm_tn call with K tripled:
 * [A_hi | A_lo | A_hi] (role 0) against [B_hi | B_hi | B_lo] (role 1), fp32 accumulation (error ~2^-17 relative).
 * gtos_split3 stages such an operand: src fp32, element (r, c) at src[r * ld_r + c * ld_c] (a transposed view is free),
 * dst bf16 [rows, 3 kp] with row stride ldd; kp = cols rounded up to 8, pad columns zero.  With ldd == 3 kp the same buffer
 * viewed as [3 rows, kp] is the row-stacked operand of gtos_gemm_nn (weight gradients), pairing role 0 with role 1.
 * The attention core has the same three-pass variant (gtos_attn_desc.precise). */
int gtos_split3(const float* src, int64_t ld_r, int64_t ld_c, int64_t rows, int32_t cols, void* dst, int64_t ldd,
                int32_t kp, int32_t role, void* stream);
/* relation scores and their gradient on fp32 values (graph_transformer.py:122-133).  PR fp32 [P = N*N*B, 2D] (row stride
 * ldpr) = relation_in_proj(relation), reference column order [ra | rb], row (j*N + i)*B + b (query i, key j);
 * q / k fp32 rows (n*B + b) with row stride ldqk.
 *   scores[b,h,j,i] = hd^-1/2 <q[i,b,h] + ra, k[j,b,h] + rb>              (layout of gtos_attn_desc.scores_jt)
 *   G[p] = hd^-1/2 dscores[b,h,j,i] * [k + rb | q + ra]                   (fp32 [P, 2D]: d ra | d rb)
 *   dq[i,b] = sum_j G[(j,i,b), 0:D]      dk[j,b] = sum_i G[(j,i,b), D:2D]   (fixed summation order) */
int gtos_rel_score_f32(const float* PR, int64_t ldpr, const float* q, const float* k, int64_t ldqk, float* scores,
                       int32_t N, int32_t B, int32_t D, int32_t H, void* stream);
int gtos_rel_grad_f32(const float* PR, int64_t ldpr, const float* q, const float* k, int64_t ldqk, const float* dscores,
                      float* G, int64_t ldg, int32_t N, int32_t B, int32_t D, int32_t H, void* stream);
int gtos_rel_dqk_f32(const float* G, int64_t ldg, float* dq, float* dk, int64_t ld, int32_t N, int32_t B, int32_t D,
                     void* stream);
/* FFN backward through dropout(relu(.)) from the fp32 post-dropout activation (graph_transformer.py:60-61) */
int gtos_relu_drop_bwd_f32(const float* dh_in, const float* act, float* dh_out, int64_t n, float p, void* stream);
/* GRU cell on fp32 gate pre-activations (nn.GRU r,z,n; encoder.py:105-106): gi = x_t W_ih^T + b_ih, gh = h W_hh^T + b_hh
 * (both [R, 3H], from K-tripled GEMMs); rows with lengths[row] <= t keep h and emit a zero output.
 * gates fp32 [R, 4H] = [r | z | n | gh_n].  Backward: dgi / dgh fp32 [R, 3H], dh_part = (dh + dout_t) * z. */
int gtos_gru_gate_fwd_f32(const float* gi, int64_t ldgi, const float* gh, int64_t ldgh, const float* h_prev,
                          const int64_t* lengths, int32_t t, float* h_new, float* out_t, int64_t ldout, float* gates,
                          int64_t R, int32_t H, void* stream);
int gtos_gru_gate_bwd_f32(const float* dh, const float* dout_t, int64_t lddout, const float* gates, const float* h_prev,
                          const int64_t* lengths, int32_t t, float* dh_part, float* dgi, int64_t lddgi, float* dgh,
                          int64_t lddgh, int64_t R, int32_t H, void* stream);

/* ---- incremental beam decode (callers generator/generator.py:120-167, generator/search.py:57-168) ----
 * Single-query attention over a bf16 K/V cache (MultiheadAttention.forward with T_q = 1, transformer.py:98-173).
 * Cache row of (position l, slot s) = l * row_stride + s; a row holds K at column 0 and V at column v_off (elements),
 * row pitch ld_kv.  Hypothesis h reads slot = slot[l * slot_ld + h] (NULL: slot = h):
 *   graph cross-attention: slot = source graph of h, slot_ld = 0  (K/V of the graph memory projected once per graph);
 *   token self-attention : slot = ancestry table anc[l][h] (gtos_beam_ancestry), slot_ld = its row pitch.
 * key_pad[l * pad_ld + slot] = 1 masks a key.  q fp32 [Hyp, H*hd] unscaled; out fp32 and/or bf16; probs optional
 * fp32 [Hyp, H, L] (the alignment weights of TokenGenerator, decoder.py:32-36). */
int gtos_attn_decode(int32_t Hyp, int32_t L, int32_t H, int32_t hd, const float* q, int64_t ldq, const void* kv,
                     int64_t ld_kv, int32_t v_off, int64_t row_stride, const int32_t* slot, int64_t slot_ld,
                     const uint8_t* key_pad, int64_t pad_ld, float scale, float* out, int64_t ldo, void* out_bf16,
                     int64_t ldob, float* probs, void* stream);
/* new_anc[l][h] = old_anc[l][parent[h]] for l < t, new_anc[t][h] = h  (tables int32 [Tmax, ld]; parent NULL = identity).
 * Replaces the per-step index_select of every cached state tensor (search.py:72-76). */
int gtos_beam_ancestry(const int32_t* old_anc, int32_t* new_anc, int64_t ld, const int32_t* parent, int32_t t, int32_t Hyp,
                       void* stream);
/* work=True tail of TokenGenerator (decoder.py:42-59): table[row, 0..W) = log(gen * softmax(logits) (+0 beyond V)
 * + copy * scatter(align by copy_seq[:, b(row)]) + 1e-12);  b(row) = src_index[row], or row % B when src_index is NULL.
 * copy_seq int64 [S, Bsrc]. */
int gtos_token_logprob(const float* logits, int64_t ldl, int32_t V, const float* gate_logits, const float* align, int32_t S,
                       const int64_t* copy_seq, int32_t Bsrc, const int32_t* src_index, int64_t rows, int32_t B,
                       float* table, int64_t ldt, int32_t W, void* stream);
/* Same row computed in shared memory with the k best tokens taken in the same kernel (generator.py:157 torch.topk):
 * top_val / top_idx [rows, K] (log-probs best first, ties -> lowest id); table optional (NULL: never written to HBM). */
int gtos_token_topk(const float* logits, int64_t ldl, int32_t V, const float* gate_logits, const float* align, int32_t S,
                    const int64_t* copy_seq, int32_t Bsrc, const int32_t* src_index, int64_t rows, int32_t B, int32_t W,
                    int32_t K, float* top_val, int32_t* top_idx, float* table, int64_t ldt, void* stream);
/* One Beam.update (search.py:57-92) for all B beams of K slots from the top-k lists of step t: candidates = live slot x
 * rank, stable descending merge, keep K - n_done, <END> completes when t >= Tmin (else dropped), UNK scores -inf.
 * State (updated in place): score f32 [B,K], live u8 [B,K], n_done / steps i32 [B], back-pointers tok / par i32
 * [Tmax,B,K], completed table done_score f32 / done_step / done_par i32 [B,K].  Outputs for the next step:
 * parent_out i32 [B*K] (row of the prefix in this step's arrangement), last_tok i64 [B*K]. */
int gtos_beam_update(int32_t B, int32_t K, int32_t t, int32_t Tmin, int32_t Tmax, int32_t end_id, int32_t unk_id,
                     const float* top_val, const int32_t* top_idx, float* score, uint8_t* live, int32_t* n_done,
                     int32_t* steps, int32_t* tok, int32_t* par, float* done_score, int32_t* done_step, int32_t* done_par,
                     int32_t* parent_out, int64_t* last_tok, void* stream);

/* ---- optimizer step on flat buffers (generator/adam.py:28-87, generator/train.py:123-132,151-153) ----
 * gtos_grad_sumsq: out[0] = sum g^2 (deterministic two-stage reduction; workspace = gtos_grad_sumsq_workspace() floats).
 * gtos_adam_step: g' = g * min(1, max_norm / (sqrt(*norm_sq) + 1e-6)) (norm_sq NULL: no clip);
 *   m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;  p -= lr * (m / (sqrt(v) + eps) + wd_i p), wd_i = wd for the first
 *   n_decay elements and 0 for the rest (no bias correction, as the reference).  lr is read from DEVICE memory so a
 *   captured CUDA graph can follow the schedule (train.py:81-83). */
int64_t gtos_grad_sumsq_workspace(void);
int gtos_grad_sumsq(const float* g, int64_t n, float* out, float* workspace, void* stream);
int gtos_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int64_t n_decay, const float* lr_ptr, float beta1,
                   float beta2, float eps, float weight_decay, const float* norm_sq, float max_norm, void* stream);

/* ---- batch construction (generator/AMRGraph.py:100-115 collect_concepts_and_relations,
 * generator/data.py:148-154 the per-pair path choice of batchify; translator/dependencyGraph.py:54-74) ----
 * For every graph b < B and every ordered pair (i, j) of its n_nodes[b] nodes: the edge labels of ONE shortest path i -> j,
 * drawn uniformly among all shortest node paths (the reference enumerates them with networkx and calls random.choice;
 * here they are counted by a BFS from j and sampled by a walk from i).  The graph is the padded adjacency
 * deg[b][v], nbr[b][v][k], lab[b][v][k] (label of the edge v -> nbr), k < deg <= deg_max, one entry per neighbour, and its
 * STRUCTURE must be symmetric (u in nbr[v] <=> v in nbr[u]) - the reference adds a `_reverse_` / `_r_` twin for every edge.
 * paths[b][i][j][0..plen) = label ids in walking order (0 beyond); the empty path gives {self_id}, a path of more than
 * max_len labels (or an unreachable pair) {tl_id}; pairs with i or j >= n_nodes[b] get plen = 0.
 * The draw of pair (b,i,j) at step s is hash(*seed_ptr + seed_off, ((b*n_max+i)*n_max+j)*max_len+s): reproducible, and
 * bit-identical to oracle/paths_oracle.py. */
int gtos_graph_paths(const int32_t* n_nodes, const int32_t* deg, const int32_t* nbr, const int32_t* lab, int32_t B,
                     int32_t n_max, int32_t deg_max, int32_t max_len, int32_t self_id, int32_t tl_id, const void* seed_ptr,
                     uint64_t seed_off, int32_t* paths, int32_t* plen, void* stream);

/* Evaluation batches keep EVERY shortest path of a pair (generator/data.py:176-225; generator.py:83-88 averages their
 * encodings).  Same graph format as gtos_graph_paths.  all_paths[b][i][j][k][0..len) = labels of the k-th shortest path
 * i -> j in depth-first adjacency order, k < min(pcount, K); pcount[b][i][j] = number of shortest paths saturated at K + 1
 * (a pair that reports K + 1 carries its first K paths), 1 for the <SELF> / <TL> pairs (data.py:197-199), 0 outside the
 * graph.  max_len <= 16. */
int gtos_graph_all_paths(const int32_t* n_nodes, const int32_t* deg, const int32_t* nbr, const int32_t* lab, int32_t B,
                         int32_t n_max, int32_t deg_max, int32_t max_len, int32_t K, int32_t self_id, int32_t tl_id,
                         int32_t* all_paths, int32_t* pcount, void* stream);

/* Node order of a batch: AMRGraph.bfs (generator/AMRGraph.py:82-98; translator/dependencyGraph.py:36-52) - the queue BFS
 * from root[b] that visits neighbours in adjacency order.  order[b][k] = k-th node of the queue (-1 beyond it),
 * depth[b][k] = its BFS depth (the `concept_depth` input, data.py:129), pos[b][v] = position of node v (-1 = not reached),
 * reached[b] = queue length (== n_nodes[b] iff connected).  One thread per graph replays the sequential queue exactly. */
int gtos_graph_bfs(const int32_t* n_nodes, const int32_t* deg, const int32_t* nbr, const int32_t* root, int32_t B, int32_t n_max,
                   int32_t deg_max, int32_t* order, int32_t* depth, int32_t* pos, int32_t* reached, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GTOS_B200_H_ */
