// gtos_b200 -- declarations of the memory-bound helper kernels (elementwise.cu, attention.cu, gru.cu)
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace gtos {

int cast_f32_bf16(const float* src, long lds, void* dst, long ldd, long rows, int cols, cudaStream_t st);
int cast_colsum(const float* src, long lds, void* dst, long ldd, float* sums, long rows, int cols, cudaStream_t st);
int weight_prep(const float* W, int R, int C, void* Wb, long ldw, void* Wt, long ldt, int perm_D, int perm_hd,
                cudaStream_t st);
int add_ln_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y, void* y_bf16,
               float* z, float* mean, float* rstd, long rows, int D, float p_drop, const void* seed_ptr,
               unsigned long long seed_off, cudaStream_t st);
int add_ln_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma, float* dres,
               float* dx, void* dx_bf16, float* dgamma, float* dbeta, long rows, int D, float p_drop,
               const void* seed_ptr, unsigned long long seed_off, cudaStream_t st);
int ln_param_grad(const float* dy, const float* z, const float* mean, const float* rstd, float* dgamma, float* dbeta,
                  long rows, int D, cudaStream_t st);
int colsum(const float* x, long ld, float* out, long rows, int cols, cudaStream_t st);
int colsum_bf16(const void* x, long ld, float* out, long rows, int cols, cudaStream_t st);
int dropout_bf16(void* h, long n, float p, const void* seed_ptr, unsigned long long seed_off, cudaStream_t st);
int dropout_f32(const float* x, float* out, long n, float p, const void* seed_ptr, unsigned long long seed_off,
                cudaStream_t st);
int relu_drop_bwd(const float* dh_in, const void* act, float* dh_f32, void* dh_bf16, long n, float p, cudaStream_t st);
int token_nll_fwd(const float* logits, long ldl, int V, const float* gate_logits, const float* align, int S,
                  const long long* copy_seq, const long long* target, long rows, int B, long long pad_idx,
                  float* loss_row, float* stats, cudaStream_t st);
int token_nll_bwd(const float* dloss_row, const float* logits, long ldl, int V, const float* align, int S,
                  const long long* copy_seq, const long long* target, long rows, int B, long long pad_idx,
                  const float* stats, float* dlogits, long lddl, float* dgate_logits, float* dalign, void* dlogits_bf16, long lddb,
                  cudaStream_t st);
int bank_gather(const float* bank, const long long* idx, long P, int D, float* out_f32, void* out_bf16, cudaStream_t st);
int bank_segsum_f32(const float* d_rel, const long long* order, const long long* keys, long P, int D, float* d_bank, long R,
                    cudaStream_t st);
int bank_gather_mean(const float* bank, const long long* idx, long P, int K, int D, float* out_f32, void* out_bf16,
                     cudaStream_t st);
int bank_scatter_add(const float* d_rel, const long long* idx, long P, int D, float* d_bank, long R, cudaStream_t st);
int rel_dqk(const void* G, const RelTiling& rt, float* dq, float* dk, long ld, void* dq_bf16, void* dk_bf16,
            cudaStream_t st);
int rel_pair_keys(const long long* idx, const RelTiling& rt, int R, int* keys, cudaStream_t st);
int rel_segsum(const void* G, const int* order, const int* keys, long n, int C, void* out_bf16, long ldo, float* spill,
               cudaStream_t st);

// ---- attention core (attention.cu) ------------------------------------------------------
struct AttnArgs {
  int T, S, B, H, hd;
  // projections, fp32, element (t, b, h, d) at ptr[(t*B + b)*ld + h*hd + d]
  const float* q; long ldq;     // unscaled; `scale` applied inside (null in encoder mode)
  const float* k; long ldk;
  const float* v; long ldv;
  float scale;
  // encoder mode: scores precomputed by the fused relation kernel, layout [B,H,S(j),T(i)]
  const float* scores_jt;
  const unsigned char* key_pad;   // [S,B] 1 = padding, or null
  const unsigned char* attn_mask; // [T,S] 1 = blocked, or null
  float p_drop;                   // dropout on the attention weights (weights_dropout=True)
  const void* seed_ptr; unsigned long long seed_off;
  float* probs;                   // [B,H,T,S] softmax (pre-dropout), saved for backward
  float* probs_dropped;           // optional [B,H,T,S] post-dropout weights (need_weights)
  float* out; long ldo;           // [T,B,H*hd] fp32
  void* out_bf16;                 // optional bf16 copy (same ld)
  int precise;                    // fp32 mode: split-bf16 operand planes, three MMA passes per product (attention.cu)
};
int attn_fwd(const AttnArgs& a, cudaStream_t st);

// several byte ranges zeroed by one launch (elementwise.cu)
constexpr int ZERO_MAX_REGIONS = 48;
struct ZeroRegions {
  void* ptr[ZERO_MAX_REGIONS];
  long bytes[ZERO_MAX_REGIONS];
};
int zero_regions(const ZeroRegions& z, int n, cudaStream_t st);

// ---- fp32 mode (precise.cu) ----
int split3(const float* src, long ld_r, long ld_c, long rows, int cols, void* dst, long ldd, int kp, int role,
           cudaStream_t st);
int rel_score_f32(const float* PR, long ldpr, const float* q, const float* k, long ldqk, float* scores, int N, int B,
                  int D, int H, cudaStream_t st);
int rel_grad_f32(const float* PR, long ldpr, const float* q, const float* k, long ldqk, const float* dscores, float* G,
                 long ldg, int N, int B, int D, int H, cudaStream_t st);
int rel_dqk_f32(const float* G, long ldg, float* dq, float* dk, long ld, int N, int B, int D, cudaStream_t st);
int relu_drop_bwd_f32(const float* dh_in, const float* act, float* dh_out, long n, float p, cudaStream_t st);
int gru_gate_fwd_f32(const float* gi, long ldgi, const float* gh, long ldgh, const float* h_prev, const long long* lengths,
                     int t, float* h_new, float* out_t, long ldout, float* gates, long R, int Hh, cudaStream_t st);
int gru_gate_bwd_f32(const float* dh, const float* dout_t, long lddout, const float* gates, const float* h_prev,
                     const long long* lengths, int t, float* dh_part, float* dgi, long lddgi, float* dgh, long lddgh,
                     long R, int Hh, cudaStream_t st);

struct AttnBwdArgs {
  AttnArgs f;                     // forward description (q,k,v,probs,masks,...)
  const float* dout; long lddo;   // [T,B,H*hd]
  const float* dprobs_extra;      // optional gradient wrt returned (post-dropout) weights [B,H,T,S]
  float* dscores_jt;              // encoder mode out: [B,H,S(j),T(i)]
  float* dscores_ts;              // scratch/out [B,H,T,S] (always written)
  float* dq; long lddq;           // decoder mode outs
  float* dk; long lddk;
  float* dv; long lddv;
  void* dq_bf16; void* dk_bf16; void* dv_bf16;   // optional bf16 copies, same element layout as dq / dk / dv
  int dk_in_q;                    // set by attn_bwd: the query-side kernel also writes dK (see attn_bwd_dk_on_query_side)
};
int attn_bwd_dk_on_query_side(const AttnArgs& a, bool decoder_mode);
int attn_bwd(const AttnBwdArgs& a, int part, cudaStream_t st);
int attn_debug_read_trace(unsigned long long* host_out, int enable);   // 3 kernels x 16 clock64 slots of CTA 0

// ---- GRU gate kernels (gru.cu) ------------------------------------------------------------
int gru_weight_prep(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int Kin, int H, int Kx,
                    void* Wcat, long ldw, float* bcat, cudaStream_t st);
int gru_gate_bwd(const float* dh, const float* dout_t, long lddout, const void* gates, const float* h_prev,
                 const long long* lengths, int t, float* dh_prev, void* dgi_bf16, long lddgi, void* dgh_bf16,
                 long lddgh, float* db_ih, float* db_hh, long R, int Hh, cudaStream_t st);
int embed_gather(const float* table, const long long* idx, long n, int dim, float* out_f32, void* out_bf16, long ldb,
                 float p_drop, const void* seed_ptr, unsigned long long seed_off, cudaStream_t st);
int embed_scatter_add(const float* dx, const long long* idx, long n, int dim, float* dtable, float p_drop,
                      const void* seed_ptr, unsigned long long seed_off, cudaStream_t st);


// ---- incremental beam decode, log-prob table, flat Adam (decode.cu) -----------------
int attn_decode(int Hyp, int L, int H, int hd, const float* q, long ldq, const void* kv, long ld_kv, int v_off,
                long row_stride, const int* slot, long slot_ld, const unsigned char* key_pad, long pad_ld, float scale,
                float* out, long ldo, void* out_bf16, long ldob, float* probs, cudaStream_t st);
int token_logprob(const float* logits, long ldl, int V, const float* gate_logits, const float* align, int S,
                  const long long* copy_seq, int Bsrc, const int* src_index, long rows, int B, float* table, long ldt, int W,
                  cudaStream_t st);
int token_topk(const float* logits, long ldl, int V, const float* gate_logits, const float* align, int S,
               const long long* copy_seq, int Bsrc, const int* src_index, long rows, int B, int W, int K, float* top_val,
               int* top_idx, float* table, long ldt, cudaStream_t st);
int beam_update_c(int B, int K, int t, int Tmin, int Tmax, int end_id, int unk_id, const float* top_val, const int* top_idx,
                  float* score, unsigned char* live, int* n_done, int* steps, int* tok, int* par, float* done_score,
                  int* done_step, int* done_par, int* parent_out, long long* last_tok, cudaStream_t st);
long grad_sumsq_workspace();
int grad_sumsq(const float* g, long n, float* out, float* workspace, cudaStream_t st);
int adam_step(float* p, const float* g, float* m, float* v, long n, long n_decay, const float* lr_ptr, float b1, float b2,
              float eps, float wd, const float* norm_sq, float max_norm, cudaStream_t st);
int beam_ancestry(const int* old_anc, int* new_anc, long ld, const int* parent, int t, int Hyp, cudaStream_t st);

// ---- shortest label paths of a graph batch (graph_paths.cu / graph_paths_core.h) ----
struct GraphPathsArgs;
int graph_paths(const GraphPathsArgs& a, const void* seed_ptr, unsigned long long seed_off, cudaStream_t st);
struct GraphAllPathsArgs;
int graph_all_paths(const GraphAllPathsArgs& a, cudaStream_t st);
struct GraphBfsArgs;
int graph_bfs(const GraphBfsArgs& a, cudaStream_t st);

}  // namespace gtos
