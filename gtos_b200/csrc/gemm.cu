// gtos_b200 -- warp-specialised tcgen05 GEMMs for sm_100a.
//
//   gemm_tn_kernel<BN, MODE>   C[128 x BN tile] = A[M,K] * B[N,K]^T, bf16 operands staged by TMA
//                              (128B swizzle, K-major), fp32 accumulators double-buffered in TMEM.
//        warp 0    : TMA producer (one elected lane)
//        warp 1    : TMEM allocator + tcgen05.mma issuer (one lane)
//        warps 2-5 : epilogue, each owns the TMEM lane quarter (warp_idx % 4)
//     MODE_SCORE / MODE_GRAD are the fused relation-attention kernels: A tiles are 4-D TMA boxes
//     of relation[j][i][b][:], B is the head-interleaved relation_in_proj weight, and the
//     epilogue combines the ra/rb accumulators with q_i / k_j slices that the producer also
//     stages through TMA (reference: generator/graph_transformer.py:122-133).
//
//   gemm_nn_kernel<BN, REL>    C[M,N] = sum_k A[k,m] B[k,n] with MN-major UMMA descriptors
//                              (weight gradients: contraction over rows), split-K partials.
#include <stdlib.h>

#include "gemm.cuh"

namespace gtos {

static constexpr int BM = 128;
static constexpr int BK = 64;  // bf16 elements per k-block = one 128-byte swizzle row
static constexpr int A_STAGE_BYTES = BM * BK * 2;
static constexpr int GEMM_THREADS = 192;
static constexpr int MAX_STAGES = 8;
static constexpr uint64_t WATCHDOG_CYCLES = 8000000000ull;  // ~4 s: turn a pipeline hang into a trap

struct PipeBars {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint64_t qfull[2];
  uint64_t qempty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > WATCHDOG_CYCLES) {
      printf("gtos_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// GTOS_DBG bit 1: per-CTA clock64 timestamps of the pipeline phases (gtos_debug_read_trace); timing experiment only
__device__ unsigned long long g_trace[148 * 16];
#define GTOS_TRACE(slot) do { if ((p.dbg & 2) && blockIdx.x < 148) g_trace[blockIdx.x * 16 + (slot)] = clock64(); } while (0)

struct TnDev {
  int M, N, K;
  int m_tiles, n_tiles, k_blocks, units;
  int stages;
  const float* bias;
  float* out_f32;
  long ldo;
  __nv_bfloat16* out_bf16;
  long ldob;
  int relu, accumulate;
  const float* addend;
  long ldadd;
  RelTiling rt;
  float* scores;
  const float* dscores;
  __nv_bfloat16* G;
  int bi8, bj8;        // q / k box rows rounded up to 8 (1024-byte swizzle atoms)
  int qk_stage_bytes;  // 4*(bi8+bj8)*128
  int tma_out;         // epilogue stages tiles in shared memory and writes them with TMA stores
  int dbg;             // GTOS_DBG bit 0: relation epilogues skip their body (pipeline-rate experiment, wrong results)
  // MODE_SCORE with the attention tail fused in (fuse != 0): masks + softmax + dropout + PV in the epilogue
  int fuse;
  const uint8_t* key_pad;
  float p_drop;
  const void* seed_ptr;
  unsigned long long seed_off;
  float* probs;
  float* probs_dropped;
  float* att;
  long ldatt;
  __nv_bfloat16* att_b;
  // MODE_GRU
  int kx_blocks;       // k-blocks that come from x_t (tmA); the rest come from h_prev (tmQ slot)
  int gru_H, gru_t;
  const float* gru_hprev;
  const long long* gru_len;
  float* gru_hnew;
  __nv_bfloat16* gru_hbnew; long gru_ldhbn;
  __nv_bfloat16* gru_out; long gru_ldout;
  __nv_bfloat16* gru_gates; long gru_ldg;
};

__device__ __forceinline__ void rel_tile_decode(const RelTiling& t, int tile, int& b, int& j0, int& i0) {
  int ib = tile % t.ni_blk;
  int r = tile / t.ni_blk;
  int jb = r % t.nj_blk;
  b = r / t.nj_blk;
  i0 = ib * t.bi;
  j0 = jb * t.bj;
}

// 16 bf16 (as fp32) from a 128B-swizzled [rows x 64 bf16] TMA box: row `row`, elements [e0, e0+16), e0 % 16 == 0
__device__ __forceinline__ void lds_sw128_bf16x16(const uint8_t* box, int row, int e0, float* out) {
  const uint8_t* rp = box + row * 128;
  const int c0 = e0 >> 3;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    uint4 v = *reinterpret_cast<const uint4*>(rp + (((c0 + t) ^ (row & 7)) << 4));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 f = __bfloat1622float2(h[u]);
      out[8 * t + 2 * u] = f.x;
      out[8 * t + 2 * u + 1] = f.y;
    }
  }
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2 (relation modes): a CTA PAIR (cluster of 2, cta_group::2) works on two
// relation tiles at once: each CTA stages its own 128-row A tile and HALF of the BN weight rows, the leader issues
// M=256 MMAs that read both CTAs' shared memory, and each CTA's TMEM receives the accumulator of its own 128 rows.
// Halves the weight traffic per FLOP and the shared-memory operand traffic per SM.
// relation and GRU modes run TWO epilogue warpgroups (warps 2-5 and 6-9; both map onto TMEM lane quarters warp & 3):
// each takes half of the unit's heads / hidden units - their epilogues (score / gradient / gate math) cost more
// issue slots per unit than the unit's MMAs take cycles, and one warp per SM sub-partition cannot hide that
template <int MODE>
constexpr int tn_epi_wgs() { return 2; }   // plain / d_relation too: the warpgroups alternate over the 32-column output chunks
template <int MODE>
constexpr int tn_threads() { return 64 + 128 * tn_epi_wgs<MODE>(); }
template <int MODE>
constexpr int tn_out_stage_bytes() { return (MODE == MODE_GRAD ? 4 : 2) * BM * 128; }  // [128 rows x 128 B] swizzled tiles

template <int BN, int MODE, int CG>
__global__ void __launch_bounds__(tn_threads<MODE>(), 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmO, const TnDev p) {
  constexpr bool REL = (MODE == MODE_SCORE || MODE == MODE_GRAD);
  static_assert(CG == 1 || REL || MODE == MODE_PLAIN || MODE == MODE_DREL,
                "CTA pairs are wired up for the relation modes, d_relation and the plain GEMM");
  constexpr int OUT_STAGE_BYTES = tn_out_stage_bytes<MODE>();  // two staging tiles per epilogue warpgroup
  constexpr int EPI_WGS = tn_epi_wgs<MODE>();
  constexpr int B_STAGE_BYTES = (BN / CG) * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t IDESC = make_idesc_bf16(BM * CG, BN, 0, 0);
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0);
  // pair scheduling: both CTAs of a pair walk the same unit list; CTA `rank` owns relation tile 2*pair_tile + rank
  const int sched_id = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int sched_n = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* qk_base = smem + p.stages * STAGE_BYTES;
  uint8_t* out_stage = qk_base + (REL ? 2 * p.qk_stage_bytes : 0);  // 1024-aligned (all regions are multiples of 1 KB)
  [[maybe_unused]] uint8_t* gru_stage = out_stage + 2 * BN * 4;   // MODE_GRU: 8 warp-private 8 KB store-staging tiles
  PipeBars* bars = reinterpret_cast<PipeBars*>(out_stage + ((p.tma_out || p.fuse) ? OUT_STAGE_BYTES : (MODE == MODE_GRU ? 2 * BN * 4 + 8 * 8192 : 0)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) GTOS_TRACE(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out || p.fuse) tma_prefetch_desc(&tmO);
    if (REL) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);       // pair: only the leader arrives (expect_tx covers both CTAs' bytes)
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->tfull[s], 1);
      mbar_init(&bars->tempty[s], 4 * CG * EPI_WGS);  // pair: epilogue warps of BOTH CTAs release the leader's accumulator stage
      mbar_init(&bars->qfull[s], 1);
      mbar_init(&bars->qempty[s], 4 * EPI_WGS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2cta(&bars->tmem_base, TMEM_COLS); else tmem_alloc(&bars->tmem_base, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();        // peer barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_launch_dependents();  // our successor may begin its own prologue as SMs free up
  if (threadIdx.x == 0) GTOS_TRACE(1);
  pdl_wait();               // everything below touches global memory written by predecessors
  if (threadIdx.x == 0) GTOS_TRACE(2);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int qs = 0;
      uint32_t qph = 0;
      for (int unit = sched_id; unit < p.units; unit += sched_n) {
        const int m_blk = (unit / p.n_tiles) * CG + (int)rank, n_blk = unit % p.n_tiles;
        int b = 0, j0 = 0, i0 = 0;
        if (REL) {
          rel_tile_decode(p.rt, m_blk, b, j0, i0);   // a dummy tile past the end decodes to b == B: TMA zero-fills it
          // q / k slices for this (tile, head group): dims [n_blk*BN/2, +BN/2)
          wait_bar(&bars->qempty[qs], qph ^ 1);
          const bool fuse_v = (MODE == MODE_SCORE) && p.fuse;
          mbar_expect_tx(&bars->qfull[qs], (uint32_t)((BN / 128) * (p.rt.bi + p.rt.bj * (fuse_v ? 2 : 1)) * 128));
          uint8_t* qb = qk_base + qs * p.qk_stage_bytes;
          const int d0 = n_blk * (BN / 2);
#pragma unroll
          for (int c = 0; c < BN / 128; ++c) {   // bf16 q/k: one 128-byte box row = 64 dims
            tma_load_3d(&tmQ, &bars->qfull[qs], qb + c * p.bi8 * 128, d0 + c * 64, b, i0);
            tma_load_3d(&tmK, &bars->qfull[qs], qb + (BN / 128) * p.bi8 * 128 + c * p.bj8 * 128, d0 + c * 64, b, j0);
            if (fuse_v)                          // the value rows of the tile's keys, same box shape as k (tmO = v map)
              tma_load_3d(&tmO, &bars->qfull[qs], qb + (BN / 128) * (p.bi8 + p.bj8) * 128 + c * p.bj8 * 128, d0 + c * 64, b, j0);
          }
          if (++qs == 2) { qs = 0; qph ^= 1; }
        }
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          wait_bar(&bars->empty[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (REL && CG == 2) {
            // the peer's loads complete_tx on the leader's barrier too; it cannot run ahead of the leader's phase because
            // it only reuses a stage after the leader's MMA committed it (multicast) - so a cluster-scope arrive
            // (a membar per k-block) is not needed
            if (leader) mbar_expect_tx(&bars->full[s], (uint32_t)(2 * (p.rt.bi * p.rt.bj * 128 + B_STAGE_BYTES)));
            tma_load_4d_2cta(&tmA, &bars->full[s], sa, kb * BK, b, i0, j0);
            tma_load_2d_2cta(&tmB, &bars->full[s], sb, kb * BK, n_blk * BN + (int)rank * (BN / 2));
            if (++s == p.stages) { s = 0; ph ^= 1; }
            continue;
          } else if (CG == 2) {
            // plain GEMM as a CTA pair: own 128 rows of A, half of the BN weight rows; a row block past M (odd number of
            // row blocks) is zero-filled by TMA and still counts its full box
            if (leader) mbar_expect_tx(&bars->full[s], (uint32_t)(2 * STAGE_BYTES));
            tma_load_2d_2cta(&tmA, &bars->full[s], sa, kb * BK, m_blk * BM);
            tma_load_2d_2cta(&tmB, &bars->full[s], sb, kb * BK, n_blk * BN + (int)rank * (BN / 2));
            if (++s == p.stages) { s = 0; ph ^= 1; }
            continue;
          } else if (REL) {
            mbar_expect_tx(&bars->full[s], (uint32_t)(p.rt.bi * p.rt.bj * 128 + B_STAGE_BYTES));
            tma_load_4d(&tmA, &bars->full[s], sa, kb * BK, b, i0, j0);
          } else if (MODE == MODE_GRU) {
            mbar_expect_tx(&bars->full[s], (uint32_t)STAGE_BYTES);
            if (kb < p.kx_blocks)
              tma_load_2d(&tmA, &bars->full[s], sa, kb * BK, m_blk * BM);
            else
              tma_load_2d(&tmQ, &bars->full[s], sa, (kb - p.kx_blocks) * BK, m_blk * BM);
          } else {
            mbar_expect_tx(&bars->full[s], (uint32_t)STAGE_BYTES);
            tma_load_2d(&tmA, &bars->full[s], sa, kb * BK, m_blk * BM);
          }
          tma_load_2d(&tmB, &bars->full[s], sb, kb * BK, n_blk * BN);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
      GTOS_TRACE(3);
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && leader) {
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int unit = sched_id; unit < p.units; unit += sched_n) {
        wait_bar(&bars->tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          wait_bar(&bars->full[s], ph);
          tc_fence_after();
          if (kb == 0 && unit == sched_id) GTOS_TRACE(4);
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sa + A_STAGE_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128B swizzle row: +2 in the (addr>>4) field
            if (CG == 2)
              umma_bf16_2cta(tacc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) != 0 ? 1u : 0u);
            else
              umma_bf16(tacc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) != 0 ? 1u : 0u);
          }
          if (CG == 2) umma_commit_2cta(&bars->empty[s]); else umma_commit(&bars->empty[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (CG == 2) umma_commit_2cta(&bars->tfull[as]); else umma_commit(&bars->tfull[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
      GTOS_TRACE(5);
    }
  } else {
    // ================= epilogue (warps 2..5, relation modes: also 6..9) =================
    const int quarter = warp & 3;
    const int wg = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;  // accumulator row == TMEM lane
    int as = 0;
    uint32_t aph = 0;
    int qs = 0;
    uint32_t qph = 0;
    for (int unit = sched_id; unit < p.units; unit += sched_n) {
      const int m_blk = (unit / p.n_tiles) * CG + (int)rank, n_blk = unit % p.n_tiles;
      // MODE_GRU: fetch this thread's slice of the previous state, the bias block and the length flag while the
      // tensor core is still producing the accumulator (4 epilogue warps cannot hide DRAM latency otherwise)
      [[maybe_unused]] float gru_hp[MODE == MODE_GRU ? (BN >= 128 ? BN / 8 : BN / 4) : 1];   // this warpgroup's half of the unit's hidden units
      [[maybe_unused]] long long gru_len_v = 0;   // compared with the time step only after the accumulator wait
      [[maybe_unused]] const float* gru_bias = nullptr;
      if constexpr (MODE == MODE_GRU) {
        constexpr int UB = BN / 4;
        const long row = (long)m_blk * BM + r;
        float* sb = reinterpret_cast<float*>(out_stage) + (as * BN);     // bias block of this unit, per accumulator stage
        for (int t = r + 128 * wg; t < BN; t += 128 * EPI_WGS) sb[t] = p.bias[n_blk * BN + t];
        gru_bias = sb;
        constexpr int UW = (UB >= 32) ? UB / 2 : UB;   // hidden units per warpgroup (a 16-unit tile is not split)
        if (row < p.M && (UB >= 32 || wg == 0)) {
          gru_len_v = p.gru_len[row];
          const int uw = n_blk * UB + wg * UW;         // first hidden unit of this warpgroup
          const float4* hpp = reinterpret_cast<const float4*>(p.gru_hprev + row * p.gru_H + uw);
#pragma unroll
          for (int t4 = 0; t4 < UW / 4; ++t4) {
            float4 v = (uw + 4 * t4 < p.gru_H) ? hpp[t4] : make_float4(0.f, 0.f, 0.f, 0.f);
            gru_hp[4 * t4] = v.x; gru_hp[4 * t4 + 1] = v.y; gru_hp[4 * t4 + 2] = v.z; gru_hp[4 * t4 + 3] = v.w;
          }
        }
        named_bar_sync(1, 128 * EPI_WGS);  // bias block visible to all epilogue warps (also keeps the warpgroups within
                                           // one unit of each other, so the per-stage bias block is never overwritten early)
      }
      // relation modes: decode the tile and (MODE_GRAD) fetch this thread's d(score) values of the warpgroup's heads
      // while the tensor core is still producing the accumulator
      [[maybe_unused]] int rb_ = 0, rj0 = 0, ri0 = 0, rjj = 0, rii = 0, rch0 = 0, rnch = 0;
      [[maybe_unused]] bool rvalid = false;
      [[maybe_unused]] long rs0 = 0;          // scores index of (b, head 0 of this unit, j, i)
      [[maybe_unused]] float gpre[4] = {0.f, 0.f, 0.f, 0.f};
      if constexpr (REL) {
        constexpr int NCH = (BN / 2) / 16;    // 16-dim chunks per unit
        rel_tile_decode(p.rt, m_blk, rb_, rj0, ri0);
        const int bi = p.rt.bi;
        rjj = r / bi; rii = r - rjj * bi;
        const int i = ri0 + rii, j = rj0 + rjj;
        rvalid = (rjj < p.rt.bj) && (i < p.rt.N) && (j < p.rt.N) && (m_blk < p.m_tiles);
        const int hd = p.rt.hd;
        const int heads_blk = (BN / 2) / hd;
        rch0 = heads_blk >= 2 ? wg * (NCH / 2) : 0;
        rnch = (p.dbg & 1) ? 0 : (heads_blk >= 2 ? NCH / 2 : (wg == 0 ? NCH : 0));
        rs0 = (((long)rb_ * p.rt.H + n_blk * heads_blk) * p.rt.N + j) * p.rt.N + i;
        if constexpr (MODE == MODE_GRAD) {
          const int h0 = (rch0 * 16) / hd;                  // first head of this warpgroup inside the unit
          const int nh = (rnch * 16 + hd - 1) / hd;         // heads it owns (<= 4)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (rvalid && q < nh) gpre[q] = p.dscores[rs0 + (long)(h0 + q) * p.rt.N * p.rt.N] * p.rt.scale;
        }
      }
      wait_bar(&bars->tfull[as], aph);
      if (REL) wait_bar(&bars->qfull[qs], qph);
      tc_fence_after();
      if (warp == 2 && lane == 0 && unit == sched_id) GTOS_TRACE(6);
      const uint32_t tacc = tmem_base + as * BN + ((uint32_t)(quarter * 32) << 16);

      if constexpr (MODE == MODE_GRU) {
        // ---- GRU gate math: this unit owns hidden units [n_blk*UB, +UB) of rows [m_blk*128, +128) ----
        // (the previous-state loads for the unit are issued BEFORE waiting on the accumulator, see gru_prefetch)
        constexpr int UB = BN / 4;
        const long row = (long)m_blk * BM + r;
        const bool row_ok = row < p.M;
        const int H = p.gru_H;
        const bool live = row_ok && gru_len_v > p.gru_t;
        constexpr int UW = (UB >= 32) ? UB / 2 : UB;
        const int c_end = (UB >= 32 || wg == 0) ? wg * UW + UW : 0;
#pragma unroll
        for (int c = wg * UW; c < c_end; c += 16) {
          float ar[16], az[16], ai[16], ah[16];
          tmem_ld16(tacc + c, ar);
          tmem_ld16(tacc + UB + c, az);
          tmem_ld16(tacc + 2 * UB + c, ai);
          tmem_ld16(tacc + 3 * UB + c, ah);
          tmem_ld_wait();
          const int u0 = n_blk * UB + c;  // first hidden unit of this chunk
          if (u0 < H) {                   // warp-uniform
            // Results go through a warp-private swizzled smem tile so that every global store instruction writes whole
            // 32/64-byte row pieces of 8-16 rows (full sectors) instead of 16 bytes of 32 different rows: the epilogue
            // was bound by the number of L2 store requests, not by bytes (ncu: long-scoreboard stalls behind STG.128)
            uint8_t* wst = gru_stage + ((warp - 2) * 8192);
            if (row_ok) {
              const float* hp = gru_hp + (c - wg * UW);
              float hn[16], gr[16], gz[16], gn[16], hh[16];
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                const float r_ = __fdividef(1.f, 1.f + __expf(-(ar[t] + gru_bias[c + t])));
                const float z_ = __fdividef(1.f, 1.f + __expf(-(az[t] + gru_bias[UB + c + t])));
                const float hn_ = ah[t] + gru_bias[3 * UB + c + t];
                const float pre = ai[t] + gru_bias[2 * UB + c + t] + r_ * hn_;
                const float n_ = 1.f - __fdividef(2.f, 1.f + __expf(2.f * pre));   // tanh
                gr[t] = live ? r_ : 0.f; gz[t] = live ? z_ : 0.f; gn[t] = live ? n_ : 0.f; hh[t] = live ? hn_ : 0.f;
                hn[t] = live ? (1.f - z_) * n_ + z_ * hp[t] : hp[t];
              }
              const int sw4 = (lane >> 1) & 3, sw2 = (lane >> 2) & 1;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<float4*>(wst + lane * 64 + ((k ^ sw4) << 4)) =
                    make_float4(hn[4 * k], hn[4 * k + 1], hn[4 * k + 2], hn[4 * k + 3]);
#define GTOS_STAGE16(off, v, zero)                                                                                     \
  do {                                                                                                                \
    uint8_t* _b = wst + (off) + lane * 32;                                                                            \
    *reinterpret_cast<uint4*>(_b + ((0 ^ sw2) << 4)) = (zero) ? make_uint4(0, 0, 0, 0) :                              \
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])); \
    *reinterpret_cast<uint4*>(_b + ((1 ^ sw2) << 4)) = (zero) ? make_uint4(0, 0, 0, 0) :                              \
        make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15])); \
  } while (0)
              GTOS_STAGE16(2048, hn, false);
              if (p.gru_out) GTOS_STAGE16(3072, hn, !live);
              GTOS_STAGE16(4096, gr, false);
              GTOS_STAGE16(5120, gz, false);
              GTOS_STAGE16(6144, gn, false);
              GTOS_STAGE16(7168, hh, false);
#undef GTOS_STAGE16
            }
            __syncwarp();
            const long row0 = (long)m_blk * BM + quarter * 32;     // first row of this warp
#pragma unroll
            for (int i = 0; i < 4; ++i) {                          // h (fp32): 8 rows x 64 B per instruction
              const int rr = i * 8 + (lane >> 2), k = lane & 3;
              const float4 v = *reinterpret_cast<const float4*>(wst + rr * 64 + ((k ^ ((rr >> 1) & 3)) << 4));
              if (row0 + rr < p.M) *reinterpret_cast<float4*>(p.gru_hnew + (row0 + rr) * H + u0 + k * 4) = v;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {                          // bf16 pieces: 16 rows x 32 B per instruction
              const int rr = i * 16 + (lane >> 1), k = lane & 1;
              const int so = rr * 32 + ((k ^ ((rr >> 2) & 1)) << 4);
              if (row0 + rr < p.M) {
                const long grow = row0 + rr;
                *reinterpret_cast<uint4*>(p.gru_hbnew + grow * p.gru_ldhbn + u0 + k * 8) =
                    *reinterpret_cast<const uint4*>(wst + 2048 + so);
                if (p.gru_out)
                  *reinterpret_cast<uint4*>(p.gru_out + grow * p.gru_ldout + u0 + k * 8) =
                      *reinterpret_cast<const uint4*>(wst + 3072 + so);
                __nv_bfloat16* gp = p.gru_gates + grow * p.gru_ldg + (long)n_blk * BN + c + k * 8;
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4)
                  *reinterpret_cast<uint4*>(gp + g4 * UB) = *reinterpret_cast<const uint4*>(wst + 4096 + g4 * 1024 + so);
              }
            }
            __syncwarp();
          }
        }
      } else if constexpr (MODE == MODE_PLAIN || MODE == MODE_DREL) {
        if (p.tma_out) {
          // fp32 tile -> swizzled smem staging tile [128 rows x 32 cols] -> one TMA store per 32-column chunk
          // (double-buffered; OOB rows / columns are clipped by the tensor map)
          // two epilogue warpgroups: wg owns staging buffer wg and every second output chunk (fp32: 32 columns, bf16: 64)
          const bool issuer = (quarter == 2 && lane == 0);       // first warp of each warpgroup
          int b0 = 0, j0 = 0, i0 = 0;
          if constexpr (MODE == MODE_DREL) rel_tile_decode(p.rt, m_blk, b0, j0, i0);
          const int n0 = n_blk * BN;
          const int own_shift = (MODE == MODE_PLAIN && p.tma_out == 2) ? 6 : 5;
          uint8_t* const wbuf = out_stage + wg * (BM * 128);
#pragma unroll 1
          for (int c = 0; c < BN; c += 32) {
            if (n0 + c >= p.N) break;
            if (((c >> own_shift) & 1) != wg) continue;
            float v[32];
            const bool trc = (warp == 2 && lane == 0 && unit == sched_id && c == 0);
            if (trc) GTOS_TRACE(10);
            tmem_ld16(tacc + c, v);
            tmem_ld16(tacc + c + 16, v + 16);
            // bias of this chunk: 8 independent 16-byte loads issued BEFORE the TMEM wait (the scalar predicated form
            // compiled to 32 loads into ONE register, each waiting out its own L1 latency: 1800 cycles per chunk)
            float bv[32];
            if (p.bias) {
              const float* bp = p.bias + n0 + c;
              if (n0 + c + 32 <= p.N && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
                for (int t4 = 0; t4 < 8; ++t4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + t4);
                  bv[4 * t4] = b4.x; bv[4 * t4 + 1] = b4.y; bv[4 * t4 + 2] = b4.z; bv[4 * t4 + 3] = b4.w;
                }
              } else {
#pragma unroll
                for (int t = 0; t < 32; ++t) bv[t] = (n0 + c + t < p.N) ? __ldg(bp + t) : 0.f;
              }
            }
            tmem_ld_wait();
            if (trc) GTOS_TRACE(11);
            if (p.bias) {
#pragma unroll
              for (int t = 0; t < 32; ++t) v[t] += bv[t];
            }
            if (p.addend) {
              const long arow = (long)m_blk * BM + r;
              if (arow < p.M) {
                const float* ap = p.addend + arow * p.ldadd + n0 + c;
                if (n0 + c + 32 <= p.N && ((reinterpret_cast<uintptr_t>(ap) & 15) == 0)) {
#pragma unroll
                  for (int t4 = 0; t4 < 8; ++t4) {
                    const float4 a4 = *reinterpret_cast<const float4*>(ap + 4 * t4);
                    v[4 * t4] += a4.x; v[4 * t4 + 1] += a4.y; v[4 * t4 + 2] += a4.z; v[4 * t4 + 3] += a4.w;
                  }
                } else {
#pragma unroll
                  for (int t = 0; t < 32; ++t)
                    if (n0 + c + t < p.N) v[t] += ap[t];
                }
              }
            }
            if (p.relu) {
#pragma unroll
              for (int t = 0; t < 32; ++t) v[t] = fmaxf(v[t], 0.f);
            }
            if constexpr (MODE == MODE_PLAIN) {
              if (p.tma_out == 2) {
                // bf16 output: two 32-column chunks fill one [128 rows x 64 bf16] swizzled tile, one TMA store per tile
                const int half = (c >> 5) & 1;
                uint8_t* buf = wbuf;
                if (half == 0) {
                  if (issuer) tma_store_wait_read<0>();
                  named_bar_sync(1 + 2 * wg, 128);
                }
                uint8_t* rowp = buf + r * 128;
#pragma unroll
                for (int t = 0; t < 4; ++t)
                  *reinterpret_cast<uint4*>(rowp + (((half * 4 + t) ^ (r & 7)) << 4)) =
                      make_uint4(pack_bf16x2(v[8 * t], v[8 * t + 1]), pack_bf16x2(v[8 * t + 2], v[8 * t + 3]),
                                 pack_bf16x2(v[8 * t + 4], v[8 * t + 5]), pack_bf16x2(v[8 * t + 6], v[8 * t + 7]));
                if (half == 1 || c + 32 >= BN || n0 + c + 32 >= p.N) {      // tile complete (or last chunk of the unit)
                  fence_proxy_async();
                  named_bar_sync(2 + 2 * wg, 128);
                  if (issuer) {
                    tma_store_2d(&tmO, buf, n0 + (c & ~63), m_blk * BM);
                    tma_store_commit();
                  }
                }
                continue;
              }
            }
            uint8_t* buf = wbuf;
            if (trc) GTOS_TRACE(12);
            if (issuer) tma_store_wait_read<0>();   // this warpgroup's previous store has finished reading the buffer
            if (trc) GTOS_TRACE(13);
            named_bar_sync(1 + 2 * wg, 128);
            if (trc) GTOS_TRACE(14);
            uint8_t* rowp = buf + r * 128;
#pragma unroll
            for (int t = 0; t < 8; ++t)
              *reinterpret_cast<float4*>(rowp + ((t ^ (r & 7)) << 4)) =
                  make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
            fence_proxy_async();
            named_bar_sync(2 + 2 * wg, 128);
            if (issuer) {
              if constexpr (MODE == MODE_DREL) {
                if (p.accumulate)
                  tma_reduce_add_4d(&tmO, buf, n0 + c, b0, i0, j0);
                else
                  tma_store_4d(&tmO, buf, n0 + c, b0, i0, j0);
              } else
                tma_store_2d(&tmO, buf, n0 + c, m_blk * BM);
              tma_store_commit();
            }
            if (trc) GTOS_TRACE(15);
          }
        } else {
        long out_row = (long)m_blk * BM + r;
        bool row_ok = out_row < p.M;
        if constexpr (MODE == MODE_DREL) {
          int b, j0, i0;
          rel_tile_decode(p.rt, m_blk, b, j0, i0);
          int jj = r / p.rt.bi, ii = r - jj * p.rt.bi;
          int i = i0 + ii, j = j0 + jj;
          row_ok = (jj < p.rt.bj) && (i < p.rt.N) && (j < p.rt.N);
          out_row = ((long)j * p.rt.N + i) * p.rt.B + b;
        }
        const int n0 = n_blk * BN;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
          if (n0 + c >= p.N) break;
          if (((c >> 4) & 1) != wg) continue;     // the two warpgroups alternate over the 16-column chunks
          float v[16];
          tmem_ld16(tacc + c, v);
          const int n = n0 + c;
          float bv[16];
          if (p.bias) {                         // independent loads, in flight during the TMEM wait (see the TMA path)
            const float* bp = p.bias + n;
            if (n + 16 <= p.N && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
#pragma unroll
              for (int t4 = 0; t4 < 4; ++t4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + t4);
                bv[4 * t4] = b4.x; bv[4 * t4 + 1] = b4.y; bv[4 * t4 + 2] = b4.z; bv[4 * t4 + 3] = b4.w;
              }
            } else {
#pragma unroll
              for (int t = 0; t < 16; ++t) bv[t] = (n + t < p.N) ? __ldg(bp + t) : 0.f;
            }
          }
          tmem_ld_wait();
          if (row_ok) {
            if (p.bias) {
#pragma unroll
              for (int t = 0; t < 16; ++t) v[t] += bv[t];
            }
            if (p.relu) {
#pragma unroll
              for (int t = 0; t < 16; ++t) v[t] = fmaxf(v[t], 0.f);
            }
            const bool full = (n + 16 <= p.N);
            if (p.out_f32) {
              float* o = p.out_f32 + out_row * p.ldo + n;
              if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  float4 w = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
                  if (p.accumulate) {
                    float4 old = *reinterpret_cast<float4*>(o + 4 * t);
                    w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
                  }
                  *reinterpret_cast<float4*>(o + 4 * t) = w;
                }
              } else {
#pragma unroll
                for (int t = 0; t < 16; ++t)
                  if (n + t < p.N) o[t] = p.accumulate ? o[t] + v[t] : v[t];
              }
            }
            if (p.out_bf16) {
              __nv_bfloat16* o = p.out_bf16 + out_row * p.ldob + n;
              if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                uint4 w0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                      pack_bf16x2(v[6], v[7]));
                uint4 w1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]),
                                      pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                *reinterpret_cast<uint4*>(o) = w0;
                *reinterpret_cast<uint4*>(o + 8) = w1;
              } else {
#pragma unroll
                for (int t = 0; t < 16; ++t)
                  if (n + t < p.N) o[t] = __float2bfloat16(v[t]);
              }
            }
          }
        }
        }  // !tma_out
      } else {
        // ---- relation epilogues: thread owns pair (j0+jj, i0+ii, b); warpgroup wg owns chunks [rch0, rch0+rnch) ----
        constexpr int NCH = (BN / 2) / 16;
        const int jjc = rjj < p.rt.bj ? rjj : p.rt.bj - 1;  // keep smem reads inside the k boxes
        const uint8_t* qb = qk_base + qs * p.qk_stage_bytes;
        const uint8_t* kbx = qb + (BN / 128) * p.bi8 * 128;
        const int hd = p.rt.hd;
        const int h0 = (rch0 * 16) / hd;
        const long hstride = (long)p.rt.N * p.rt.N;
        const bool valid = rvalid;
        [[maybe_unused]] uint8_t* stage_wg = out_stage + wg * (2 * BM * 128);
        [[maybe_unused]] const bool issuer = (quarter == 2 && lane == 0);   // first warp of each warpgroup
        float acc = 0.f;
        float g = 0.f;
        [[maybe_unused]] float s_val = 0.f;     // fused attention: the finished score of this warpgroup's head
        float ra[2][16], rb[2][16];
        // software pipeline: the TMEM loads of chunk t+1 are in flight while chunk t is being computed
        auto issue_ld = [&](int it, float* a_, float* b_) {
          const int dl = it * 16;
          const int hh = dl / hd, c = dl - hh * hd;
          tmem_ld16(tacc + hh * 2 * hd + c, a_);
          tmem_ld16(tacc + hh * 2 * hd + hd + c, b_);
        };
        if (rnch > 0) issue_ld(rch0, ra[0], rb[0]);
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          if (t < rnch) {
            const int it = rch0 + t;
            const int dl = it * 16;                  // dim inside this unit's BN/2-wide slice
            const int hh = dl / hd, c = dl - hh * hd;
            float qv[16], kv[16];
            lds_sw128_bf16x16(qb + (dl >> 6) * p.bi8 * 128, rii, dl & 63, qv);
            lds_sw128_bf16x16(kbx + (dl >> 6) * p.bj8 * 128, jjc, dl & 63, kv);
            tmem_ld_wait();
            if (t + 1 < rnch) issue_ld(it + 1, ra[(t + 1) & 1], rb[(t + 1) & 1]);
            const float* ra_ = ra[t & 1];
            const float* rb__ = rb[t & 1];
            if (c == 0) {
              acc = 0.f;
              if constexpr (MODE == MODE_GRAD) {
                const int hq = hh - h0;
                g = hq == 0 ? gpre[0] : hq == 1 ? gpre[1] : hq == 2 ? gpre[2] : gpre[3];
              }
            }
            if constexpr (MODE == MODE_SCORE) {
#pragma unroll
              for (int u = 0; u < 16; ++u) acc = fmaf(qv[u] + ra_[u], kv[u] + rb__[u], acc);
              if (c + 16 == hd) {
                if (p.fuse) s_val = acc * p.rt.scale;
                else if (valid) p.scores[rs0 + (long)hh * hstride] = acc * p.rt.scale;
              }
            } else {
              // G columns (permuted order): [d(q+ra) = g*(k+rb) | d(k+rb) = g*(q+ra)]
              uint32_t wx[8], wy[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                float x0 = qv[2 * u] + ra_[2 * u], x1 = qv[2 * u + 1] + ra_[2 * u + 1];
                float y0 = kv[2 * u] + rb__[2 * u], y1 = kv[2 * u + 1] + rb__[2 * u + 1];
                wx[u] = valid ? pack_bf16x2(g * y0, g * y1) : 0u;
                wy[u] = valid ? pack_bf16x2(g * x0, g * x1) : 0u;
              }
              const int gcx = hh * 2 * hd + c, gcy = gcx + hd;   // G columns inside this 256-wide block
              if (p.tma_out) {
                // stage 64-column (128-byte) spans of G in this warpgroup's two swizzled smem tiles, one TMA store
                // per finished span
                const bool wide = hd >= 64;                      // X and Y halves live in different spans
                uint8_t* bx = stage_wg + (wide ? 0 : ((gcx >> 6) & 1)) * (BM * 128);
                uint8_t* by = stage_wg + (wide ? 1 : ((gcy >> 6) & 1)) * (BM * 128);
                if ((gcx & 63) == 0) {                           // first chunk of a span: buffer is reused
                  if (issuer) {
                    if (wide) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
                  }
                  named_bar_sync(1 + 2 * wg, 128);
                }
                uint8_t* rx = bx + r * 128;
                uint8_t* ry = by + r * 128;
                const int cx = (gcx & 63) >> 3, cy = (gcy & 63) >> 3;
                *reinterpret_cast<uint4*>(rx + (((cx) ^ (r & 7)) << 4)) = make_uint4(wx[0], wx[1], wx[2], wx[3]);
                *reinterpret_cast<uint4*>(rx + (((cx + 1) ^ (r & 7)) << 4)) = make_uint4(wx[4], wx[5], wx[6], wx[7]);
                *reinterpret_cast<uint4*>(ry + (((cy) ^ (r & 7)) << 4)) = make_uint4(wy[0], wy[1], wy[2], wy[3]);
                *reinterpret_cast<uint4*>(ry + (((cy + 1) ^ (r & 7)) << 4)) = make_uint4(wy[4], wy[5], wy[6], wy[7]);
                if (((gcy + 16) & 63) == 0) {                    // span(s) complete
                  fence_proxy_async();
                  named_bar_sync(2 + 2 * wg, 128);
                  if (issuer) {
                    if (wide) tma_store_2d(&tmO, bx, n_blk * BN + (gcx & ~63), m_blk * BM);
                    tma_store_2d(&tmO, by, n_blk * BN + (gcy & ~63), m_blk * BM);
                    tma_store_commit();
                  }
                }
              } else {
                __nv_bfloat16* grow = p.G + ((long)m_blk * BM + r) * (2 * p.rt.D) + (long)n_blk * BN + gcx;
                *reinterpret_cast<uint4*>(grow) = make_uint4(wx[0], wx[1], wx[2], wx[3]);
                *reinterpret_cast<uint4*>(grow + 8) = make_uint4(wx[4], wx[5], wx[6], wx[7]);
                *reinterpret_cast<uint4*>(grow + hd) = make_uint4(wy[0], wy[1], wy[2], wy[3]);
                *reinterpret_cast<uint4*>(grow + hd + 8) = make_uint4(wy[4], wy[5], wy[6], wy[7]);
              }
            }
          }
        }
        if constexpr (MODE == MODE_SCORE) {
          if (p.fuse) {
            // ---- attention tail (graph_transformer.py:136-159) on the tile's complete softmax rows: the tile holds ALL
            // keys of its bi queries (bj == N) and this warpgroup owns ONE head (hd = 64, BN = 256), so mask, softmax,
            // dropout and P.V finish here - the [B,H,N,N] scores never reach HBM and no second kernel runs.
            float* ssm = reinterpret_cast<float*>(out_stage) + wg * 256;      // [128 scores | 128 dropped probabilities]
            const int bi = p.rt.bi, bj = p.rt.bj, Nn = p.rt.N, Bb = p.rt.B;
            const int head = n_blk * 2 + wg;
            const int qi_ = ri0 + rii, kj_ = rj0 + rjj;
            const bool live = valid && !(p.key_pad && p.key_pad[(long)kj_ * Bb + rb_]);
            ssm[r] = live ? s_val : -INFINITY;
            // the accumulator is consumed: hand the TMEM stage back so the next unit's MMAs run under the softmax / P.V tail
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2 && !leader) mbar_arrive_remote(&bars->tempty[as], 0); else mbar_arrive(&bars->tempty[as]);
            }
            named_bar_sync(1 + 2 * wg, 128);
            // row statistics: every warp computes them for all bi (<= 4) queries, 8 lanes per query striding the keys,
            // and a thread picks the pair of its own query with one shuffle (no second barrier)
            float mx = -INFINITY, sum = 0.f;
            {
              const int sq = lane >> 3, su = lane & 7;
              const int sqc = sq < bi ? sq : bi - 1;
              for (int jj = su; jj < bj; jj += 8) mx = fmaxf(mx, ssm[jj * bi + sqc]);
#pragma unroll
              for (int o = 1; o < 8; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
              for (int jj = su; jj < bj; jj += 8) {
                const float sv = ssm[jj * bi + sqc];
                sum += (sv == -INFINITY) ? 0.f : __expf(sv - mx);
              }
#pragma unroll
              for (int o = 1; o < 8; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
              const int src = (rii < bi ? rii : 0) * 8;
              mx = __shfl_sync(0xffffffffu, mx, src);
              sum = __shfl_sync(0xffffffffu, sum, src);
            }
            float pr = (live && sum > 0.f) ? __expf(s_val - mx) / sum : 0.f;
            if (valid) {
              const long pidx = (((long)rb_ * p.rt.H + head) * Nn + qi_) * Nn + kj_;
              p.probs[pidx] = pr;
              if (p.p_drop > 0.f) {
                const unsigned long long seed = reinterpret_cast<const unsigned long long*>(p.seed_ptr)[0] + p.seed_off;
                pr = (rng_uniform(seed, (unsigned long long)pidx) >= p.p_drop) ? pr * (1.f / (1.f - p.p_drop)) : 0.f;
              }
              if (p.probs_dropped) p.probs_dropped[pidx] = pr;
            }
            ssm[128 + r] = valid ? pr : 0.f;
            named_bar_sync(2 + 2 * wg, 128);
            // o_i = sum_j w_ij v_j for the head's 64 features: thread = (query ii = r / 32, feature pair r % 32)
            const int ii2 = r >> 5, d2 = r & 31;
            if (ii2 < bi && ri0 + ii2 < Nn && m_blk < p.m_tiles) {
              const uint8_t* vbx = qb + (BN / 128) * (p.bi8 + p.bj8) * 128 + wg * p.bj8 * 128;   // 64-dim chunk of head wg
              const int e = 2 * d2;
              float o0 = 0.f, o1 = 0.f;
#pragma unroll 8
              for (int jj = 0; jj < bj; ++jj) {
                const float w = ssm[128 + jj * bi + ii2];
                const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(vbx + jj * 128 + (((e >> 3) ^ (jj & 7)) << 4) + (e & 7) * 2);
                const float2 vf = __bfloat1622float2(v2);
                o0 = fmaf(w, vf.x, o0);
                o1 = fmaf(w, vf.y, o1);
              }
              const long orow = (long)(ri0 + ii2) * Bb + rb_;
              *reinterpret_cast<float2*>(p.att + orow * p.ldatt + head * 64 + e) = make_float2(o0, o1);
              if (p.att_b) *reinterpret_cast<uint32_t*>(p.att_b + orow * p.ldatt + head * 64 + e) = pack_bf16x2(o0, o1);
            }
          }
        }
      }

      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!(MODE == MODE_SCORE && p.fuse)) {      // fused attention: released before the softmax / P.V tail
          if (CG == 2 && !leader) mbar_arrive_remote(&bars->tempty[as], 0); else mbar_arrive(&bars->tempty[as]);
        }
        if (REL) mbar_arrive(&bars->qempty[qs]);
      }
      if (++as == 2) { as = 0; aph ^= 1; }
      if (REL) { if (++qs == 2) { qs = 0; qph ^= 1; } }
    }
    if (warp == 2 && lane == 0) GTOS_TRACE(7);
    if (p.tma_out && quarter == 2 && lane == 0) tma_store_wait_all();  // smem must outlive the bulk stores
    if (warp == 2 && lane == 0) GTOS_TRACE(8);
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // the peer may still multicast into our barriers / read our smem until here
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0) GTOS_TRACE(9);
}

int debug_read_trace(unsigned long long* host_out, int n) {
  if (n > 148 * 16) n = 148 * 16;
  GTOS_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_trace, sizeof(unsigned long long) * n));
  return GTOS_OK;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int choose_rel_tiling(RelTiling* t, int N, int B, int D, int H) {
  if (N <= 0 || B <= 0 || D <= 0 || H <= 0 || D % H != 0) {
    set_error("bad relation-attention shape N=%d B=%d D=%d H=%d", N, B, D, H);
    return GTOS_ERR_ARG;
  }
  int hd = D / H;
  if (D % 128 != 0 || !(hd == 16 || hd == 32 || hd == 64 || hd == 128)) {
    set_error("relation attention needs D %% 128 == 0 and head_dim in {16,32,64,128} (got D=%d hd=%d)", D, hd);
    return GTOS_ERR_UNSUPPORTED;
  }
  long best_tiles = -1;
  int best_bi = 1, best_bj = 1;
  for (int bi = 1; bi <= 128 && bi <= N; ++bi) {
    for (int bj = 1; bj * bi <= 128 && bj <= N; ++bj) {
      int bi8 = (bi + 7) & ~7, bj8 = (bj + 7) & ~7;
      if (bi8 + bj8 > 48) continue;  // q/k staging budget: 2*(bi8+bj8)*128 B per stage <= 12 KB (bf16 q/k)
      long tiles = (long)((N + bi - 1) / bi) * ((N + bj - 1) / bj);
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && bi + bj < best_bi + best_bj)) {
        best_tiles = tiles;
        best_bi = bi;
        best_bj = bj;
      }
    }
  }
  // A tile that holds EVERY key of its queries (bj = N, up to four queries) lets the score kernel finish the attention
  // in its epilogue (gtos_rel_attn_fwd: softmax rows are complete inside the tile).  Take it whenever it costs no extra
  // tiles - it needs a larger q / k / v staging area (the 48-row budget above does not apply), paid for with fewer
  // stages of the main A/B ring.  N = 41 -> 3 x 41 (14 tiles per graph, same as 6 x 21), N = 61 -> 2 x 61.
  // (only when the fused tail is switched on - GTOS_REL_FUSED_FWD=1: it measured 0.16 ms per step SLOWER than the
  // two-kernel path at config 2, because the epilogue warps already are this kernel's bottleneck, so the default keeps the
  // tiling with the smaller staging area and the deeper A/B ring.)
  static const bool full_rows = getenv("GTOS_REL_FUSED_FWD") && getenv("GTOS_REL_FUSED_FWD")[0] == '1';
  if (full_rows && N <= 128) {
    int bif = 128 / N;
    if (bif > 4) bif = 4;
    if (bif > N) bif = N;
    if (bif >= 1) {
      long tiles = (long)((N + bif - 1) / bif);
      if (tiles <= best_tiles && (((bif + 7) & ~7) + 2 * ((N + 7) & ~7)) * 256 * 2 <= 96 * 1024) {
        best_tiles = tiles;
        best_bi = bif;
        best_bj = N;
      }
    }
  }
  t->N = N; t->B = B; t->D = D; t->H = H; t->hd = hd;
  t->bi = best_bi; t->bj = best_bj;
  t->ni_blk = (N + best_bi - 1) / best_bi;
  t->nj_blk = (N + best_bj - 1) / best_bj;
  t->tiles = B * t->ni_blk * t->nj_blk;
  t->scale = 1.0f / sqrtf((float)hd);
  return GTOS_OK;
}

static int g_num_sms = 0;
static int g_sm_reserve = -1;   // SMs the persistent tcgen05 kernels leave free (for NCCL CTAs running beside them)
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_num_sms = 148;
  }
  if (g_sm_reserve < 0) g_sm_reserve = getenv("GTOS_SM_RESERVE") ? atoi(getenv("GTOS_SM_RESERVE")) : 0;
  int n = g_num_sms - g_sm_reserve;
  n &= ~1;                      // CTA pairs
  return n < 2 ? 2 : n;
}
int set_sm_reserve(int n) {
  if (n < 0 || n > 128) {
    set_error("sm_reserve must be in [0, 128] (got %d)", n);
    return GTOS_ERR_ARG;
  }
  g_sm_reserve = n;
  return GTOS_OK;
}

bool rel_attn_fusable(const RelTiling& rt) {
  // all keys of a query in one tile (softmax rows complete), one 64-wide head per epilogue warpgroup of a 256-column unit,
  // and one thread per (query, feature pair) of the head for the P.V product
  return rt.nj_blk == 1 && rt.bj == rt.N && rt.hd == 64 && rt.bi * 32 <= 128 && rt.D % 128 == 0;
}

int make_rel_tmaps(const RelTiling& rt, const void* relb, const void* q, const void* k, long ldqk,
                   CUtensorMap* tmA, CUtensorMap* tmQ, CUtensorMap* tmK) {
  {
    uint64_t dims[4] = {(uint64_t)rt.D, (uint64_t)rt.B, (uint64_t)rt.N, (uint64_t)rt.N};
    uint64_t str[4] = {0, (uint64_t)rt.D * 2, (uint64_t)rt.B * rt.D * 2, (uint64_t)rt.N * rt.B * rt.D * 2};
    uint32_t box[4] = {64, 1, (uint32_t)rt.bi, (uint32_t)rt.bj};
    int e = make_tmap_nd(tmA, relb, 2, 4, dims, str, box, true);
    if (e) return e;
  }
  if (q) {
    uint64_t dims[3] = {(uint64_t)rt.D, (uint64_t)rt.B, (uint64_t)rt.N};
    uint64_t str[3] = {0, (uint64_t)ldqk * 2, (uint64_t)rt.B * ldqk * 2};
    uint32_t boxq[3] = {64, 1, (uint32_t)rt.bi};
    uint32_t boxk[3] = {64, 1, (uint32_t)rt.bj};
    int e = make_tmap_nd(tmQ, q, 2, 3, dims, str, boxq, true);
    if (e) return e;
    e = make_tmap_nd(tmK, k, 2, 3, dims, str, boxk, true);
    if (e) return e;
  }
  return GTOS_OK;
}

template <int BN, int MODE, int CG = 1>
static int launch_tn(const GemmTnArgs& a, cudaStream_t stream) {
  constexpr bool REL = (MODE == MODE_SCORE || MODE == MODE_GRAD);
  TnDev p;
  memset(&p, 0, sizeof(p));
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.n_tiles = (a.N + BN - 1) / BN;
  p.k_blocks = (a.K + BK - 1) / BK;
  p.bias = a.bias; p.out_f32 = a.out_f32; p.ldo = a.ldo;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16); p.ldob = a.ldob;
  p.relu = a.relu; p.accumulate = a.accumulate; p.addend = a.addend; p.ldadd = a.ldadd;
  p.rt = a.rt; p.scores = a.scores; p.dscores = a.dscores; p.G = reinterpret_cast<__nv_bfloat16*>(a.G);
  static const int dbg = getenv("GTOS_DBG") ? atoi(getenv("GTOS_DBG")) : 0;
  p.dbg = dbg;
  CUtensorMap tmA, tmB, tmQ, tmK;
  int e;
  CUtensorMap tmV;
  bool have_v = false;
  if (REL) {
    p.m_tiles = a.rt.tiles;
    p.bi8 = (a.rt.bi + 7) & ~7;
    p.bj8 = (a.rt.bj + 7) & ~7;
    if (MODE == MODE_SCORE && a.fuse) {
      GTOS_REQUIRE(rel_attn_fusable(a.rt) && BN == 256, "rel_attn_fwd: shape not fusable (N=%d D=%d H=%d)", a.rt.N, a.rt.D, a.rt.H);
      GTOS_REQUIRE(a.v && a.probs && a.att && a.ldv % 8 == 0 && a.ldatt % 2 == 0, "rel_attn_fwd: v / probs / att are required");
      GTOS_REQUIRE(a.p_drop == 0.f || a.seed_ptr, "rel_attn_fwd: dropout needs a device seed pointer");
      p.fuse = 1;
      p.key_pad = a.key_pad; p.p_drop = a.p_drop; p.seed_ptr = a.seed_ptr; p.seed_off = a.seed_off;
      p.probs = a.probs; p.probs_dropped = a.probs_dropped; p.att = a.att; p.ldatt = a.ldatt;
      p.att_b = reinterpret_cast<__nv_bfloat16*>(a.att_bf16);
      const RelTiling& rt = a.rt;
      uint64_t dims[3] = {(uint64_t)rt.D, (uint64_t)rt.B, (uint64_t)rt.N};
      uint64_t str[3] = {0, (uint64_t)a.ldv * 2, (uint64_t)rt.B * a.ldv * 2};
      uint32_t boxv[3] = {64, 1, (uint32_t)rt.bj};
      e = make_tmap_nd(&tmV, a.v, 2, 3, dims, str, boxv, true);
      if (e) return e;
      have_v = true;
    }
    p.qk_stage_bytes = (BN / 128) * (p.bi8 + p.bj8 * (p.fuse ? 2 : 1)) * 128;
    e = make_rel_tmaps(a.rt, a.A, a.q, a.k, a.ldqk, &tmA, &tmQ, &tmK);
    if (e) return e;
  } else {
    p.m_tiles = (MODE == MODE_DREL) ? a.rt.tiles : (a.M + BM - 1) / BM;
    e = make_tmap_2d_bf16(&tmA, a.A, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda, BM);
    if (e) return e;
    tmQ = tmA;
    tmK = tmA;
  }
  e = make_tmap_2d_bf16(&tmB, a.Bm, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb, BN / CG);
  if (e) return e;
  p.units = ((p.m_tiles + CG - 1) / CG) * p.n_tiles;   // CG = 2: units are (tile pair, n block)
  // ---- output path: TMA stores from swizzled staging tiles where the layout allows it ----
  CUtensorMap tmO = have_v ? tmV : tmB;
  constexpr int OUT_STAGE_BYTES = tn_out_stage_bytes<MODE>();
  p.tma_out = 0;
  static const bool grad_tma = !(getenv("GTOS_GRAD_TMA") && getenv("GTOS_GRAD_TMA")[0] == '0');
  static const bool tma_enabled = !(getenv("GTOS_TMA_OUT") && getenv("GTOS_TMA_OUT")[0] == '0');
  static const bool bf16_tma = !(getenv("GTOS_TMA_BF16") && getenv("GTOS_TMA_BF16")[0] == '0');
  if (!tma_enabled) {
  } else if (MODE == MODE_PLAIN && a.out_f32 && !a.out_bf16 && !a.accumulate && a.ldo % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(a.out_f32) & 15) == 0) {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    uint64_t str[2] = {0, (uint64_t)a.ldo * 4};
    uint32_t box[2] = {32, (uint32_t)BM};
    e = make_tmap_nd(&tmO, a.out_f32, 4, 2, dims, str, box, true);
    if (e) return e;
    p.tma_out = 1;
  } else if (MODE == MODE_PLAIN && bf16_tma && a.out_bf16 && !a.out_f32 && !a.accumulate && !a.addend && a.ldob % 8 == 0 &&
             (reinterpret_cast<uintptr_t>(a.out_bf16) & 15) == 0) {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    uint64_t str[2] = {0, (uint64_t)a.ldob * 2};
    uint32_t box[2] = {64, (uint32_t)BM};
    e = make_tmap_nd(&tmO, a.out_bf16, 2, 2, dims, str, box, true);
    if (e) return e;
    p.tma_out = 2;
  } else if (MODE == MODE_DREL && a.ldo == a.rt.D) {
    const RelTiling& rt = a.rt;
    uint64_t dims[4] = {(uint64_t)rt.D, (uint64_t)rt.B, (uint64_t)rt.N, (uint64_t)rt.N};
    uint64_t str[4] = {0, (uint64_t)rt.D * 4, (uint64_t)rt.B * rt.D * 4, (uint64_t)rt.N * rt.B * rt.D * 4};
    uint32_t box[4] = {32, 1, (uint32_t)rt.bi, (uint32_t)rt.bj};
    e = make_tmap_nd(&tmO, a.out_f32, 4, 4, dims, str, box, true);
    if (e) return e;
    p.tma_out = 1;
  } else if (MODE == MODE_GRAD && grad_tma) {
    uint64_t dims[2] = {(uint64_t)(2 * a.rt.D), (uint64_t)a.rt.tiles * BM};
    uint64_t str[2] = {0, (uint64_t)(2 * a.rt.D) * 2};
    uint32_t box[2] = {64, (uint32_t)BM};
    e = make_tmap_nd(&tmO, a.G, 2, 2, dims, str, box, true);
    if (e) return e;
    p.tma_out = 1;
  }
  constexpr int STAGE_BYTES = A_STAGE_BYTES + (BN / CG) * BK * 2;
  const int budget = 227 * 1024 - 1024 /*align*/ - (int)sizeof(PipeBars) - (REL ? 2 * p.qk_stage_bytes : 0) -
                     ((p.tma_out || p.fuse) ? OUT_STAGE_BYTES : 0);
  int stages = budget / STAGE_BYTES;
  if (stages > 6) stages = 6;
  if (stages < 2) {
    set_error("not enough shared memory for the GEMM pipeline");
    return GTOS_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const int smem_bytes = 1024 + stages * STAGE_BYTES + (REL ? 2 * p.qk_stage_bytes : 0) +
                         ((p.tma_out || p.fuse) ? OUT_STAGE_BYTES : 0) + (int)sizeof(PipeBars);
  auto kern = gemm_tn_kernel<BN, MODE, CG>;
  GTOS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  int grid = p.units * CG < num_sms() ? p.units * CG : (num_sms() / CG) * CG;
  if (grid <= 0) return GTOS_OK;
  GTOS_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(tn_threads<MODE>()), (size_t)smem_bytes, stream, CG, tmA, tmB, tmQ, tmK, tmO, p));
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int launch_gemm_tn(int mode, const GemmTnArgs& a, cudaStream_t stream) {
  GTOS_REQUIRE(a.K % 8 == 0 && a.lda % 8 == 0 && a.ldb % 8 == 0, "gemm_tn: K, lda, ldb must be multiples of 8 (K=%d lda=%ld ldb=%ld)",
               a.K, a.lda, a.ldb);
  GTOS_REQUIRE(!a.addend || (mode == MODE_PLAIN && a.out_f32 && !a.out_bf16 && !a.accumulate && a.ldo % 4 == 0),
               "gemm_tn: addend needs the fp32 TMA-store epilogue");
  static const bool pair = !(getenv("GTOS_REL_2CTA") && getenv("GTOS_REL_2CTA")[0] == '0');
  if (mode == MODE_SCORE) return pair ? launch_tn<256, MODE_SCORE, 2>(a, stream) : launch_tn<256, MODE_SCORE, 1>(a, stream);
  if (mode == MODE_GRAD) return pair ? launch_tn<256, MODE_GRAD, 2>(a, stream) : launch_tn<256, MODE_GRAD, 1>(a, stream);
  if (mode == MODE_DREL) {
    // d_relation = G Wperm as a CTA pair too: two pair tiles per unit, each CTA stages half of Wperm^T (L2 -> SM operand
    // traffic per FLOP down by a third): 142 -> 133 us per launch at config 2 (A/B on one box).  GTOS_DREL_CG2=0: single CTAs.
    static const bool drel_pair = !(getenv("GTOS_DREL_CG2") && getenv("GTOS_DREL_CG2")[0] == '0');
    return drel_pair ? launch_tn<256, MODE_DREL, 2>(a, stream) : launch_tn<256, MODE_DREL>(a, stream);
  }
  // plain: pick the N tile with the lowest estimated time.  Model fitted to `tools/gemm_probe.py --sweep` on a B200
  // (profiles/r02_gemm_tile_sweep.txt; CUDA-graph back-to-back launches of every plain-GEMM shape of the step at every tile
  // width):   t [us] = base(bn) + waves x (k_blocks x ck(bn) + e(bn)),   floored by the HBM time of the operands and the result.
  // A k-block costs ~0.32-0.36 us whatever the tile width - the pipeline is bound by the latency of the bytes in flight
  // (6 stages of <= 32 KB per SM), not by the MMA - so the widest tile that still fits ONE wave wins (fewer waves, and a
  // k-block of a 256-wide tile carries twice the FLOPs of a 128-wide one), unless its larger fixed cost (epilogue of a
  // 128 x 256 fp32 tile, ramp) outweighs it: [2624,1024] x K=512 takes bn = 256 (9.4 us against 9.8 at 128 and 12.0 at 64,
  // which the previous cycle-count estimate chose), [2624,512] x K=512 stays at 128 (7.0 against 8.9).
  // GTOS_TILE_MODEL=0 restores the previous estimate (MMA cycles + operand ingest).
  const long mt = (a.M + BM - 1) / BM;
  const int sms = num_sms();
  const long kb = (a.K + BK - 1) / BK;
  int best_bn = 64;
  double best = 1e30;
  const int cands[3] = {128, 256, 64};                              // ties go to the earlier candidate
  static const bool fitted = !(getenv("GTOS_TILE_MODEL") && getenv("GTOS_TILE_MODEL")[0] == '0');
  const double hbm_floor = ((double)a.M * a.N * 4 + (double)a.M * a.K * 2 + (double)a.N * a.K * 2) / 5.0e6;   // us at 5 TB/s
  for (int ci = 0; ci < 3; ++ci) {
    const int bn = cands[ci];
    if (bn > 64 && a.N <= bn / 2) continue;                       // tile mostly empty
    const long units = mt * ((a.N + bn - 1) / bn);
    const long waves = (units + sms - 1) / sms;
    double cost;
    if (fitted) {
      const double base = bn == 256 ? 5.4 : (bn == 128 ? 3.7 : 3.0);
      const double ck = bn == 256 ? 0.36 : (bn == 128 ? 0.35 : 0.32);
      double e = bn == 256 ? 1.4 : (bn == 128 ? 0.4 : 0.35);
      if (bn == 256 && kb < 8) e += (8 - kb) * 0.4;               // short K: the 128 x 256 epilogue is not hidden by the next tile
      cost = base + waves * (kb * ck + e);
      if (cost < hbm_floor) cost = hbm_floor;
    } else {
      double mma = (double)kb * 4 * (bn >= 128 ? bn / 2 : 48);  // UMMA floor M128: N/2 cycles per K=16; N=64 is smem-bound
      static const bool ingest = !(getenv("GTOS_TILE_INGEST") && getenv("GTOS_TILE_INGEST")[0] == '0');
      if (ingest) {                                               // (128 + bn) x 64 bf16 per k-block at ~100 B/clk per SM
        const double ing = (double)kb * (BM + bn) * BK * 2 / 100.0;
        if (ing > mma) mma = ing;
      }
      const double epi = (bn / 32) * 350.0;
      cost = waves * (mma > epi ? mma : epi) + 2500.0 + (mma > epi ? epi : mma);
    }
    if (cost < best) { best = cost; best_bn = bn; }
  }
  if (const char* f = getenv("GTOS_FORCE_BN")) {               // tools/gemm_probe.py --sweep: time every tile width
    const int v = atoi(f);
    if (v == 64 || v == 128 || v == 256) best_bn = v;
  }
  // CTA pairs (cta_group::2: each CTA stages its own A rows and HALF of the weight rows) pay ~1 us of cluster launch / sync
  // and only win where the L2 -> SM operand traffic is the bound: many waves of 256-wide tiles ([98636,512] x K=1536:
  // 149 -> 136 us; every small shape of the step is 0.7-1.1 us slower as a pair - profiles/r02_gemm_tile_sweep_pairs.txt)
  int cg = 1;
  {
    const long units256 = mt * ((a.N + 255) / 256);
    if (best_bn == 256 && units256 >= 8L * sms && kb >= 16) cg = 2;
  }
  if (const char* f = getenv("GTOS_FORCE_CG")) cg = atoi(f) == 2 ? 2 : 1;
  if (cg == 2 && mt >= 2 && !a.accumulate) {
    if (best_bn == 256) return launch_tn<256, MODE_PLAIN, 2>(a, stream);
    if (best_bn == 128) return launch_tn<128, MODE_PLAIN, 2>(a, stream);
    return launch_tn<64, MODE_PLAIN, 2>(a, stream);
  }
  if (best_bn == 256) return launch_tn<256, MODE_PLAIN>(a, stream);
  if (best_bn == 128) return launch_tn<128, MODE_PLAIN>(a, stream);
  return launch_tn<64, MODE_PLAIN>(a, stream);
}

template <int BN>
static int launch_gru_bn(const GruStepArgs& a, cudaStream_t stream) {
  TnDev p;
  memset(&p, 0, sizeof(p));
  const int H = a.H;
  p.M = a.R; p.N = 4 * H; p.K = a.Kx + H;
  p.m_tiles = (a.R + BM - 1) / BM;
  p.n_tiles = (4 * H + BN - 1) / BN;
  p.kx_blocks = a.Kx / BK;
  p.k_blocks = p.kx_blocks + (H + BK - 1) / BK;
  p.units = p.m_tiles * p.n_tiles;
  p.bias = a.bcat;
  p.gru_H = H; p.gru_t = a.t; p.gru_hprev = a.h_prev; p.gru_len = a.lengths; p.gru_hnew = a.h_new;
  p.gru_hbnew = reinterpret_cast<__nv_bfloat16*>(a.hb_new); p.gru_ldhbn = a.ldhbn;
  p.gru_out = reinterpret_cast<__nv_bfloat16*>(a.out_t); p.gru_ldout = a.ldout;
  p.gru_gates = reinterpret_cast<__nv_bfloat16*>(a.gates); p.gru_ldg = a.ldg;
  CUtensorMap tmA, tmB, tmH;
  int e = make_tmap_2d_bf16(&tmA, a.x, (uint64_t)a.R, (uint64_t)((a.Kin + 7) / 8 * 8), (uint64_t)a.ldx, BM);
  if (e) return e;
  e = make_tmap_2d_bf16(&tmH, a.hb, (uint64_t)a.R, (uint64_t)H, (uint64_t)a.ldhb, BM);
  if (e) return e;
  e = make_tmap_2d_bf16(&tmB, a.Wcat, (uint64_t)(4 * H), (uint64_t)(a.Kx + (H + 7) / 8 * 8), (uint64_t)a.ldw, BN);
  if (e) return e;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + BN * BK * 2;
  constexpr int GRU_STAGE_BYTES = 8 * 8192;   // warp-private store-staging tiles of the epilogue
  int stages = (227 * 1024 - 1024 - (int)sizeof(PipeBars) - 2 * BN * 4 - GRU_STAGE_BYTES) / STAGE_BYTES;
  if (stages > 6) stages = 6;
  p.stages = stages;
  const int smem_bytes = 1024 + stages * STAGE_BYTES + 2 * BN * 4 + GRU_STAGE_BYTES + (int)sizeof(PipeBars);
  auto kern = gemm_tn_kernel<BN, MODE_GRU, 1>;
  GTOS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  int grid = p.units < num_sms() ? p.units : num_sms();
  if (grid <= 0) return GTOS_OK;
  GTOS_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(tn_threads<MODE_GRU>()), (size_t)smem_bytes, stream, 1, tmA, tmB, tmH, tmH, tmB, p));
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int launch_gru_step(const GruStepArgs& a, cudaStream_t stream) {
  GTOS_REQUIRE(a.H % 16 == 0 && a.Kx % 64 == 0 && a.Kx >= a.Kin && a.ldx % 8 == 0 && a.ldhb % 8 == 0 && a.ldw % 8 == 0,
               "gru_step: need H %% 16 == 0, Kx = 64*ceil(Kin/64), 8-element row strides (H=%d Kx=%d Kin=%d)", a.H, a.Kx, a.Kin);
  if (a.R == 0) return GTOS_OK;
  if (a.H % 64 == 0) return launch_gru_bn<256>(a, stream);
  return launch_gru_bn<64>(a, stream);
}

// =======================================================================================
// gemm_nn: C[M,N] = sum_k A[k,m] * B[k,n]   (contraction over rows; both operands MN-major)
// =======================================================================================
struct NnDev {
  int M, N, Kd;
  int m_tiles, n_tiles, k_blocks, splits, kb_per_split;
  int stages;
  float* out;  // [M][N] fp32, zero-initialised: every split reduces into it with vector atomics
  long ldo;
  int perm_D, perm_hd;  // rel: accumulator row m is permuted row -> reference row order
  RelTiling rt;
};

template <int BN, int KB, int REL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_nn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const NnDev p) {
  // smem stage: A = (BM/64) boxes of [KB rows x 128 B]; B = (BN/64) boxes of [KB rows x 128 B]
  constexpr int BOX_BYTES = KB * 128;
  constexpr int A_BYTES = (BM / 64) * BOX_BYTES;
  constexpr int B_BYTES = (BN / 64) * BOX_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  constexpr uint32_t IDESC = make_idesc_bf16(BM, BN, 1, 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  PipeBars* bars = reinterpret_cast<PipeBars*>(smem + p.stages * STAGE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
  const int split = blockIdx.y;
  const int kb0 = split * p.kb_per_split;
  int kb1 = kb0 + p.kb_per_split;
  if (kb1 > p.k_blocks) kb1 = p.k_blocks;

  if (REL) {
    // relation tiles carry only bi*bj (< 128) rows per k-block: rows the TMA box never writes must be 0
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int o = threadIdx.x * 16; o < p.stages * STAGE_BYTES; o += GEMM_THREADS * 16)
      *reinterpret_cast<uint4*>(smem + o) = z;
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        wait_bar(&bars->empty[s], ph ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        int b = 0, j0 = 0, i0 = 0;
        uint32_t bbytes = B_BYTES;
        if (REL) {
          rel_tile_decode(p.rt, kb, b, j0, i0);
          bbytes = (uint32_t)((BN / 64) * p.rt.bi * p.rt.bj * 128);
        }
        mbar_expect_tx(&bars->full[s], (uint32_t)A_BYTES + bbytes);
#pragma unroll
        for (int c = 0; c < BM / 64; ++c) tma_load_2d(&tmA, &bars->full[s], sa + c * BOX_BYTES, m_blk * BM + c * 64, kb * KB);
#pragma unroll
        for (int c = 0; c < BN / 64; ++c) {
          if (REL)
            tma_load_4d(&tmB, &bars->full[s], sb + c * BOX_BYTES, n_blk * BN + c * 64, b, i0, j0);
          else
            tma_load_2d(&tmB, &bars->full[s], sb + c * BOX_BYTES, n_blk * BN + c * 64, kb * KB);
        }
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        wait_bar(&bars->full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        // MN-major, 128B swizzle: LBO = next 64-element chunk along M/N (one box), SBO = next 8 k-rows
        const uint64_t da = make_smem_desc_sw128(sa, BOX_BYTES, 1024);
        const uint64_t db = make_smem_desc_sw128(sa + A_BYTES, BOX_BYTES, 1024);
#pragma unroll
        for (int k = 0; k < KB / 16; ++k) {
          // 16 k-rows = 2048 bytes = +128 in the (addr>>4) field
          umma_bf16(tmem_base, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), IDESC,
                    (kb > kb0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&bars->empty[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(&bars->tfull[0]);
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const bool any = kb1 > kb0;
    if (any) {
      wait_bar(&bars->tfull[0], 0);
      tc_fence_after();
    }
    const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const long m = (long)m_blk * BM + r;
    const long mo = (p.perm_D && m < p.M) ? rel_perm_to_orig((int)m, p.perm_D, p.perm_hd) : m;
    float* orow = p.out + mo * p.ldo + (long)n_blk * BN;
    if (any) {
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        if (n_blk * BN + c >= p.N) break;
        float v[16];
        tmem_ld16(tacc + c, v);
        tmem_ld_wait();
        if (m < p.M) {
          const int n = n_blk * BN + c;
          if (n + 16 <= p.N && ((reinterpret_cast<uintptr_t>(orow + c) & 15) == 0)) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + c + 4 * t), "f"(v[4 * t]),
                           "f"(v[4 * t + 1]), "f"(v[4 * t + 2]), "f"(v[4 * t + 3])
                           : "memory");
          } else {
#pragma unroll
            for (int t = 0; t < 16; ++t)
              if (n + t < p.N) atomicAdd(orow + c + t, v[t]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static void nn_plan(int M, int N, int Kd, int rel, int* BN, int* KB, int* splits, int* kb_per) {
  *BN = (N >= 256) ? 256 : ((N > 64 || rel) ? 128 : 64);
  if (const char* f = getenv("GTOS_FORCE_NN_BN")) {
    const int v = atoi(f);
    if (!rel && (v == 64 || v == 128 || v == 256)) *BN = v;
  }
  *KB = rel ? 128 : 64;
  int tiles = ((M + BM - 1) / BM) * ((N + *BN - 1) / *BN);
  int kblocks = (Kd + *KB - 1) / *KB;
  int want = num_sms() / tiles;  // floor: one wave (ceil gave 160 CTAs on 148 SMs = 2 waves)
  if (want > kblocks) want = kblocks;
  if (want < 1) want = 1;
  int per = (kblocks + want - 1) / want;
  if (per < 1) per = 1;
  // Small contractions (the weight gradients of a 2.6-3.8 k row Linear: 41-60 k-blocks over 8 tiles) used to be split 18
  // ways: 144 CTAs of 3-4 k-blocks each, every one of them reducing its whole 128 x BN fp32 tile into the output with
  // vector atomics - 18 MB of atomic traffic for a 1 MB result (CUPTI timeline: 24.5 us per launch, atomics-bound) on
  // all SMs, while the kernel only runs BESIDE the critical chain (side stream).  At least NN_MIN_KB k-blocks per CTA:
  // 6-8 splits, a third of the atomics, and more than half of the SMs stay with the main chain.  Large contractions
  // (GRU weight gradients: 1 541 k-blocks) keep one full wave.
  static const int min_kb = getenv("GTOS_NN_MIN_KB") ? atoi(getenv("GTOS_NN_MIN_KB")) : 8;
  if (!rel && per < min_kb) per = kblocks < min_kb ? kblocks : min_kb;
  // (tile width, split count) by a cost model fitted to `tools/gemm_probe.py --sweep-nn`
  // (profiles/r02_gemm_nn_sweep.txt):  t [us] = 2.1 + waves x (k-blocks per CTA x 0.35 + bn x 0.027) + 0.15 x splits,
  // floored by the HBM time of the operands.  The atomic epilogue costs in proportion to the tile width and a k-block
  // ~0.35 us whatever the width, so narrow tiles with few splits win at the step's sizes (dW [512,512] over 2624 rows:
  // 8.3 us at bn 64 / 4 splits against 11.0 us at bn 256 / 6 splits; every weight-gradient shape of the step 5-25 % faster
  // alone, the cfg2 step 0.04 ms faster).  GTOS_NN_MODEL=0 restores the previous plan (one wave, >= GTOS_NN_MIN_KB k-blocks).
  static const bool nn_model = !(getenv("GTOS_NN_MODEL") && getenv("GTOS_NN_MODEL")[0] == '0');
  if (nn_model && !rel) {
    const int sms = num_sms();
    const int mt = (M + BM - 1) / BM;
    const double floor_us = ((double)Kd * (M + N) * 2 + (double)M * N * 4) / 5.0e6;
    double best = 1e30;
    int best_bn = *BN, best_per = per;
    const int cands[3] = {64, 128, 256};
    for (int ci = 0; ci < 3; ++ci) {
      const int bn = cands[ci];
      if (bn > 64 && N <= bn / 2) continue;
      const int tl = mt * ((N + bn - 1) / bn);
      for (int sp = 1; sp <= 18 && sp <= kblocks; ++sp) {
        const int pr = (kblocks + sp - 1) / sp;
        const int real_sp = (kblocks + pr - 1) / pr;
        const long ctas = (long)tl * real_sp;
        const long waves = (ctas + sms - 1) / sms;
        double c = 2.1 + waves * (pr * 0.35 + bn * 0.027) + 0.15 * real_sp;
        if (c < floor_us) c = floor_us + 0.01 * real_sp;
        if (c < best) { best = c; best_bn = bn; best_per = pr; }
      }
    }
    *BN = best_bn;
    per = best_per;
  }
  if (const char* f = getenv("GTOS_FORCE_NN_SPLITS")) {        // tools/gemm_probe.py --sweep-nn
    const int v = atoi(f);
    if (v >= 1 && !rel) per = (kblocks + v - 1) / v;
  }
  *kb_per = per;
  *splits = (kblocks + per - 1) / per;
  if (*splits < 1) *splits = 1;
}

long gemm_nn_workspace_elems(int M, int N, int Kd, int rel) {
  (void)M; (void)N; (void)Kd; (void)rel;
  return 0;  // split-K now reduces in place (vector atomics); kept in the ABI for callers that size a workspace
}

template <int BN, int KB, int REL>
static int launch_nn(const GemmNnArgs& a, int splits, int per, cudaStream_t stream) {
  NnDev p;
  memset(&p, 0, sizeof(p));
  p.M = a.M; p.N = a.N; p.Kd = a.Kd;
  p.m_tiles = (a.M + BM - 1) / BM;
  p.n_tiles = (a.N + BN - 1) / BN;
  p.k_blocks = (a.Kd + KB - 1) / KB;
  p.splits = splits; p.kb_per_split = per;
  p.out = a.out; p.ldo = a.ldo; p.rt = a.rt;
  p.perm_D = a.rel ? a.rt.D : a.perm_D; p.perm_hd = a.rel ? a.rt.hd : a.perm_hd;
  CUtensorMap tmA, tmB;
  int e;
  {
    uint64_t dims[2] = {(uint64_t)a.M, (uint64_t)a.Kd};
    uint64_t str[2] = {0, (uint64_t)a.lda * 2};
    uint32_t box[2] = {64, (uint32_t)KB};
    e = make_tmap_nd(&tmA, a.A, 2, 2, dims, str, box, true);
    if (e) return e;
  }
  if (REL) {
    const RelTiling& rt = a.rt;
    uint64_t dims[4] = {(uint64_t)rt.D, (uint64_t)rt.B, (uint64_t)rt.N, (uint64_t)rt.N};
    uint64_t str[4] = {0, (uint64_t)rt.D * 2, (uint64_t)rt.B * rt.D * 2, (uint64_t)rt.N * rt.B * rt.D * 2};
    uint32_t box[4] = {64, 1, (uint32_t)rt.bi, (uint32_t)rt.bj};
    e = make_tmap_nd(&tmB, a.Bm, 2, 4, dims, str, box, true);
  } else {
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.Kd};
    uint64_t str[2] = {0, (uint64_t)a.ldb * 2};
    uint32_t box[2] = {64, (uint32_t)KB};
    e = make_tmap_nd(&tmB, a.Bm, 2, 2, dims, str, box, true);
  }
  if (e) return e;
  constexpr int STAGE_BYTES = (BM / 64 + BN / 64) * KB * 128;
  int stages = (227 * 1024 - 1024 - (int)sizeof(PipeBars)) / STAGE_BYTES;
  if (stages > 6) stages = 6;
  if (stages < 2) {
    set_error("gemm_nn: pipeline does not fit in shared memory");
    return GTOS_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const int smem_bytes = 1024 + stages * STAGE_BYTES + (int)sizeof(PipeBars);
  auto kern = gemm_nn_kernel<BN, KB, REL>;
  GTOS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  // the split-K partial sums are reduced straight into the (zeroed) output with 16-byte vector atomics
  if (a.ldo == a.N)
    GTOS_CHECK_CUDA(cudaMemsetAsync(a.out, 0, sizeof(float) * (size_t)a.M * a.N, stream));
  else
    GTOS_CHECK_CUDA(cudaMemset2DAsync(a.out, sizeof(float) * a.ldo, 0, sizeof(float) * a.N, a.M, stream));
  dim3 grid(p.m_tiles * p.n_tiles, splits);
  GTOS_CHECK_CUDA(launch_pdl(kern, grid, dim3(GEMM_THREADS), (size_t)smem_bytes, stream, 1, tmA, tmB, p));
  GTOS_LAUNCH_CHECK();
  return GTOS_OK;
}

int launch_gemm_nn(const GemmNnArgs& a, cudaStream_t stream) {
  GTOS_REQUIRE(a.lda % 8 == 0 && (a.rel || a.ldb % 8 == 0), "gemm_nn: lda/ldb must be multiples of 8");
  int BN, KB, splits, per;
  nn_plan(a.M, a.N, a.Kd, a.rel, &BN, &KB, &splits, &per);
  if (a.rel) {
    if (BN == 256) return launch_nn<256, 128, 1>(a, splits, per, stream);
    return launch_nn<128, 128, 1>(a, splits, per, stream);
  }
  if (BN == 256) return launch_nn<256, 64, 0>(a, splits, per, stream);
  if (BN == 128) return launch_nn<128, 64, 0>(a, splits, per, stream);
  return launch_nn<64, 64, 0>(a, splits, per, stream);
}

}  // namespace gtos
