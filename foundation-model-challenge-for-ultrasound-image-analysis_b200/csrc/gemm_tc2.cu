// Persistent, warp-specialised bf16 GEMM on the 5th-gen tensor cores (the main engine of the bf16 mode).
//
//   * one CTA per SM, static round-robin tile scheduler over (split, m-tile, n-tile) work items, n fastest so
//     the CTAs running at the same time share A row-blocks / the whole weight matrix in L2;
//   * warp 0 = TMA producer (A and B tiles, SWIZZLE_128B, 4-stage mbarrier ring that runs ahead across tiles),
//     warp 1 = tcgen05.mma issuer (one elected thread, fp32 accumulators in TMEM, DOUBLE-BUFFERED: the epilogue of
//     tile i overlaps the main loop of tile i+1), warp 2 = TMEM allocator, warp 3 = TMA loader of the epilogue
//     operand (residual / saved pre-activation), warps 4-7 = epilogue;
//   * epilogue: tcgen05.ld -> registers -> bias / GELU / GELU' / drop-path scale / residual -> swizzled shared
//     memory -> TMA store (bf16) or TMA reduce-add (fp32, split-K weight gradients).  No per-thread global
//     loads or stores: every byte of HBM traffic of this kernel moves through TMA, fully coalesced.
//
// Operand forms: K-major / MN-major 2-D tiles, implicit-im2col 4-D boxes (3x3 convolutions), and the conv3x3 weight
// gradient (both operands pixel-major 4-D boxes).  Replaces the cuBLAS / cuDNN calls behind timm's nn.Linear
// layers and smp's FPN convolutions.
//
// Roofline: tensor pipe for K >= 512 (tcgen05 floor: 64 cycles per 128x128x16 MMA), HBM for the K = 128 / 256
// layers of stages 1-2 whose epilogue traffic dominates (algorithmic bytes = A + B + every epilogue tensor once).
#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"
#include <stdlib.h>

#define T2_BM 128
#define T2_BK 64
// Epilogue warps: T2_NQ column groups x 4 TMEM lane quarters.  With 8 warps (T2_NQ = 2) the epilogues that do real math
// (GELU / GELU' at K = 128 / 256, where the main loop is short) ran at ~47 % issue-slot utilisation: two warps per
// scheduler cannot cover the MUFU / LDS / TMEM-load latencies.  16 warps (T2_NQ = 4) give every scheduler four epilogue
// warps and halve the accumulator registers per thread (16 columns instead of 32).
#ifndef T2_NQ
#define T2_NQ 4
#endif
#define T2_EPI_WARPS (4 * T2_NQ)
#define T2_EPI_THREADS (32 * T2_EPI_WARPS)
#define T2_THREADS (128 + T2_EPI_THREADS)
#define T2_EPI_BAR 1

struct T2Conv {   // geometry of conv operands (NHWC [B,H,W,C]); pixel tiles of th x tw
  int H, W, C;
  int th, tw, tiles_x, tiles_y;        // 128-pixel tile of the M dimension (fwd / dgrad)
  int kth, ktw, ktiles_x, ktiles_y;    // 64-pixel tile of the K dimension (wgrad)
};

struct T2Epi {
  const float* bias;      // [N] or null
  const float* rowscale;  // per-sample scale or null
  int rows_per_sample;
  int act;                // 0 none | 1 GELU (pre-activation -> X out) | 2 multiply by GELU'(X in) | 3 GELU (GELU' -> X out) | 4 multiply by X in
  int x_mode;             // 0 none | 1 residual in (out += X) | 2 aux in (act 2) | 3 aux out (act 1)
  int reduce;             // fp32 output: 1 = TMA reduce-add into the destination (split-K weight gradients), 0 = store
  float* colsum;          // bf16 output only: colsum[n] += sum over rows of the fp32 (un-rounded) output values (bias gradients)
};

template <int BN, bool OUT_F32, int CS = 1>
struct T2Smem {
  static constexpr int A_BYTES = T2_BM * T2_BK * 2;
  static constexpr int B_BYTES = (BN / CS) * T2_BK * 2;             // CTA pair (CS = 2): each CTA stages half of the B tile
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int SUB_BYTES = 128 * 128;                       // epilogue sub-tile: 128 rows x 128 bytes
  static constexpr int EPI_BUFS = 2;   // (measured: a deeper ring with a single-buffered epilogue is not faster)
  static constexpr int STAGES = (CS == 2) ? 4 : ((BN == 256) ? 3 : ((BN >= 128) ? 4 : 6));
  static constexpr int TMEM_COLS = (2 * BN > 256) ? 512 : ((2 * BN > 128) ? 256 : 128);   // two accumulators, power of two
  static constexpr int EPI_OFF = STAGES * STAGE;
  static constexpr int BAR_OFF = EPI_OFF + 2 * EPI_BUFS * SUB_BYTES;   // X[EPI_BUFS], D[EPI_BUFS]
  static constexpr int BIAS_OFF = BAR_OFF + 256;                       // bias of this / the next tile's BN columns (fp32) [2][BN]
  static constexpr int TOTAL = BIAS_OFF + 2 * BN * 4 + 1024;
  static constexpr int SUB_COLS = OUT_F32 ? 32 : 64;
  static constexpr int NSUB = BN / SUB_COLS;
};

// A_MODE: 0 K-major 2-D | 1 MN-major 2-D | 2 conv im2col (K-major, 128-pixel row tiles) | 3 conv pixel-major (wgrad)
// B_MODE: 0 K-major 2-D | 1 MN-major 2-D | 2 conv pixel-major with tap shift (wgrad)
// CS = 2: CTA PAIR (tcgen05 cta_group::2).  The two CTAs of a cluster work on two consecutive m-blocks of the same n-tile:
// each loads its own 128-row A tile and HALF of the B tile into its own shared memory (transaction bytes signalled on the
// leader's barrier), the leader (cluster rank 0) issues 256 x BN MMAs that read both CTAs' shared memory and leave 128
// accumulator rows in each CTA's TMEM, and each CTA runs its own epilogue.  Per SM and k-block the operand traffic drops
// from A + B to A + B/2 -- these GEMMs are bound by the L2 -> SM operand rate.  (Round 1 tried multicasting B halves to
// both CTAs with cta_group::1 MMAs: the same bytes still enter every SM, so it did not help and was removed.)
template <int BN, int A_MODE, int B_MODE, bool OUT_F32, int CS>
__global__ void __launch_bounds__(T2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, int M, int N, int K,
                int kb_per_split, int total_kb, int m_tiles, int n_tiles, int n_work, T2Conv cv, T2Epi ep) {
  using S = T2Smem<BN, OUT_F32, CS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  // barrier map: full[STAGES] empty[STAGES] tfull[2] tempty[2] xfull[2] xempty[2]
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * S::STAGES, bar_tfull = bar_empty + 8 * S::STAGES,
                 bar_tempty = bar_tfull + 16, bar_xfull = bar_tempty + 16, bar_xempty = bar_xfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 8);
  const uint32_t smem_x = smem_base + S::EPI_OFF, smem_d = smem_x + S::EPI_BUFS * S::SUB_BYTES;
  float* s_bias = reinterpret_cast<float*>(smem + S::BIAS_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool x_in = (ep.x_mode == 1 || ep.x_mode == 2);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
    if (ep.x_mode) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  const int cl_rank = (CS > 1) ? (int)cluster_ctarank() : 0;
  const int cl_id = blockIdx.x / CS, n_cl = gridDim.x / CS;
  constexpr uint16_t cl_mask = (uint16_t)((1u << CS) - 1);
  if (warp == 1 && lane == 0) {
    // pair mode: full / tempty are used in the leader only (fed by both CTAs); empty / tfull get one multicast commit each
    for (int s = 0; s < S::STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, CS * T2_EPI_WARPS);
      mbar_init(bar_xfull + 8 * b, 1); mbar_init(bar_xempty + 8 * b, T2_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) { if (CS == 2) tmem_alloc_2cta(smem_u32(tmem_slot), S::TMEM_COLS); else tmem_alloc(smem_u32(tmem_slot), S::TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();          // barrier inits visible cluster-wide before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();                              // everything above overlapped the tail of the previous kernel

  // work units enumerate (split z, m-group, n-tile), n fastest; the CTA of rank r takes m-block mg*CS + r (a block
  // beyond m_tiles is a dummy: its loads are zero-filled and its stores clipped, it only keeps the cluster in step)
  const int mgn_tiles = ((m_tiles + CS - 1) / CS) * n_tiles;
#define T2_DECODE(u)                                                                                   \
  const int z = (u) / mgn_tiles, r_ = (u) - z * mgn_tiles, m_blk = (r_ / n_tiles) * CS + cl_rank,      \
            n_blk = r_ - (r_ / n_tiles) * n_tiles;                                                     \
  (void)z; (void)m_blk; (void)n_blk;

  if (warp == 0 && lane == 0) {
    // ================================ TMA producer: A / B tiles ================================
    uint32_t it = 0;
    for (int t = cl_id; t < n_work; t += n_cl) {
      T2_DECODE(t)
      const int kb0 = z * kb_per_split, nkb = min(kb_per_split, total_kb - kb0);
      int cb = 0, cy0 = 0, cx0 = 0;
      if (A_MODE == 2) {
        int q = m_blk; const int tx = q % cv.tiles_x; q /= cv.tiles_x; const int ty = q % cv.tiles_y; cb = q / cv.tiles_y;
        cy0 = ty * cv.th; cx0 = tx * cv.tw;
      }
      for (int i = 0; i < nkb; ++i, ++it) {
        const uint32_t s = it % S::STAGES, ph = (it / S::STAGES) & 1;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t full = bar_full + 8 * s;
        const uint32_t sa = smem_base + s * S::STAGE, sb = sa + S::A_BYTES;
        const int kb = kb0 + i, k0 = kb * T2_BK;
        if (CS == 2) {
          // CTA pair: only the leader's barrier counts, and it expects the bytes of BOTH CTAs
          static_assert(CS == 1 || (A_MODE == 0 && B_MODE != 2), "pair mode: K-major A, 2-D B");
          if (cl_rank == 0) mbar_expect_tx(full, 2 * S::STAGE);
          tma_load_2d_2cta(sa, &tmA, k0, m_blk * T2_BM, full);
          if (B_MODE == 0) {          // tmB box = (64 k, BN/2 rows)
            tma_load_2d_2cta(sb, &tmB, k0, n_blk * BN + cl_rank * (BN / 2), full);
          } else {                    // MN-major: BN/128 chunks of 64 n per CTA
#pragma unroll
            for (int j = 0; j < BN / 128; ++j)
              tma_load_2d_2cta(sb + j * 8192, &tmB, n_blk * BN + cl_rank * (BN / 2) + j * 64, k0, full);
          }
          continue;
        }
        mbar_expect_tx(full, S::STAGE);
        int pb = 0, py0 = 0, px0 = 0;                    // pixel patch of this k-block (conv wgrad)
        if (A_MODE == 3 || B_MODE == 2) {
          int q = kb; const int tx = q % cv.ktiles_x; q /= cv.ktiles_x; const int ty = q % cv.ktiles_y; pb = q / cv.ktiles_y;
          py0 = ty * cv.kth; px0 = tx * cv.ktw;
        }
        if (A_MODE == 0) {
          tma_load_2d(sa, &tmA, k0, m_blk * T2_BM, full);
        } else if (A_MODE == 1) {
          tma_load_2d(sa, &tmA, m_blk * T2_BM, k0, full);
          tma_load_2d(sa + 8192, &tmA, m_blk * T2_BM + 64, k0, full);
        } else if (A_MODE == 2) {   // K index = tap*C + c, 64 channels of one tap per k-block (C % 64 == 0)
          const int tap = k0 / cv.C, c0 = k0 - tap * cv.C;
          tma_load_4d(sa, &tmA, c0, cx0 + tap % 3 - 1, cy0 + tap / 3 - 1, cb, full);
        } else {                    // dy [pixels, Cout]: 64 pixels x 2 chunks of 64 output channels
          tma_load_4d(sa, &tmA, m_blk * T2_BM, px0, py0, pb, full);
          tma_load_4d(sa + 8192, &tmA, m_blk * T2_BM + 64, px0, py0, pb, full);
        }
        if (B_MODE == 0) {
          tma_load_2d(sb, &tmB, k0, n_blk * BN, full);
        } else if (B_MODE == 1) {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, n_blk * BN + j * 64, k0, full);
        } else {                    // x [pixels, Cin] shifted by the tap of this n-tile: N index = tap*C + c
          const int n0 = n_blk * BN;
          const int tap = n0 / cv.C, c0 = n0 - tap * cv.C;
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(sb + j * 8192, &tmB, c0 + j * 64, px0 + tap % 3 - 1, py0 + tap / 3 - 1, pb, full);
        }
      }
    }
  } else if (warp == 1 && lane == 0 && (CS == 1 || cl_rank == 0)) {
    // ================================ MMA issuer (pair mode: the leader CTA only) ================================
    constexpr uint32_t idesc = umma_idesc(T2_BM * CS, BN, (A_MODE == 1 || A_MODE == 3) ? 1 : 0, B_MODE != 0 ? 1 : 0);
    uint32_t it = 0, tc = 0;
    for (int t = cl_id; t < n_work; t += n_cl, ++tc) {
      T2_DECODE(t)
      const int kb0 = z * kb_per_split, nkb = min(kb_per_split, total_kb - kb0);
      const uint32_t ab = tc & 1;
      mbar_wait(bar_tempty + 8 * ab, ((tc >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ab * BN;
      for (int i = 0; i < nkb; ++i, ++it) {
        const uint32_t s = it % S::STAGES, ph = (it / S::STAGES) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * S::STAGE, sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < T2_BK / 16; ++k) {
          const uint64_t ad = (A_MODE == 1 || A_MODE == 3) ? umma_smem_desc(sa + k * 2048, 8192, 1024) : umma_smem_desc(sa + k * 32, 0, 1024);
          const uint64_t bd = (B_MODE != 0) ? umma_smem_desc(sb + k * 2048, 8192, 1024) : umma_smem_desc(sb + k * 32, 0, 1024);
          if (CS == 1) umma_bf16(tacc, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
          else umma_bf16_2cta(tacc, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        if (CS == 1) umma_commit(bar_empty + 8 * s);
        else umma_commit_2cta(bar_empty + 8 * s, cl_mask);   // frees this stage in BOTH CTAs once the pair's MMAs have read it
      }
      if (CS == 1) umma_commit(bar_tfull + 8 * ab);
      else umma_commit_2cta(bar_tfull + 8 * ab, cl_mask);     // accumulators complete: both CTAs' epilogues may start
    }
  } else if (warp == 3 && lane == 0) {
    // ================================ TMA loader of the epilogue operand ================================
    if (x_in) {
      uint32_t e = 0;
      for (int t = cl_id; t < n_work; t += n_cl) {
        T2_DECODE(t)
        for (int sub = 0; sub < S::NSUB; ++sub, ++e) {
          const uint32_t b = e % S::EPI_BUFS;
          mbar_wait(bar_xempty + 8 * b, ((e / S::EPI_BUFS) & 1) ^ 1);
          mbar_expect_tx(bar_xfull + 8 * b, S::SUB_BYTES);
          tma_load_2d(smem_x + b * S::SUB_BYTES, &tmX, n_blk * BN + sub * S::SUB_COLS, m_blk * T2_BM, bar_xfull + 8 * b);
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (4 T2_NQ warps: lane quarter q, column group hf) ================================
    const int ew = warp - 4, q = ew & 3, hf = ew >> 2, row = q * 32 + lane, etid = threadIdx.x - 128;
    constexpr int CH = S::SUB_COLS / T2_NQ;   // accumulator columns per thread per sub-tile
    constexpr int CPT = 8 / T2_NQ;            // 16-byte chunks of the 128-byte output row per thread
    uint32_t tc = 0, e = 0;
    // One named barrier per sub-tile: thread 0 waits until every TMA store issued so far has finished READING shared
    // memory right before the barrier, so after it the other buffer (last stored one sub-tile ago) may be overwritten
    // while the buffer just filled is handed to the TMA store.  The bias of a tile's BN columns is fetched one tile ahead
    // and parked in shared memory ([2][BN], published by the last barrier of the previous tile), then read as broadcast LDS.
    auto tile_bias = [&](int u) -> float {
      if (!ep.bias || u >= n_work || etid >= BN) return 0.f;
      const int r2 = u % mgn_tiles, nb2 = r2 - (r2 / n_tiles) * n_tiles;
      const int nb = nb2 * BN + etid;
      return nb < N ? __ldg(ep.bias + nb) : 0.f;
    };
    if (ep.bias) {
      if (etid < BN) s_bias[etid] = tile_bias(cl_id);
      named_bar_sync(T2_EPI_BAR, T2_EPI_THREADS);
    }
    for (int t = cl_id; t < n_work; t += n_cl, ++tc) {
      T2_DECODE(t)
      const uint32_t ab = tc & 1;
      const float* sb_cur = s_bias + (tc & 1) * BN;
      const float bias_next = tile_bias(t + n_cl);
      mbar_wait(bar_tfull + 8 * ab, (tc >> 1) & 1);
      tc_fence_after();
      float rs = 1.0f;
      if (ep.rowscale) {
        const int64_t m = (int64_t)m_blk * T2_BM + row;
        if (m < M) rs = __ldg(ep.rowscale + m / ep.rows_per_sample);
      }
#pragma unroll 1
      for (int sub = 0; sub < S::NSUB; ++sub, ++e) {
        const uint32_t b = e % S::EPI_BUFS;
        const uint32_t sx = smem_x + b * S::SUB_BYTES + row * 128, sd = smem_d + b * S::SUB_BYTES + row * 128;
        const int n_sub = n_blk * BN + sub * S::SUB_COLS;
        const int n0 = n_sub + hf * CH;
        if (x_in) mbar_wait(bar_xfull + 8 * b, (e / S::EPI_BUFS) & 1);
        uint32_t v[CH];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * BN + sub * S::SUB_COLS + hf * CH;
        tmem_ld_n_nowait<CH>(taddr, v);
        uint32_t xr[4 * CPT];
        if (x_in) {
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const uint32_t src = sx + (((hf * CPT + c) ^ (row & 7)) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xr[4 * c]), "=r"(xr[4 * c + 1]), "=r"(xr[4 * c + 2]), "=r"(xr[4 * c + 3]) : "r"(src));
          }
        }
        tmem_ld_wait();
        if (sub == S::NSUB - 1) {       // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CS == 1) mbar_arrive(bar_tempty + 8 * ab); else mbar_arrive_cluster(bar_tempty + 8 * ab, 0); }
        }
        float f[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) f[j] = __uint_as_float(v[j]);
        if (ep.bias) {
          const float4* sb4 = reinterpret_cast<const float4*>(sb_cur + sub * S::SUB_COLS + hf * CH);   // zero beyond N
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) {
            const float4 bv = sb4[j];
            f[4 * j] += bv.x; f[4 * j + 1] += bv.y; f[4 * j + 2] += bv.z; f[4 * j + 3] += bv.w;
          }
        }
        if (OUT_F32) {
          if (ep.rowscale) {
#pragma unroll
            for (int j = 0; j < CH; ++j) f[j] *= rs;
          }
          if (ep.x_mode == 1) {            // fp32 residual stream
#pragma unroll
            for (int j = 0; j < CH; ++j) f[j] += __uint_as_float(xr[j]);
          }
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const uint32_t off = (((hf * CPT + c) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sd + off), "r"(__float_as_uint(f[4 * c])), "r"(__float_as_uint(f[4 * c + 1])),
                         "r"(__float_as_uint(f[4 * c + 2])), "r"(__float_as_uint(f[4 * c + 3])) : "memory");
          }
        } else {
          uint32_t pre[CH / 2];
          if (ep.act == 1) {
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) {
              pre[j] = pack2_bf16(f[2 * j], f[2 * j + 1]);
              f[2 * j] = gelu_fast_f(f[2 * j]); f[2 * j + 1] = gelu_fast_f(f[2 * j + 1]);
            }
          } else if (ep.act == 3) {         // the MLP pair of the bf16 training path: the backward's GELU' is computed here, next to
#pragma unroll                             // the sigmoid it shares, and stored instead of the pre-activation
            for (int j = 0; j < CH / 2; ++j) {
              float d0, d1;
              f[2 * j] = gelu_both_fast_f(f[2 * j], d0); f[2 * j + 1] = gelu_both_fast_f(f[2 * j + 1], d1);
              pre[j] = pack2_bf16(d0, d1);
            }
          } else if (ep.act == 4) {
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) {
              const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr[j]));
              f[2 * j] *= h.x; f[2 * j + 1] *= h.y;
            }
          } else if (ep.act == 2) {
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) {
              const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr[j]));
              f[2 * j] *= gelu_grad_fast_f(h.x); f[2 * j + 1] *= gelu_grad_fast_f(h.y);
            }
          }
          if (ep.rowscale) {
#pragma unroll
            for (int j = 0; j < CH; ++j) f[j] *= rs;
          }
          if (ep.x_mode == 1) {
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) {
              const float2 h = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xr[j]));
              f[2 * j] += h.x; f[2 * j + 1] += h.y;
            }
          }
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const uint32_t off = (((hf * CPT + c) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sd + off), "r"(pack2_bf16(f[8 * c], f[8 * c + 1])), "r"(pack2_bf16(f[8 * c + 2], f[8 * c + 3])),
                         "r"(pack2_bf16(f[8 * c + 4], f[8 * c + 5])), "r"(pack2_bf16(f[8 * c + 6], f[8 * c + 7])) : "memory");
            if (ep.act == 1 || ep.act == 3)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sx + off), "r"(pre[4 * c]), "r"(pre[4 * c + 1]), "r"(pre[4 * c + 2]), "r"(pre[4 * c + 3]) : "memory");
          }
          if (ep.colsum) {
            // column sums of the 32 rows of this warp: recursive halving over the lanes while more than one column is
            // held, plain butterfly adds afterwards; lane l ends up with column l >> (5 - log2 CH) of this CH-column chunk
            // (the lowest lane of each group of 32 / CH writes); rows beyond M contribute nothing
            const bool row_ok = (int64_t)m_blk * T2_BM + row < M;
            float cs[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) cs[j] = row_ok ? f[j] : 0.f;   // fp32 values: the exact bias gradient, no F2F round trips
            int held = CH;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              if (held > 1) {
                const int hlf = held >> 1;
                const bool hi = (lane & off) != 0;
#pragma unroll
                for (int j = 0; j < CH / 2; ++j) {
                  if (j < hlf) {
                    const float send = hi ? cs[j] : cs[j + hlf];
                    const float recv = __shfl_xor_sync(0xffffffffu, send, off);
                    cs[j] = (hi ? cs[j + hlf] : cs[j]) + recv;
                  }
                }
                held = hlf;
              } else {
                cs[0] += __shfl_xor_sync(0xffffffffu, cs[0], off);
              }
            }
            constexpr int LPC = 32 / CH;          // lanes per column
            const int col = lane / LPC;
            if ((lane % LPC) == 0 && n0 + col < N) atomicAdd(ep.colsum + n0 + col, cs[0]);
          }
        }
        if (x_in) { __syncwarp(); if (lane == 0) mbar_arrive(bar_xempty + 8 * b); }
        if (sub == S::NSUB - 1 && ep.bias && etid < BN) s_bias[((tc & 1) ^ 1) * BN + etid] = bias_next;
        fence_proxy_async();
        if (etid == 0) bulk_wait_read<0>();
        named_bar_sync(T2_EPI_BAR, T2_EPI_THREADS);
        if (etid == 0) {
          if (A_MODE == 2) {
            int qq = m_blk; const int tx = qq % cv.tiles_x; qq /= cv.tiles_x; const int ty = qq % cv.tiles_y; const int cb = qq / cv.tiles_y;
            tma_store_4d(&tmD, smem_d + b * S::SUB_BYTES, n_sub, tx * cv.tw, ty * cv.th, cb);
          } else if (OUT_F32 && ep.reduce) {
            tma_reduce_add_2d(&tmD, smem_d + b * S::SUB_BYTES, n_sub, m_blk * T2_BM);
          } else {
            tma_store_2d(&tmD, smem_d + b * S::SUB_BYTES, n_sub, m_blk * T2_BM);
            if (ep.act == 1 || ep.act == 3) tma_store_2d(&tmX, smem_x + b * S::SUB_BYTES, n_sub, m_blk * T2_BM);
          }
          bulk_commit();
        }
      }
    }
    if (etid == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();          // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == 2) { tc_fence_after(); if (CS == 2) tmem_dealloc_2cta(tmem_base, S::TMEM_COLS); else tmem_dealloc(tmem_base, S::TMEM_COLS); }
#undef T2_DECODE
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static int make_map_2d_f32(CUtensorMap* m, const void* p, int64_t inner, int64_t outer, int64_t ld_elems, int box_inner,
                           int box_outer) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return MTUS_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(p) & 15) || (ld_elems % 4)) return MTUS_ERR_BAD_ARG;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(p), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MTUS_OK : MTUS_ERR_DRIVER;
}

// Measured on B200 at the Swin-B stage-3 shapes (M = 6272; profiles/r2_gemm_cta_pair_ab.txt): the CTA-pair engine is parity-
// clean but SLOWER than single-CTA tiles (qkv fwd 13.2 -> 18.9 us, fc1 21.5 -> 25.6 us): 25 m-pairs x n-tiles quantise worse on
// 74 pairs than 49 x n-tiles on 148 CTAs (qkv: 3 waves instead of 2), and at 1-3 tiles per CTA these kernels are bound by the
// epilogue and by fill / drain latency, not by the operand stream the pairing halves.  Kept for large-M GEMMs, off by default.
static bool t2_pair_default() { return false; }

static int g_sm_count = 0;
static int sm_count() {
  if (!g_sm_count) {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_sm_count = n > 0 ? n : 148;
    const char* e = getenv("MTUS_GEMM_SMS");       // cap of the persistent grid (leaves SMs to a concurrent NCCL kernel)
    if (e && atoi(e) > 0 && atoi(e) < g_sm_count) g_sm_count = atoi(e);
  }
  return g_sm_count;
}

template <int BN, int A_MODE, int B_MODE, bool OUT_F32, int CS>
static int t2_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tx, const CUtensorMap& td, int M,
                     int N, int K, int kbps, int total_kb, int m_tiles, int n_tiles, int splits, const T2Conv& cv,
                     const T2Epi& ep, cudaStream_t st) {
  auto kern = gemm_tc2_kernel<BN, A_MODE, B_MODE, OUT_F32, CS>;
  static mtus_per_device_flag configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T2Smem<BN, OUT_F32, CS>::TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured.set();
  }
  const int64_t units64 = (int64_t)splits * ceil_div(m_tiles, CS) * n_tiles;
  if (units64 > (1ll << 30)) return MTUS_ERR_UNSUPPORTED;
  const int n_units = (int)units64;
  const int max_cl = sm_count() / CS;
  const int n_cl = n_units < max_cl ? n_units : max_cl;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_cl * CS); cfg.blockDim = dim3(T2_THREADS);
  cfg.dynamicSmemBytes = T2Smem<BN, OUT_F32, CS>::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (CS == 1 && mtus_pdl_enabled()) ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tx, td, M, N, K, kbps, total_kb, m_tiles, n_tiles, n_units, cv, ep);
  if (e != cudaSuccess) return (int)e;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// Tile width: wider tiles move fewer operand bytes from L2 per flop (the main-loop limiter of these GEMMs: about
// 42 B/clk/SM when every SM pulls), narrower ones quantise better on the 148 SMs.  Model: waves x (k-blocks x L2 time of
// one k-block + a fixed per-tile cost), all in cycles.  128x192 is what makes N = 512 at M = 6272 (49 x 3 = 147 tiles) a
// single full wave.
static int t2_pick_bn(int M, int N, int K, bool allow_192_256) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("MTUS_BN"); forced = e ? atoi(e) : 0; }
  if (N <= 64) return 64;
  if (!allow_192_256 || N <= 128) return 128;
  if (forced == 128 || forced == 192 || forced == 256) return forced;
  const int sms = sm_count();
  const int64_t mt = ceil_div(M, T2_BM), kb = ceil_div(K, T2_BK);
  int best = 128;
  double best_t = 0;
  for (int bn = 128; bn <= 256; bn += 64) {
    const int64_t tiles = mt * ceil_div(N, bn);
    const double waves = (double)((tiles + sms - 1) / sms);
    const double t = waves * ((double)kb * (16384.0 + 128.0 * bn) / 42.5 + 600.0 + 3.0 * bn);
    if (bn == 128 || t < best_t) { best = bn; best_t = t; }
  }
  return best;
}

bool mtus_gemm_tc2_supported(const mtus_gemm_desc* d) {
  if (d->dtype != MTUS_BF16) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(d->a) || !al16(d->b) || !al16(d->out)) return false;
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return false;
  const bool wgrad_conv = d->b_conv && d->a_mn_major && d->b_mn_major;
  if (d->a_conv && d->b_conv) return false;
  if (d->b_conv && !wgrad_conv) return false;
  if (d->a_conv && (d->a_mn_major || d->b_mn_major)) return false;
  if ((d->a_conv || d->b_conv) && (d->conv_c % 64)) return false;
  if (d->a_mn_major && !d->b_mn_major) return false;
  if (!d->a_conv && !wgrad_conv && (d->lda % 8)) return false;
  if (!d->b_conv && (d->ldb % 8)) return false;
  if (d->out_f32) {
    // fp32 output: either accumulate (TMA reduce-add, split-K weight gradients) or a plain store with an optional
    // bias / drop-path scale / fp32 residual (the fp32 residual stream of the bf16 mode)
    if (d->act) return false;
    if (d->ld_out % 4) return false;
    if (d->atomic && (d->bias || d->res || d->rowscale)) return false;
    if (d->res && (!d->res_f32 || d->res_mode != 1 || d->ld_res % 4 || !al16(d->res))) return false;
    if (d->a_conv || d->out_colsum) return false;
    if (wgrad_conv && (!d->atomic || d->lda % 64 || d->conv_c % 64 || d->M % 64)) return false;   // an n-tile must lie inside one tap: BN | Cin
  } else {
    if (d->atomic || d->res_f32) return false;
    if (d->ld_out % 8) return false;
    if (d->res && (d->res_mode != 1 || d->ld_res % 8 || !al16(d->res))) return false;
    if (d->act && (!d->aux || d->ld_aux % 8 || !al16(d->aux))) return false;
    if (d->act && d->res) return false;             // one epilogue operand tile
    if (d->act < 0 || d->act > 4) return false;
    if (d->a_conv && (d->bias || d->act || d->res || d->rowscale)) return false;
    if (d->a_conv && (d->N % 64 || d->ld_out != d->N)) return false;
    if (wgrad_conv) return false;
  }
  return true;
}

int mtus_gemm_tc2(const mtus_gemm_desc* d, cudaStream_t st) {
  if (!mtus_gemm_tc2_supported(d)) return MTUS_ERR_UNSUPPORTED;
  const int M = d->M, N = d->N, K = d->K;
  const bool wgrad_conv = d->b_conv != 0;
  CUtensorMap ta, tb, tx, td;
  T2Conv cv{};
  int rc;
  // wide tiles: plain Linear shapes (K-major A; fp32 output = residual stream only with K-major B) and the Linear weight
  // gradients (both operands MN-major, fp32 reduce-add output): a 128x256 tile pulls 48 KB per k-block for twice the MACs of
  // a 128x128 tile's 32 KB -- the weight gradients are L2->SM operand-bound (K = all tokens), so this is +33 % flop per byte
  const bool lin_wgrad = !d->a_conv && !d->b_conv && d->a_mn_major && d->b_mn_major && d->out_f32 && d->atomic;
  const bool wide_ok = (!d->a_conv && !d->b_conv && !d->a_mn_major && (!d->out_f32 || !d->b_mn_major)) || lin_wgrad;
  int BN = t2_pick_bn(M, N, K, wide_ok && !lin_wgrad);
  if (lin_wgrad) {
    static int wg_bn = -1;
    if (wg_bn < 0) { const char* e = getenv("MTUS_WGRAD_BN"); wg_bn = e ? atoi(e) : 128; }   // measured: 128x256 weight-gradient tiles are 5-20 % SLOWER (fewer, longer work items)
    BN = (N <= 64) ? 64 : ((wg_bn == 256 && N % 256 == 0) ? 256 : 128);
  }
  if (d->a_conv) {
    // implicit-GEMM conv forward / data gradient: 256-wide tiles when the output channels allow (the data gradient of the
    // segmentation head's Conv3x3(512 -> 128) has N = 512); MTUS_CONV_BN=128 restores the narrower tile
    static int cv_bn = -1;
    if (cv_bn < 0) { const char* e = getenv("MTUS_CONV_BN"); cv_bn = e ? atoi(e) : 256; }
    if (cv_bn == 256 && N % 256 == 0) BN = 256;
  }
  if (wgrad_conv) {
    // conv weight gradient: N = 9 Cin with the tap in the high part of the index, so the tile width must divide Cin
    // (64 for Cin % 128 != 0; 256 where Cin % 256 == 0: 48 KB per k-block for twice the MACs of the 128-wide tile's 32 KB --
    // the segmentation step gains 80 us; MTUS_CONV_WGRAD_BN=128 restores the narrower tile)
    static int cw_bn = -1;
    if (cw_bn < 0) { const char* e = getenv("MTUS_CONV_WGRAD_BN"); cw_bn = e ? atoi(e) : 256; }
    BN = (d->conv_c % 128) ? 64 : ((cw_bn == 256 && d->conv_c % 256 == 0) ? 256 : (cw_bn == 64 ? 64 : 128));
  }
  int m_tiles = ceil_div(M, T2_BM), n_tiles = ceil_div(N, BN), total_kb = ceil_div(K, T2_BK);
  if (d->a_conv || d->b_conv) {
    cv.H = d->conv_h; cv.W = d->conv_w; cv.C = d->conv_c;
    cv.tw = (cv.W >= 16) ? 16 : 8; cv.th = 128 / cv.tw;
    cv.tiles_x = ceil_div(cv.W, cv.tw); cv.tiles_y = ceil_div(cv.H, cv.th);
    cv.ktw = 8; cv.kth = 8;
    cv.ktiles_x = ceil_div(cv.W, 8); cv.ktiles_y = ceil_div(cv.H, 8);
  }
  int CS = 1;
  {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("MTUS_CLUSTER"); forced = e ? atoi(e) : 0; }
    // CTA pairs (cta_group::2) for the plain Linear shapes with 256-wide tiles: MTUS_CLUSTER=1 disables, =2 forces where legal
    const bool pair_ok = !d->a_conv && !d->b_conv && !d->a_mn_major && BN == 256 && m_tiles >= 2;
    CS = (forced == 1) ? 1 : ((pair_ok && (forced == 2 || t2_pair_default())) ? 2 : 1);
  }
  int am, bm;
  if (d->a_conv) {
    const int B = M / (cv.H * cv.W);
    rc = make_map_conv(&ta, d->a, B, cv.H, cv.W, cv.C, cv.tw, cv.th);
    m_tiles = B * cv.tiles_x * cv.tiles_y;
    am = 2;
  } else if (wgrad_conv) {           // A = dy [B,H,W,Cout] pixel-major, Cout = M = lda
    const int B = K / (cv.H * cv.W);
    rc = make_map_conv(&ta, d->a, B, cv.H, cv.W, (int)d->lda, cv.ktw, cv.kth);
    total_kb = B * cv.ktiles_x * cv.ktiles_y;
    am = 3;
  } else if (!d->a_mn_major) { rc = make_map_2d(&ta, d->a, K, M, d->lda, T2_BK, T2_BM); am = 0; }
  else { rc = make_map_2d(&ta, d->a, M, K, d->lda, 64, T2_BK); am = 1; }
  if (rc) return rc;
  if (wgrad_conv) {
    const int B = K / (cv.H * cv.W);
    rc = make_map_conv(&tb, d->b, B, cv.H, cv.W, cv.C, cv.ktw, cv.kth);
    bm = 2;
  } else if (!d->b_mn_major) { rc = make_map_2d(&tb, d->b, K, N, d->ldb, T2_BK, BN / CS); bm = 0; }
  else { rc = make_map_2d(&tb, d->b, N, K, d->ldb, 64, T2_BK); bm = 1; }
  if (rc) return rc;
  // D / X maps
  T2Epi ep{};
  ep.bias = d->bias; ep.rowscale = d->rowscale; ep.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  ep.act = d->act;
  ep.x_mode = d->res ? 1 : ((d->act == 2 || d->act == 4) ? 2 : ((d->act == 1 || d->act == 3) ? 3 : 0));
  ep.reduce = d->atomic ? 1 : 0;
  ep.colsum = d->out_colsum;
  if (d->out_f32) rc = make_map_2d_f32(&td, d->out, N, M, d->ld_out, 32, T2_BM);
  else if (d->a_conv) rc = make_map_conv(&td, d->out, M / (cv.H * cv.W), cv.H, cv.W, N, cv.tw, cv.th);
  else rc = make_map_2d(&td, d->out, N, M, d->ld_out, 64, T2_BM);
  if (rc) return rc;
  if (ep.x_mode == 1 && d->out_f32) rc = make_map_2d_f32(&tx, d->res, N, M, d->ld_res, 32, T2_BM);
  else if (ep.x_mode == 1) rc = make_map_2d(&tx, d->res, N, M, d->ld_res, 64, T2_BM);
  else if (ep.x_mode) rc = make_map_2d(&tx, d->aux, N, M, d->ld_aux, 64, T2_BM);
  else tx = td;
  if (rc) return rc;

  int splits = d->split_k > 0 ? d->split_k : 1;
  if (!d->atomic) splits = 1;
  if (lin_wgrad && d->split_k > 0 && BN != 128) {
    // the caller sized split_k for 128x128 tiles; keep the same number of work items (~4 waves) with this tile width
    const int64_t tiles = (int64_t)m_tiles * n_tiles;
    int64_t want = (4ll * sm_count() + tiles - 1) / tiles;
    const int64_t maxs = total_kb / 4 > 0 ? total_kb / 4 : 1;
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    splits = (int)want;
  }
  if (wgrad_conv && d->atomic && BN != 128) {       // the caller sized split_k for 128-wide tiles: keep the number of work items
    const int64_t want = BN == 256 ? 2ll * splits : (splits + 1) / 2;
    splits = (int)(want < 1 ? 1 : want);
  }
  if (splits > total_kb) splits = total_kb;
  const int kbps = ceil_div(total_kb, splits);
  splits = ceil_div(total_kb, kbps);

#define T2_GO2(BN_, AM_, BM_, F32_)                                                                                                  \
  {                                                                                                                                  \
    if constexpr (BN_ == 256 && AM_ == 0 && BM_ != 2) {                                                                              \
      if (CS == 2) return t2_launch<BN_, AM_, BM_, F32_, 2>(ta, tb, tx, td, M, N, K, kbps, total_kb, m_tiles, n_tiles, splits, cv, ep, st); \
    }                                                                                                                                \
    return t2_launch<BN_, AM_, BM_, F32_, 1>(ta, tb, tx, td, M, N, K, kbps, total_kb, m_tiles, n_tiles, splits, cv, ep, st);              \
  }
#define T2_GO(BN_, AM_, BM_, F32_) T2_GO2(BN_, AM_, BM_, F32_)
  if (d->out_f32) {
    if (BN == 256) {
      if (am == 0 && bm == 0) T2_GO(256, 0, 0, true);
      if (am == 1 && bm == 1) T2_GO(256, 1, 1, true);
      if (am == 3 && bm == 2) T2_GO(256, 3, 2, true);
    } else if (BN == 192) {
      if (am == 0 && bm == 0) T2_GO(192, 0, 0, true);
    } else if (BN == 128) {
      if (am == 0 && bm == 0) T2_GO(128, 0, 0, true);
      if (am == 1 && bm == 1) T2_GO(128, 1, 1, true);
      if (am == 3 && bm == 2) T2_GO(128, 3, 2, true);
    } else {
      if (am == 0 && bm == 0) T2_GO(64, 0, 0, true);
      if (am == 1 && bm == 1) T2_GO(64, 1, 1, true);
      if (am == 3 && bm == 2) T2_GO(64, 3, 2, true);
    }
  } else if (BN == 256) {
    if (am == 0 && bm == 0) T2_GO(256, 0, 0, false);
    if (am == 0 && bm == 1) T2_GO(256, 0, 1, false);
    if (am == 2 && bm == 0) T2_GO(256, 2, 0, false);
  } else if (BN == 192) {
    if (am == 0 && bm == 0) T2_GO(192, 0, 0, false);
    if (am == 0 && bm == 1) T2_GO(192, 0, 1, false);
  } else if (BN == 128) {
    if (am == 0 && bm == 0) T2_GO(128, 0, 0, false);
    if (am == 0 && bm == 1) T2_GO(128, 0, 1, false);
    if (am == 2 && bm == 0) T2_GO(128, 2, 0, false);
  } else {
    if (am == 0 && bm == 0) T2_GO(64, 0, 0, false);
    if (am == 0 && bm == 1) T2_GO(64, 0, 1, false);
    if (am == 2 && bm == 0) T2_GO(64, 2, 0, false);
  }
#undef T2_GO2
#undef T2_GO
  return MTUS_ERR_UNSUPPORTED;
}
