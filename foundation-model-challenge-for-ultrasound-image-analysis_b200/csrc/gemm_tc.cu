// bf16 GEMM on the 5th-gen tensor cores: tcgen05.mma (cta_group::1, kind::f16) with the fp32
// accumulator in TMEM, operands staged in shared memory by TMA (SWIZZLE_128B), warp-specialised
// (TMA producer / MMA issuer / 4 epilogue warps), mbarrier full/empty ring, fused epilogue.
//
// Replaces: cuBLAS calls behind the nn.Linear modules of timm WindowAttention/Mlp/PatchMerging and
// the smp FPN convolutions in the bf16 mode.
//
// One 128 x BN output tile per CTA; 2 CTAs are resident per SM (3-stage ring, 96 KB smem, 128 TMEM
// columns each) so one CTA's epilogue overlaps the other's main loop -- K is short in Swin
// (128..4096), the epilogue (bias/GELU/residual, 1-2 stores) is not.
//
// Operand forms (both bf16):  K-major  : stored [rows][k], k contiguous   (TMA box 64k x 128 rows)
//                             MN-major : stored [k][rows], rows contiguous (TMA boxes 64 rows x 64 k)
//   forward  y = x w^T   : A K-major (x),  B K-major (w)
//   dgrad    dx = dy w   : A K-major (dy), B MN-major (w)
//   wgrad    dw = dy^T x : A MN-major (dy), B MN-major (x), split-K over the token dimension
//   conv3x3 fwd / dgrad  : A = implicit im2col of an NHWC map via 4-D TMA boxes (OOB zero fill = padding)
//   conv3x3 wgrad        : B = the same 4-D boxes used MN-major
#include "common.cuh"
#include "tc_ptx.cuh"

#define TC_BM 128
#define TC_BK 64
#define TC_STAGES 3
#define TC_THREADS 256


// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
struct TcConv {  // implicit-im2col geometry of the conv operand (NHWC [B,H,W,C])
  int H, W, C;
  int th, tw;    // pixel tile: th rows x tw cols = 128 pixels (tw in {8,16}); tile index -> (b, ty, tx)
  int tiles_x, tiles_y;
};

template <int BN>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int TOTAL = TC_STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
};

// A_MODE / B_MODE: 0 K-major 2-D, 1 MN-major 2-D, 2 conv (4-D boxes; K-major for A, MN-major for B)
template <typename OutT_unused, int BN, int A_MODE, int B_MODE>
__global__ void __launch_bounds__(TC_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, int M, int N,
                                                             int K, int kb_per_split, TcConv cv, EpiParams ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using S = TcSmem<BN>;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * S::STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + TC_STAGES),
                 bar_tmem = smem_u32(bars + 2 * TC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.y, n_blk = blockIdx.x;
  const int total_kb = (K + TC_BK - 1) / TC_BK;
  const int kb0 = blockIdx.z * kb_per_split;
  const int nkb = min(kb_per_split, total_kb - kb0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tmem, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    // conv operand: this CTA's 128 "rows" are a th x tw pixel tile of one image
    int cb = 0, cy0 = 0, cx0 = 0;
    if (A_MODE == 2) {
      int t = m_blk; const int tx = t % cv.tiles_x; t /= cv.tiles_x; const int ty = t % cv.tiles_y; cb = t / cv.tiles_y;
      cy0 = ty * cv.th; cx0 = tx * cv.tw;
    }
    for (int i = 0; i < nkb; ++i) {
      const int s = i % TC_STAGES, ph = (i / TC_STAGES) & 1;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      const uint32_t full = bar_full + 8 * s;
      const uint32_t sa = smem_base + s * S::STAGE, sb = sa + S::A_BYTES;
      mbar_expect_tx(full, S::STAGE);
      const int k0 = (kb0 + i) * TC_BK;
      if (A_MODE == 0) {
        tma_load_2d(sa, &tmA, k0, m_blk * TC_BM, full);
      } else if (A_MODE == 1) {
        tma_load_2d(sa, &tmA, m_blk * TC_BM, k0, full);
        tma_load_2d(sa + 8192, &tmA, m_blk * TC_BM + 64, k0, full);
      } else {  // conv, K index = tap*C + c, 64 channels of one tap per k-block (C % 64 == 0)
        const int tap = k0 / cv.C, c0 = k0 - tap * cv.C;
        tma_load_4d(sa, &tmA, c0, cx0 + tap % 3 - 1, cy0 + tap / 3 - 1, cb, full);
      }
      if (B_MODE == 0) {
        tma_load_2d(sb, &tmB, k0, n_blk * BN, full);
      } else if (B_MODE == 1) {
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, n_blk * BN + j * 64, k0, full);
      } else {  // conv wgrad: N index = tap*C + c (BN channels of one tap), K index = pixel (one image-row-tile per k-block)
        const int n0 = n_blk * BN;
        const int tap = n0 / cv.C, c0 = n0 - tap * cv.C;
        int t = kb0 + i; const int tx = t % cv.tiles_x; t /= cv.tiles_x; const int ty = t % cv.tiles_y; const int b = t / cv.tiles_y;
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_4d(sb + j * 8192, &tmB, c0 + j * 64, tx * cv.tw + tap % 3 - 1, ty * cv.th + tap / 3 - 1, b, full);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = umma_idesc(TC_BM, BN, A_MODE == 1 ? 1 : 0, B_MODE != 0 ? 1 : 0);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % TC_STAGES, ph = (i / TC_STAGES) & 1;
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      const uint32_t sa = smem_base + s * S::STAGE, sb = sa + S::A_BYTES;
#pragma unroll
      for (int k = 0; k < TC_BK / 16; ++k) {
        // K-major: 8-row groups 1024 B apart, +32 B per UMMA_K inside the 128-B swizzle atom.
        // MN-major: 64-element MN chunks 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO), +2048 B per UMMA_K.
        const uint64_t ad = (A_MODE == 1) ? umma_smem_desc(sa + k * 2048, 8192, 1024) : umma_smem_desc(sa + k * 32, 0, 1024);
        const uint64_t bd = (B_MODE != 0) ? umma_smem_desc(sb + k * 2048, 8192, 1024) : umma_smem_desc(sb + k * 32, 0, 1024);
        umma_bf16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(bar_empty + 8 * s);   // frees the smem slot once these MMAs have read it
    }
    umma_commit(bar_tmem);              // accumulator complete
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> fused epilogue -> global =====
    const int q = warp - 4;             // TMEM lane quarter this warp may touch
    mbar_wait(bar_tmem, 0);
    tc_fence_after();
    const int row_in_tile = q * 32 + lane;
    int64_t m;
    bool row_ok;
    if (A_MODE == 2) {
      int t = m_blk; const int tx = t % cv.tiles_x; t /= cv.tiles_x; const int ty = t % cv.tiles_y; const int b = t / cv.tiles_y;
      const int y = ty * cv.th + row_in_tile / cv.tw, x = tx * cv.tw + row_in_tile % cv.tw;
      row_ok = (y < cv.H) && (x < cv.W);
      m = ((int64_t)b * cv.H + y) * cv.W + x;
    } else {
      m = (int64_t)m_blk * TC_BM + row_in_tile;
      row_ok = m < M;
    }
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
      if (row_ok && nkb > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = n_blk * BN + c * 32 + j * 4;
          const int nvalid = min(4, N - n);
          if (nvalid > 0) {
            float v[4] = {__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3])};
            epilogue4<bf16>(ep, m, n, nvalid, v);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, BN); }
}


template <int BN, int A_MODE, int B_MODE>
static int tc_launch(const CUtensorMap& ta, const CUtensorMap& tb, dim3 grid, int M, int N, int K, int kbps, TcConv cv,
                     const EpiParams& ep, cudaStream_t st) {
  auto kern = gemm_tc_kernel<bf16, BN, A_MODE, B_MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::TOTAL);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  kern<<<grid, TC_THREADS, TcSmem<BN>::TOTAL, st>>>(ta, tb, M, N, K, kbps, cv, ep);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

static void conv_tiling(int H, int W, TcConv& cv) {
  cv.tw = (W >= 16) ? 16 : 8;
  cv.th = 128 / cv.tw;
  cv.tiles_x = (W + cv.tw - 1) / cv.tw;
  cv.tiles_y = (H + cv.th - 1) / cv.th;
}

bool mtus_gemm_tc_supported(const mtus_gemm_desc* d) {
  if (d->dtype != MTUS_BF16) return false;
  if (d->a_conv && d->b_conv) return false;
  if ((reinterpret_cast<uintptr_t>(d->a) & 15) || (reinterpret_cast<uintptr_t>(d->b) & 15)) return false;
  if (d->a_conv || d->b_conv) {
    if (d->conv_c % 64) return false;
    if (d->a_conv && d->a_mn_major) return false;
    if (d->b_conv) return false;   // conv wgrad on tcgen05 needs 64-pixel boxes for both operands: not wired yet
  }
  if (!d->a_conv && (d->lda % 8)) return false;
  if (!d->b_conv && (d->ldb % 8)) return false;
  if (d->N % 4) return false;
  return true;
}

int mtus_gemm_tc(const mtus_gemm_desc* d, const EpiParams& ep, cudaStream_t st) {
  if (!mtus_gemm_tc_supported(d)) return MTUS_ERR_UNSUPPORTED;
  CUtensorMap ta, tb;
  TcConv cv{};
  int rc;
  const int M = d->M, N = d->N, K = d->K;
  const int BN = (N % 128 == 0 || N > 192) ? 128 : 64;
  int m_tiles = ceil_div(M, TC_BM);
  int total_kb = ceil_div(K, TC_BK);
  if (d->a_conv || d->b_conv) {
    cv.H = d->conv_h; cv.W = d->conv_w; cv.C = d->conv_c;
    conv_tiling(cv.H, cv.W, cv);
  }
  // A
  if (d->a_conv) {
    const int B = M / (cv.H * cv.W);
    rc = make_map_conv(&ta, d->a, B, cv.H, cv.W, cv.C, cv.tw, cv.th);
    m_tiles = B * cv.tiles_x * cv.tiles_y;
  } else if (!d->a_mn_major) rc = make_map_2d(&ta, d->a, K, M, d->lda, TC_BK, TC_BM);
  else rc = make_map_2d(&ta, d->a, M, K, d->lda, 64, TC_BK);
  if (rc) return rc;
  // B
  if (d->b_conv) {
    const int B = K / (cv.H * cv.W);
    rc = make_map_conv(&tb, d->b, B, cv.H, cv.W, cv.C, cv.tw, cv.th);
    total_kb = B * cv.tiles_x * cv.tiles_y;   // one pixel tile (128 pixels = 2 x 64-row k-blocks) ... see below
  } else if (!d->b_mn_major) rc = make_map_2d(&tb, d->b, K, N, d->ldb, TC_BK, BN);
  else rc = make_map_2d(&tb, d->b, N, K, d->ldb, 64, TC_BK);
  if (rc) return rc;
  if (d->b_conv) return MTUS_ERR_UNSUPPORTED;  // conv wgrad on tcgen05: enabled once the 64-pixel box variant lands

  int splits = d->split_k > 0 ? d->split_k : 1;
  if (splits > total_kb) splits = total_kb;
  int kbps = ceil_div(total_kb, splits);
  splits = ceil_div(total_kb, kbps);
  dim3 grid(ceil_div(N, BN), m_tiles, splits);
  if (grid.y > 65535) { /* swap to x-major M for very tall problems */ return MTUS_ERR_UNSUPPORTED; }

#define TC_GO(BN_, AM_, BM_) return tc_launch<BN_, AM_, BM_>(ta, tb, grid, M, N, K, kbps, cv, ep, st)
  const int am = d->a_conv ? 2 : (d->a_mn_major ? 1 : 0);
  const int bm = d->b_mn_major ? 1 : 0;
  if (BN == 128) {
    if (am == 0 && bm == 0) TC_GO(128, 0, 0);
    if (am == 0 && bm == 1) TC_GO(128, 0, 1);
    if (am == 1 && bm == 1) TC_GO(128, 1, 1);
    if (am == 2 && bm == 0) TC_GO(128, 2, 0);
  } else {
    if (am == 0 && bm == 0) TC_GO(64, 0, 0);
    if (am == 0 && bm == 1) TC_GO(64, 0, 1);
    if (am == 1 && bm == 1) TC_GO(64, 1, 1);
    if (am == 2 && bm == 0) TC_GO(64, 2, 0);
  }
#undef TC_GO
  return MTUS_ERR_UNSUPPORTED;
}
