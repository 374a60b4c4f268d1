// Shifted-window attention FORWARD on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), windows of up to 64 tokens.
//
// Same contract as attention_mma.cu / attention.cu (timm _attn + WindowAttention minus the two Linear layers; SURVEY
// section 8a rows a5, a6): cyclic shift, window partition / reverse, relative-position bias, shift mask, softmax and P@V
// happen inside the kernel; nothing window-shaped touches HBM.  What changes is the engine of the two contractions.
//
// Work unit = (two windows, two heads).  Packing is what makes the 49x32x49 per-(window, head) products fill a tcgen05
// tile:
//   * rows: two windows of <= 64 tokens stacked -> M = 128 (one TMEM lane per query row, one thread per row);
//   * channels: the q / k / v slices of two adjacent heads are 64 contiguous bf16 = one 128-byte SWIZZLE_128B row, so a tile
//     [128 tokens][64 channels] is exactly the K-major operand tile of the GEMM engine; head h of the pair is selected by
//     advancing the descriptor start address by 64 bytes (k-steps 2h, 2h + 1 of the swizzle atom).
//   S_h [128 x 128] = Q_h K_h^T   (2 tcgen05.mma of K = 16 per head; only the two 64 x 64 diagonal blocks are used: the padded
//                                  MMA issues (128*128) / (2*49*49) = 3.4x the attention-only flops, which are 2 % of a block)
//   P_h = softmax rows (fp32, exp2), written as bf16 into a K-major [128][128] tile whose off-diagonal 64 x 64 blocks stay 0
//   O_h [128 x 64] = P_h V       (8 tcgen05.mma of K = 16, V tile [128 tokens][64 channels] as the MN-major B operand;
//                                  the 32 columns of head h are read back, the other 32 are discarded)
// Accumulators live in TMEM (S: 2 x 128 columns, O: 2 x 64 columns) and are read with tcgen05.ld by the thread that owns
// the row, so the softmax needs no shuffles at all.
//
// Loads: one TMA box per window ROW (ww tokens x 128 bytes of the NHWC qkv tensor viewed as (3C, W, H, B)), landing in
// rows of the operand tiles; the cyclic shift is coordinate arithmetic on the box origin, and a window row that wraps
// around the right edge is two narrower boxes (tensor maps with box widths ww, ww - shift, shift).  A 2-stage ring keeps
// the loads of unit i+1 in flight while unit i computes.  (MTUS_ATTN_TC_LOADS=cpasync selects per-thread 16-byte gathers.)
//
// Eligible: bf16, window <= 64 tokens, map divisible by the window (no padded tokens), even head count.  Everything else
// (512x512 padded maps, window 12, fp32 mode) stays on attention_mma*.cu / attention.cu.  Backward: attention_mma.cu.
//
// STATUS (measured on B200, profiles/r2_attention_tc_vs_mma.txt): parity-clean with both load paths, but 3x SLOWER than the
// mma.sync engine at the Swin-B shapes (stage 1: 115 us vs 37.5 us; stage 3: 51 us vs 14.2 us), so it is OPT-IN
// (MTUS_ATTN_TC=1) and the product path keeps attention_mma.cu.  Why: one CTA of 4 softmax warps per SM runs the unit as a
// serial chain (TMA wait -> S MMAs -> commit / wait -> 128-element softmax per thread, one warp per scheduler -> P to shared
// memory -> O MMAs -> commit / wait -> store) with two exposed tensor-core round trips per unit, while the mma.sync engine
// keeps 16 independent warps per SM busy with no TMEM round trip.  Making tcgen05 win here needs the FA4-style pipeline
// (dedicated MMA / load warp, the two heads of a unit ping-ponging between softmax and MMA, 8 softmax warps) or, better,
// the fusion with the qkv / proj GEMMs that removes the HBM round trip that bounds both engines (what comes next).
#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <string.h>

#define AT_THREADS 128
#define AT_TILE 16384                       // [128 rows][64 bf16] SWIZZLE_128B
#define AT_LOG2E 1.4426950408889634f

struct AtGeom {
  int B, H, W, C, heads, wh, ww, sh, sw, nwy, nwx, N, ntab, lin_stride, lin_off;
  int windows, pairs, head_pairs, groups;
  float scale2;
};

// shared memory map (bytes from the 1024-aligned base)
#define AT_RING_OFF 0                       // [2 stages][q, k, v][16384]
#define AT_P_OFF (2 * 3 * AT_TILE)          // [2 heads][2 k-atoms][16384]
#define AT_BIAS_OFF (AT_P_OFF + 4 * AT_TILE)  // [2 heads][64 keys][64 queries] fp32, log2 domain
#define AT_MISC_OFF (AT_BIAS_OFF + 2 * 64 * 64 * 4)
#define AT_SMEM_BYTES (AT_MISC_OFF + 1024 + 1024)

__device__ __forceinline__ uint32_t at_sw128(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
__device__ __forceinline__ float at_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void at_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

struct AtWin { int b, wy, wx, valid; };
__device__ __forceinline__ AtWin at_window(const AtGeom& g, int w) {
  AtWin r;
  r.valid = w < g.windows;
  if (!r.valid) w = g.windows - 1;
  r.wx = w % g.nwx; w /= g.nwx; r.wy = w % g.nwy; r.b = w / g.nwy;
  return r;
}
// shift-mask region of token t of window win (0 unless the window sits in the last window row / column of a shifted map)
__device__ __forceinline__ int at_region(const AtGeom& g, const AtWin& win, int t) {
  const int ty = t / g.ww, tx = t - ty * g.ww;
  int r = 0;
  if (g.sh > 0 && win.wy == g.nwy - 1) r += (ty < g.wh - g.sh) ? 3 : 6;
  if (g.sw > 0 && win.wx == g.nwx - 1) r += (tx < g.ww - g.sw) ? 1 : 2;
  return r;
}

template <bool USE_TMA>
__global__ void __launch_bounds__(AT_THREADS, 1)
window_attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmFull, const __grid_constant__ CUtensorMap tmLeft,
                          const __grid_constant__ CUtensorMap tmRight, const bf16* __restrict__ qkv, const float* __restrict__ table,
                          bf16* __restrict__ out, float* __restrict__ lse, AtGeom g) {
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* s_bias = reinterpret_cast<float*>(smem + AT_BIAS_OFF);
  int* s_lin = reinterpret_cast<int*>(smem + AT_MISC_OFF);                  // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_MISC_OFF + 512);   // full[2], s_done, o_done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + AT_MISC_OFF + 512 + 64);
  const uint32_t bar_full = smem_u32(bars), bar_s = bar_full + 16, bar_o = bar_full + 24;
  const int hp = blockIdx.x % g.head_pairs, grp = blockIdx.x / g.head_pairs;

  if (USE_TMA && tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmFull) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmLeft) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmRight) : "memory");
  }
  if (tid == 0) {
    mbar_init(bar_full, 1); mbar_init(bar_full + 8, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  // operand tiles start as zeros: the rows beyond the window (49..63 of each 64-row slot) and the off-diagonal blocks of P are
  // never written afterwards, and they must be finite for the MMAs that sweep over them
  for (int i = tid; i < (AT_BIAS_OFF) / 16; i += AT_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int t = tid; t < 64; t += AT_THREADS) s_lin[t] = (t < g.N) ? (t / g.ww) * g.lin_stride + (t % g.ww) : 0;
  __syncthreads();
  // relative-position bias of the two heads, transposed ([key][query]: a warp's 32 query rows read consecutive words)
  float* s_tab = reinterpret_cast<float*>(smem + AT_P_OFF);                 // scratch (P is not in use yet): [2][ntab]
  for (int e = tid; e < 2 * g.ntab; e += AT_THREADS) s_tab[e] = __ldg(table + (e % g.ntab) * g.heads + hp * 2 + e / g.ntab) * AT_LOG2E;
  __syncthreads();
  for (int e = tid; e < 2 * 64 * 64; e += AT_THREADS) {
    const int q = e & 63, k = (e >> 6) & 63, h = e >> 12;
    float v = -INFINITY;
    if (k < g.N) v = (q < g.N) ? s_tab[h * g.ntab + s_lin[q] - s_lin[k] + g.lin_off] : 0.f;
    s_bias[e] = v;
  }
  __syncthreads();
  for (int e = tid; e < 2 * g.ntab; e += AT_THREADS) s_tab[e] = 0.f;       // P's off-diagonal blocks must be zero again
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int n_units = (g.pairs - grp + g.groups - 1) / g.groups;        // this CTA's window pairs: grp, grp + groups, ...
  const int slot = tid >> 6, t = tid & 63;                              // this thread's query row: window slot, token
  const int c0 = hp * 64;                                               // first channel of the head pair inside q / k / v

  auto issue_loads = [&](int unit, int st) {
    const int wp = grp + unit * g.groups;
    const uint32_t ring = sbase + AT_RING_OFF + st * 3 * AT_TILE;
    if (USE_TMA) {
      if (tid != 0) return;
      const AtWin w0 = at_window(g, 2 * wp), w1 = at_window(g, 2 * wp + 1);
      const uint32_t full = bar_full + 8 * st;
      mbar_expect_tx(full, (uint32_t)((w0.valid + w1.valid) * g.N * 128 * 3));
      for (int s = 0; s < 2; ++s) {
        const AtWin& w = s ? w1 : w0;
        if (!w.valid) continue;
        int x0 = w.wx * g.ww + g.sw;
        if (x0 >= g.W) x0 -= g.W;
        const bool wrap = x0 + g.ww > g.W;                              // only the last window column of a shifted map
        for (int ty = 0; ty < g.wh; ++ty) {
          int y = w.wy * g.wh + ty + g.sh;
          if (y >= g.H) y -= g.H;
          const uint32_t row0 = (uint32_t)((s * 64 + ty * g.ww) * 128);
#pragma unroll
          for (int which = 0; which < 3; ++which) {
            const uint32_t dst = ring + which * AT_TILE + row0;
            const int cc = which * g.C + c0;
            if (!wrap) tma_load_4d(dst, &tmFull, cc, x0, y, w.b, full);
            else {
              tma_load_4d(dst, &tmLeft, cc, x0, y, w.b, full);                                   // ww - sw tokens up to the edge
              tma_load_4d(dst + (uint32_t)((g.ww - g.sw) * 128), &tmRight, cc, 0, y, w.b, full);  // sw tokens from x = 0
            }
          }
        }
      }
    } else {
      // per-thread gathers: thread (row, chunk) moves 16 bytes of q, k and v; rows beyond the window are never touched
      for (int i = tid; i < 128 * 8; i += AT_THREADS) {
        const int r = i >> 3, ch = i & 7, s = r >> 6, tt = r & 63;
        const AtWin w = at_window(g, 2 * wp + s);
        if (!w.valid || tt >= g.N) continue;
        const int ty = tt / g.ww, tx = tt - ty * g.ww;
        int y = w.wy * g.wh + ty + g.sh, x = w.wx * g.ww + tx + g.sw;
        if (y >= g.H) y -= g.H;
        if (x >= g.W) x -= g.W;
        const bf16* p = qkv + ((int64_t)(w.b * g.H + y) * g.W + x) * (3 * g.C) + c0 + ch * 8;
        const uint32_t off = at_sw128(r, ch);
        at_cp_async16(ring + off, p);
        at_cp_async16(ring + AT_TILE + off, p + g.C);
        at_cp_async16(ring + 2 * AT_TILE + off, p + 2 * g.C);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };

  constexpr uint32_t idesc_s = umma_idesc(128, 128, 0, 0);
  constexpr uint32_t idesc_o = umma_idesc(128, 64, 0, 1);
  if (n_units > 0) issue_loads(0, 0);

  for (int u = 0; u < n_units; ++u) {
    const int st = u & 1;
    const int wp = grp + u * g.groups;
    const uint32_t ring = sbase + AT_RING_OFF + st * 3 * AT_TILE;
    if (u + 1 < n_units) issue_loads(u + 1, st ^ 1);      // stage st^1 was last read by the MMAs of unit u-1 (completed: bar_o)
    if (USE_TMA) {
      mbar_wait(bar_full + 8 * st, (u >> 1) & 1);
    } else {
      if (u + 1 < n_units) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async();
      __syncthreads();
    }
    // ---- S_h = Q_h K_h^T -------------------------------------------------------------------------------------------
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t ad = umma_smem_desc(ring + h * 64 + k * 32, 0, 1024);
          const uint64_t bd = umma_smem_desc(ring + AT_TILE + h * 64 + k * 32, 0, 1024);
          umma_bf16(tmem_base + h * 128, ad, bd, idesc_s, k);
        }
      umma_commit(bar_s);
    }
    const AtWin win = at_window(g, 2 * wp + slot);
    const bool row_ok = win.valid && t < g.N;
    const bool masked = (g.sh > 0 && win.wy == g.nwy - 1) || (g.sw > 0 && win.wx == g.nwx - 1);   // uniform per warp
    // shift mask of this row as 64 bits (bit c: key c lies in another region): built in a rolled loop, used as predicates --
    // evaluating the region arithmetic inside the unrolled softmax blew the kernel up to 150 KB of SASS (instruction-cache bound)
    uint32_t mlo = 0u, mhi = 0u;
    if (masked) {
      const int my_reg = at_region(g, win, t < g.N ? t : 0);
#pragma unroll 1
      for (int c = 0; c < g.N; ++c) {
        const uint32_t d = at_region(g, win, c) != my_reg ? 1u : 0u;
        if (c < 32) mlo |= d << c; else mhi |= d << (c - 32);
      }
    }
    mbar_wait(bar_s, u & 1);
    tc_fence_after();
    // ---- softmax of this thread's row, both heads; P (un-normalised, bf16) -> shared memory -------------------------------
    float inv_l[2], lse2[2];
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      uint32_t v[64];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + h * 128 + slot * 64;
      tmem_ld32_nowait(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32_nowait(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      const float* bq = s_bias + h * 4096 + t;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float s = fmaf(__uint_as_float(v[c]), g.scale2, bq[c * 64]);           // -inf beyond the window
        if (((c < 32 ? mlo : mhi) >> (c & 31)) & 1u) s += -100.0f * AT_LOG2E;
        v[c] = __float_as_uint(s);
        mx = fmaxf(mx, s);
      }
      float sum = 0.f;
      const uint32_t prow = sbase + AT_P_OFF + h * 2 * AT_TILE + slot * AT_TILE;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float e0 = at_ex2(__uint_as_float(v[ch * 8 + 2 * j]) - mx), e1 = at_ex2(__uint_as_float(v[ch * 8 + 2 * j + 1]) - mx);
          sum += e0 + e1;
          pk[j] = pack2_bf16(e0, e1);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + at_sw128(tid, ch)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
      }
      if (h == 0) { inv_l[0] = 1.0f / sum; lse2[0] = mx + log2f(sum); }
      else { inv_l[1] = 1.0f / sum; lse2[1] = mx + log2f(sum); }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- O_h = P_h V -----------------------------------------------------------------------------------------------------
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = umma_smem_desc(sbase + AT_P_OFF + h * 2 * AT_TILE + (ks >> 2) * AT_TILE + (ks & 3) * 32, 0, 1024);
          const uint64_t bd = umma_smem_desc(ring + 2 * AT_TILE + ks * 2048, 8192, 1024);
          umma_bf16(tmem_base + 256 + h * 64, ad, bd, idesc_o, ks);
        }
      umma_commit(bar_o);
    }
    // output row of this thread (cyclic shift undone)
    int64_t src = 0;
    if (row_ok) {
      const int ty = t / g.ww, tx = t - ty * g.ww;
      int y = win.wy * g.wh + ty + g.sh, x = win.wx * g.ww + tx + g.sw;
      if (y >= g.H) y -= g.H;
      if (x >= g.W) x -= g.W;
      src = (int64_t)(win.b * g.H + y) * g.W + x;
    }
    mbar_wait(bar_o, u & 1);
    tc_fence_after();
    {
      uint32_t o0[32], o1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + 256;
      tmem_ld32_nowait(taddr, o0);                    // head 0: columns 0..31 of O_0
      tmem_ld32_nowait(taddr + 64 + 32, o1);          // head 1: columns 32..63 of O_1
      tmem_ld_wait();
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(out + src * g.C + c0);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 a, b;
          a.x = pack2_bf16(__uint_as_float(o0[ch * 8 + 0]) * inv_l[0], __uint_as_float(o0[ch * 8 + 1]) * inv_l[0]);
          a.y = pack2_bf16(__uint_as_float(o0[ch * 8 + 2]) * inv_l[0], __uint_as_float(o0[ch * 8 + 3]) * inv_l[0]);
          a.z = pack2_bf16(__uint_as_float(o0[ch * 8 + 4]) * inv_l[0], __uint_as_float(o0[ch * 8 + 5]) * inv_l[0]);
          a.w = pack2_bf16(__uint_as_float(o0[ch * 8 + 6]) * inv_l[0], __uint_as_float(o0[ch * 8 + 7]) * inv_l[0]);
          b.x = pack2_bf16(__uint_as_float(o1[ch * 8 + 0]) * inv_l[1], __uint_as_float(o1[ch * 8 + 1]) * inv_l[1]);
          b.y = pack2_bf16(__uint_as_float(o1[ch * 8 + 2]) * inv_l[1], __uint_as_float(o1[ch * 8 + 3]) * inv_l[1]);
          b.z = pack2_bf16(__uint_as_float(o1[ch * 8 + 4]) * inv_l[1], __uint_as_float(o1[ch * 8 + 5]) * inv_l[1]);
          b.w = pack2_bf16(__uint_as_float(o1[ch * 8 + 6]) * inv_l[1], __uint_as_float(o1[ch * 8 + 7]) * inv_l[1]);
          dst[ch] = a;
          dst[4 + ch] = b;
        }
        if (lse) {
          lse[src * g.heads + hp * 2] = lse2[0];
          lse[src * g.heads + hp * 2 + 1] = lse2[1];
        }
      }
    }
    tc_fence_before();
    __syncthreads();          // everyone has read S / O and P may be overwritten: the next unit's MMAs may start
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static int64_t g_at_launches = 0;

static int at_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n = v > 0 ? v : 148;
  }
  return n;
}

bool mtus_window_attn_tc_eligible(int B, int H, int W, int C, int heads, int wh, int ww, int sh, int sw, int dtype) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MTUS_ATTN_TC"); enabled = (e && atoi(e) != 0) ? 1 : 0; }   // opt-in, see STATUS above
  if (!enabled || dtype != MTUS_BF16) return false;
  if (B <= 0 || wh <= 0 || ww <= 0 || wh * ww > 64 || heads <= 0 || (heads & 1) || C != heads * 32) return false;
  if (H % wh || W % ww) return false;                   // padded tokens (qkv bias rows): attention_mma.cu
  if (sh < 0 || sw < 0 || sh >= wh || sw >= ww) return false;
  if ((int64_t)B * H * W * 3 * C >= (1ll << 31)) return false;
  return true;
}

int mtus_window_attn_tc_fwd(const void* qkv, const float* rel_table, void* out, float* lse, int B, int H, int W, int C, int heads,
                            int wh, int ww, int sh, int sw, cudaStream_t st) {
  AtGeom g;
  g.B = B; g.H = H; g.W = W; g.C = C; g.heads = heads; g.wh = wh; g.ww = ww; g.sh = sh; g.sw = sw;
  g.nwy = H / wh; g.nwx = W / ww; g.N = wh * ww; g.ntab = (2 * wh - 1) * (2 * ww - 1);
  g.lin_stride = 2 * ww - 1; g.lin_off = (wh - 1) * (2 * ww - 1) + (ww - 1);
  g.scale2 = (1.0f / sqrtf(32.0f)) * AT_LOG2E;
  g.windows = B * g.nwy * g.nwx; g.pairs = (g.windows + 1) / 2; g.head_pairs = heads / 2;
  int G = at_sm_count() / g.head_pairs;
  if (G < 1) G = 1;
  if (G > g.pairs) G = g.pairs;
  const int per = (g.pairs + G - 1) / G;
  g.groups = (g.pairs + per - 1) / per;
  static int use_tma = -1;
  if (use_tma < 0) { const char* e = getenv("MTUS_ATTN_TC_LOADS"); use_tma = (e && !strcmp(e, "cpasync")) ? 0 : 1; }
  CUtensorMap tf, tl, tr;
  memset(&tf, 0, sizeof(tf)); tl = tf; tr = tf;
  if (use_tma) {
    int rc = make_map_conv(&tf, qkv, B, H, W, 3 * C, ww, 1);
    if (rc) return rc;
    tl = tf; tr = tf;
    if (sw > 0) {
      rc = make_map_conv(&tl, qkv, B, H, W, 3 * C, ww - sw, 1);
      if (rc) return rc;
      rc = make_map_conv(&tr, qkv, B, H, W, 3 * C, sw, 1);
      if (rc) return rc;
    }
  }
  static mtus_per_device_flag configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attn_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured.set();
  }
  const dim3 grid(g.groups * g.head_pairs), block(AT_THREADS);
  cudaError_t le;
  if (use_tma) le = mtus_launch_pdl(window_attn_tc_fwd_kernel<true>, grid, block, (size_t)AT_SMEM_BYTES, st, tf, tl, tr, (const bf16*)qkv, rel_table, (bf16*)out, lse, g);
  else le = mtus_launch_pdl(window_attn_tc_fwd_kernel<false>, grid, block, (size_t)AT_SMEM_BYTES, st, tf, tl, tr, (const bf16*)qkv, rel_table, (bf16*)out, lse, g);
  if (le != cudaSuccess) return (int)le;
  MTUS_LAUNCH_STATUS();
  ++g_at_launches;
  return MTUS_OK;
}

extern "C" int64_t mtus_window_attn_tc_launch_count(void) { return g_at_launches; }
