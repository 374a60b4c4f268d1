// Shifted-window attention FORWARD on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), windows of up to 64 tokens.
//
// Same contract as attention_mma.cu / attention.cu (timm _attn + WindowAttention minus the two Linear layers; SURVEY
// section 8a rows a5, a6): cyclic shift, window partition / reverse, relative-position bias, shift mask, softmax and P@V
// happen inside the kernel; nothing window-shaped touches HBM.  What changes is the engine of the two contractions.
//
// Work unit = (two windows, two heads).  Packing is what makes the 49x32x49 per-(window, head) products fill a tcgen05 tile:
//   * rows: two windows of <= 64 tokens stacked -> M = 128 (one TMEM lane per query row, one thread per row);
//   * channels: the q / k / v slices of two adjacent heads are 64 contiguous bf16 = one 128-byte SWIZZLE_128B row, so a tile
//     [128 tokens][64 channels] is exactly the K-major operand tile of the GEMM engine; head h of the pair is selected by
//     advancing the descriptor start address by 64 bytes (k-steps 2h, 2h + 1 of the swizzle atom).
//   S_h [128 x 128] = Q_h K_h^T   (2 tcgen05.mma of K = 16 per head; only the two 64 x 64 diagonal blocks are used: the padded
//                                  MMA issues (128*128) / (2*49*49) = 3.4x the attention-only flops, which are 2 % of a block)
//   P_h = softmax rows (fp32, exp2), written back INTO TENSOR MEMORY as bf16 (tcgen05.st) over the columns S_h occupied, with
//         the off-diagonal 64 x 64 blocks zero: P never exists in shared memory and needs no proxy fence
//   O_h [128 x 64] = P_h V       (8 tcgen05.mma of K = 16 with the A operand read from TENSOR MEMORY, V tile [128 tokens]
//                                  [64 channels] as the MN-major B operand; the 32 columns of head h are read back)
//
// Version 2 (round 2) is a PIPELINE instead of the serial chain of version 1 (which was 3x slower than the mma.sync engine,
// profiles/r2_attention_tc_vs_mma.txt): warp 0 = TMA producer running up to three units ahead through a 3-stage q/k/v ring,
// warp 1 = MMA issuer, warps 4-7 and 8-11 = two softmax groups that own alternate units and alternate halves of TMEM.  The
// issuer interleaves S(u) with O(u - 1), so while one group runs the softmax of unit u - 1 the tensor pipe computes S of unit
// u and the loads of units u + 1, u + 2 are in flight.  TMEM (512 columns) = 2 slots x 2 heads x 128 columns:
// [S_h | S_h] -> [P_h (64 columns of packed bf16) | O_h (64 fp32 columns)].
//
// Loads: one TMA box per window ROW (ww tokens x 128 bytes of the NHWC qkv tensor viewed as (3C, W, H, B)), landing in
// rows of the operand tiles; the cyclic shift is coordinate arithmetic on the box origin, and a window row that wraps
// around the right edge is two narrower boxes (tensor maps with box widths ww, ww - shift, shift).
//
// STATUS (measured on B200, profiles/r2_attention_tc_vs_mma.txt): parity-clean with both load paths.  Version 2 is 1.8x faster than
// version 1 (stage 1: 113 -> 62 us, stage 3: 48 -> 23 us) and still slower than the mma.sync engine (37 / 14 us), so it stays OPT-IN
// (MTUS_ATTN_TC=1).  Cycle stamps of one CTA: a softmax group needs ~12500 cycles per unit (S -> P 4500-6500: 2 heads x 64 scores
// per thread at 2 resident warps per scheduler; P -> O 1300-2000; O read-back + row stores 2000-3300; window / address arithmetic
// 1200-1800; waiting for the next S 1500), i.e. ~6500 per unit with both groups, against ~5000 per two windows x two heads for the
// mma.sync kernel, whose 16 warps per SM hide every latency.  At stage 3 (4 units per CTA) the 144 KB zero fill, the 32 KB bias
// expansion and the TMEM allocation per CTA are not amortised.  What it would take: per-row constants hoisted out of the unit
// loop, a staged (coalesced) output store, a softmax that needs fewer issue slots -- or the fusion with the qkv / proj GEMMs.
//
// Eligible: bf16, window <= 64 tokens, map divisible by the window (no padded tokens), even head count.  Everything else
// (512x512 padded maps, window 12, fp32 mode) stays on attention_mma*.cu / attention.cu.  Backward: attention_mma.cu.
#include "common.cuh"
#include "internal.h"
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <string.h>

#define AT_THREADS 384                       // warp 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4-7 / 8-11 softmax groups
#define AT_STAGES 3
#define AT_TILE 16384                       // [128 rows][64 bf16] SWIZZLE_128B
#define AT_LOG2E 1.4426950408889634f

struct AtGeom {
  int B, H, W, C, heads, wh, ww, sh, sw, nwy, nwx, N, ntab, lin_stride, lin_off;
  int windows, pairs, head_pairs, groups;
  float scale2;
};

// shared memory map (bytes from the 1024-aligned base)
#define AT_RING_OFF 0                                   // [3 stages][q, k, v][16384]
#define AT_BIAS_OFF (AT_STAGES * 3 * AT_TILE)           // [2 heads][64 keys][64 queries] fp32, log2 domain
#define AT_MISC_OFF (AT_BIAS_OFF + 2 * 64 * 64 * 4)     // s_lin[64] | barriers | tmem slot | bias-table scratch [2][ntab <= 225]
#define AT_SMEM_BYTES (AT_MISC_OFF + 4096 + 1024)

__device__ __forceinline__ uint32_t at_sw128(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
__device__ __forceinline__ float at_lds(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
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

// tcgen05.mma with the A operand in tensor memory (cute SM100_MMA_F16BF16_TS): D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void at_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void at_tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void at_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <bool USE_TMA>
__global__ void __launch_bounds__(AT_THREADS, 1)
window_attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmFull, const __grid_constant__ CUtensorMap tmTop,
                          const __grid_constant__ CUtensorMap tmBot, const __grid_constant__ CUtensorMap tmLeft,
                          const __grid_constant__ CUtensorMap tmRight, const bf16* __restrict__ qkv, const float* __restrict__ table,
                          bf16* __restrict__ out, float* __restrict__ lse, AtGeom g) {
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* s_bias = reinterpret_cast<float*>(smem + AT_BIAS_OFF);
  int* s_lin = reinterpret_cast<int*>(smem + AT_MISC_OFF);                  // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_MISC_OFF + 256);
  // barriers: full[3] empty[3] s_ready[2] p_ready[2] o_ready[2] slot_free[2]
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * AT_STAGES, bar_s = bar_empty + 8 * AT_STAGES, bar_p = bar_s + 16,
                 bar_o = bar_p + 16, bar_free = bar_o + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + AT_MISC_OFF + 256 + 128);
  float* s_tab = reinterpret_cast<float*>(smem + AT_MISC_OFF + 512);        // [2][ntab]
  const int hp = blockIdx.x % g.head_pairs, grp = blockIdx.x / g.head_pairs;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmFull) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmTop) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBot) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmLeft) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmRight) : "memory");
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(bar_full + 8 * i, USE_TMA ? 1 : 96); mbar_init(bar_empty + 8 * i, 1); }
    for (int j = 0; j < 2; ++j) {
      mbar_init(bar_s + 8 * j, 1); mbar_init(bar_p + 8 * j, 4); mbar_init(bar_o + 8 * j, 1); mbar_init(bar_free + 8 * j, 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  // operand tiles start as zeros: the rows beyond the window (49..63 of each 64-row slot) are never written afterwards, and
  // they must be finite for the MMAs that sweep over them
  for (int i = tid; i < (AT_BIAS_OFF) / 16; i += AT_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int t = tid; t < 64; t += AT_THREADS) s_lin[t] = (t < g.N) ? (t / g.ww) * g.lin_stride + (t % g.ww) : 0;
  for (int e = tid; e < 2 * g.ntab; e += AT_THREADS) s_tab[e] = __ldg(table + (e % g.ntab) * g.heads + hp * 2 + e / g.ntab) * AT_LOG2E;
  __syncthreads();
  // relative-position bias of the two heads, transposed ([key][query]: a warp's 32 query rows read consecutive words)
  for (int e = tid; e < 2 * 64 * 64; e += AT_THREADS) {
    const int q = e & 63, k = (e >> 6) & 63, h = e >> 12;
    float v = -INFINITY;
    if (k < g.N) v = (q < g.N) ? s_tab[h * g.ntab + s_lin[q] - s_lin[k] + g.lin_off] : 0.f;
    s_bias[e] = v;
  }
  fence_proxy_async();                                  // the zero fill is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int n_units = (g.pairs - grp + g.groups - 1) / g.groups;        // this CTA's window pairs: grp, grp + groups, ...
  const int c0 = hp * 64;                                               // first channel of the head pair inside q / k / v
  constexpr uint32_t idesc_s = umma_idesc(128, 128, 0, 0);
  constexpr uint32_t idesc_o = umma_idesc(128, 64, 0, 1);

  if (!USE_TMA && (warp == 0 || warp == 2 || warp == 3)) {
    // ================================ gather loaders (96 lanes, the default): one valid tile row per lane ================================
    // 16-byte cp.async gathers of q, k and v (24 per row), unit u into ring stage u % 3; a unit's arrival on the stage's barrier
    // follows one unit behind its issue (cp.async.wait_group 1 -> fence.proxy.async -> arrive), so two units of loads are in
    // flight per lane.  Measured against the TMA form below: window-shaped 4-D boxes (7 x 7 tokens x 128 bytes, or one box per
    // window row) cost the TMA unit ~16 cycles per 128-byte row, 5000-9000 cycles per unit, which made the producer the
    // bottleneck of the pipeline (profiles/r2_attention_tc_vs_mma.txt).
    const int li = (warp == 0 ? 0 : warp - 1) * 32 + lane;               // 0 .. 95
    auto arrive_unit = [&](int v) { fence_proxy_async(); mbar_arrive(bar_full + 8 * (v % AT_STAGES)); };
    for (int u = 0; u < n_units; ++u) {
      const int st = u % AT_STAGES;
      const int wp = grp + u * g.groups;
      const uint32_t ring = sbase + AT_RING_OFF + st * 3 * AT_TILE;
      mbar_wait(bar_empty + 8 * st, ((u / AT_STAGES) & 1) ^ 1);
      for (int rv = li; rv < 2 * g.N; rv += 96) {
        const int sl = rv >= g.N ? 1 : 0, tt = rv - sl * g.N;
        const AtWin w = at_window(g, 2 * wp + sl);
        if (!w.valid) continue;
        const int ty = tt / g.ww, tx = tt - ty * g.ww;
        int y = w.wy * g.wh + ty + g.sh, x = w.wx * g.ww + tx + g.sw;
        if (y >= g.H) y -= g.H;
        if (x >= g.W) x -= g.W;
        const bf16* p = qkv + ((int64_t)(w.b * g.H + y) * g.W + x) * (3 * g.C) + c0;
        const int r = sl * 64 + tt;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t off = at_sw128(r, ch);
          at_cp_async16(ring + off, p + ch * 8);
          at_cp_async16(ring + AT_TILE + off, p + g.C + ch * 8);
          at_cp_async16(ring + 2 * AT_TILE + off, p + 2 * g.C + ch * 8);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (u >= 1) { asm volatile("cp.async.wait_group 1;" ::: "memory"); arrive_unit(u - 1); }
    }
    if (n_units >= 1) { asm volatile("cp.async.wait_group 0;" ::: "memory"); arrive_unit(n_units - 1); }
  } else if (USE_TMA && warp == 0) {
    // ================================ TMA producer (all 32 lanes issue; MTUS_ATTN_TC_LOADS=tma): unit u -> ring stage u % 3 ================================
    // A window that does not wrap around the right edge arrives as ONE box per tensor (ww x wh tokens x 128 bytes; two boxes when
    // the cyclic shift wraps it around the bottom edge); only the last window column of a shifted map needs one pair of narrower
    // boxes per window row.  A single thread issuing 42 one-row boxes per unit was the bottleneck of the first pipelined build
    // (5000-9000 cycles per unit in the cycle-stamp trace): the boxes are now fewer and spread over the lanes.
    const int ops_per = 2 * g.wh;                                       // slots per (window, tensor): enough for the per-row form
    for (int u = 0; u < n_units; ++u) {
      const int st = u % AT_STAGES;
      const int wp = grp + u * g.groups;
      const uint32_t ring = sbase + AT_RING_OFF + st * 3 * AT_TILE;
      const AtWin w0 = at_window(g, 2 * wp), w1 = at_window(g, 2 * wp + 1);
      const uint32_t full = bar_full + 8 * st;
      if (lane == 0) {
        mbar_wait(bar_empty + 8 * st, ((u / AT_STAGES) & 1) ^ 1);
        mbar_expect_tx(full, (uint32_t)((w0.valid + w1.valid) * g.N * 128 * 3));
      }
      __syncwarp();
      for (int op = lane; op < 2 * 3 * ops_per; op += 32) {
        const int idx = op % ops_per, which = (op / ops_per) % 3, sl = op / (3 * ops_per);
        const AtWin& w = sl ? w1 : w0;
        if (!w.valid) continue;
        int x0 = w.wx * g.ww + g.sw, y0 = w.wy * g.wh + g.sh;
        if (x0 >= g.W) x0 -= g.W;
        if (y0 >= g.H) y0 -= g.H;
        const bool wrapx = x0 + g.ww > g.W, wrapy = y0 + g.wh > g.H;    // last window column / row of a shifted map
        const uint32_t dst = ring + which * AT_TILE + (uint32_t)(sl * 64 * 128);
        const int cc = which * g.C + c0;
        if (!wrapx) {
          if (!wrapy) { if (idx == 0) tma_load_4d(dst, &tmFull, cc, x0, y0, w.b, full); }
          else if (idx == 0) tma_load_4d(dst, &tmTop, cc, x0, y0, w.b, full);                                       // wh - sh rows up to the edge
          else if (idx == 1) tma_load_4d(dst + (uint32_t)((g.wh - g.sh) * g.ww * 128), &tmBot, cc, x0, 0, w.b, full);  // sh rows from y = 0
        } else {
          const int ty = idx >> 1;
          int y = y0 + ty;
          if (y >= g.H) y -= g.H;
          const uint32_t drow = dst + (uint32_t)(ty * g.ww * 128);
          if ((idx & 1) == 0) tma_load_4d(drow, &tmLeft, cc, x0, y, w.b, full);                               // ww - sw tokens up to the edge
          else tma_load_4d(drow + (uint32_t)((g.ww - g.sw) * 128), &tmRight, cc, 0, y, w.b, full);            // sw tokens from x = 0
        }
      }
      __syncwarp();
    }
  } else if (warp == 1 && lane == 0) {
    // ================================ MMA issuer: S(u), then O(u - 1) ================================
    auto issue_o = [&](int v) {
      const int jv = v & 1, kv = v >> 1, stv = v % AT_STAGES;
      mbar_wait(bar_p + 8 * jv, kv & 1);                    // the slot's softmax group has written P_0, P_1 into tensor memory
      tc_fence_after();
      const uint32_t vt = sbase + AT_RING_OFF + stv * 3 * AT_TILE + 2 * AT_TILE;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t bd = umma_smem_desc(vt + ks * 2048, 8192, 1024);
          at_umma_ts(tmem_base + jv * 256 + h * 128 + 64, tmem_base + jv * 256 + h * 128 + ks * 8, bd, idesc_o, ks);
        }
      umma_commit(bar_o + 8 * jv);
      umma_commit(bar_empty + 8 * stv);                     // q / k / v of this unit are consumed: the stage may be refilled
    };
    for (int u = 0; u < n_units; ++u) {
      const int st = u % AT_STAGES, j = u & 1, k = u >> 1;
      mbar_wait(bar_full + 8 * st, (u / AT_STAGES) & 1);
      mbar_wait(bar_free + 8 * j, (k & 1) ^ 1);             // the group has read O of the slot's previous unit
      tc_fence_after();
      const uint32_t ring = sbase + AT_RING_OFF + st * 3 * AT_TILE;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint64_t ad = umma_smem_desc(ring + h * 64 + kk * 32, 0, 1024);
          const uint64_t bd = umma_smem_desc(ring + AT_TILE + h * 64 + kk * 32, 0, 1024);
          umma_bf16(tmem_base + j * 256 + h * 128, ad, bd, idesc_s, kk);
        }
      umma_commit(bar_s + 8 * j);
      if (u >= 1) issue_o(u - 1);
    }
    if (n_units >= 1) issue_o(n_units - 1);
  } else if (warp >= 4) {
    // ================================ softmax group j: units j, j + 2, ... in TMEM slot j ================================
    const int j = (warp - 4) >> 2, row = ((warp - 4) & 3) * 32 + lane;    // TMEM lane quarter = warp % 4
    const int slot = row >> 6, t = row & 63;                              // this thread's query row: window slot, token
    const uint32_t lane_addr = tmem_base + ((uint32_t)(((warp - 4) & 3) * 32) << 16) + j * 256;
    for (int u = j; u < n_units; u += 2) {
      const int k = u >> 1;
      const int wp = grp + u * g.groups;
      const AtWin win = at_window(g, 2 * wp + slot);
      const bool row_ok = win.valid && t < g.N;
      const bool masked = (g.sh > 0 && win.wy == g.nwy - 1) || (g.sw > 0 && win.wx == g.nwx - 1);   // uniform per warp
      // shift mask of this row as 64 bits (bit c: key c lies in another region): built in a rolled loop, used as predicates
      uint32_t mlo = 0u, mhi = 0u;
      if (masked) {
        const int my_reg = at_region(g, win, t < g.N ? t : 0);
#pragma unroll 1
        for (int c = 0; c < g.N; ++c) {
          const uint32_t d = at_region(g, win, c) != my_reg ? 1u : 0u;
          if (c < 32) mlo |= d << c; else mhi |= d << (c - 32);
        }
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
      mbar_wait(bar_s + 8 * j, k & 1);
      tc_fence_after();
      // ---- softmax of this thread's row, both heads; P (un-normalised, bf16) -> tensor memory, over S ----------------------
      float inv_l[2], lse2[2];
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        uint32_t v[64];
        const uint32_t taddr = lane_addr + h * 128;
        tmem_ld32_nowait(taddr + slot * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));       // the diagonal block of this row's window
        tmem_ld32_nowait(taddr + slot * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
        const uint32_t bq = sbase + AT_BIAS_OFF + (uint32_t)(h * 4096 + t) * 4;     // [key][query]: stride 256 bytes per key
        float mx = -INFINITY;
        if (masked) {                       // warp-uniform: only the last window row / column of a shifted map pays for the mask
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            float sc = fmaf(__uint_as_float(v[c]), g.scale2, at_lds(bq + c * 256));           // -inf beyond the window
            if (((c < 32 ? mlo : mhi) >> (c & 31)) & 1u) sc += -100.0f * AT_LOG2E;
            v[c] = __float_as_uint(sc);
            mx = fmaxf(mx, sc);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            const float sc = fmaf(__uint_as_float(v[c]), g.scale2, at_lds(bq + c * 256));
            v[c] = __float_as_uint(sc);
            mx = fmaxf(mx, sc);
          }
        }
        float sum = 0.f;
        uint32_t pk[32], zz[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float e0 = at_ex2(__uint_as_float(v[2 * c]) - mx), e1 = at_ex2(__uint_as_float(v[2 * c + 1]) - mx);
          sum += e0 + e1;
          pk[c] = pack2_bf16(e0, e1);
          zz[c] = 0u;
        }
        // P_h [128 x 128 keys] bf16 = 64 packed columns: this row's own window in its half, zeros in the other window's half
        at_tmem_st32(taddr + slot * 32, pk);
        at_tmem_st32(taddr + (slot ^ 1) * 32, zz);
        inv_l[h] = 1.0f / sum; lse2[h] = mx + log2f(sum);
      }
      at_tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + 8 * j);
      mbar_wait(bar_o + 8 * j, k & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld32_nowait(lane_addr + 64, o0);                  // head 0: columns 0..31 of O_0
      tmem_ld32_nowait(lane_addr + 128 + 64 + 32, o1);       // head 1: columns 32..63 of O_1
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free + 8 * j);          // O is in registers: the slot may take the S of its next unit
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
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
  CUtensorMap tf, tt, tb, tl, tr;
  int rc = make_map_conv(&tf, qkv, B, H, W, 3 * C, ww, wh);
  if (rc) return rc;
  tt = tf; tb = tf; tl = tf; tr = tf;
  if (sh > 0) {
    rc = make_map_conv(&tt, qkv, B, H, W, 3 * C, ww, wh - sh);
    if (rc) return rc;
    rc = make_map_conv(&tb, qkv, B, H, W, 3 * C, ww, sh);
    if (rc) return rc;
  }
  if (sw > 0) {
    rc = make_map_conv(&tl, qkv, B, H, W, 3 * C, ww - sw, 1);
    if (rc) return rc;
    rc = make_map_conv(&tr, qkv, B, H, W, 3 * C, sw, 1);
    if (rc) return rc;
  }
  static int use_tma = -1;
  if (use_tma < 0) { const char* e = getenv("MTUS_ATTN_TC_LOADS"); use_tma = (e && !strcmp(e, "tma")) ? 1 : 0; }
  static mtus_per_device_flag configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attn_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured.set();
  }
  const dim3 grid(g.groups * g.head_pairs), block(AT_THREADS);
  cudaError_t le;
  if (use_tma) le = mtus_launch_pdl(window_attn_tc_fwd_kernel<true>, grid, block, (size_t)AT_SMEM_BYTES, st, tf, tt, tb, tl, tr, (const bf16*)qkv, rel_table, (bf16*)out, lse, g);
  else le = mtus_launch_pdl(window_attn_tc_fwd_kernel<false>, grid, block, (size_t)AT_SMEM_BYTES, st, tf, tt, tb, tl, tr, (const bf16*)qkv, rel_table, (bf16*)out, lse, g);
  if (le != cudaSuccess) return (int)le;
  MTUS_LAUNCH_STATUS();
  ++g_at_launches;
  return MTUS_OK;
}

extern "C" int64_t mtus_window_attn_tc_launch_count(void) { return g_at_launches; }
