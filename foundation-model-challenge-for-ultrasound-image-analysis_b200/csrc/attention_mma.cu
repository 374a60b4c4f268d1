// Fused shifted-window attention on tensor cores, bf16 storage / fp32 softmax (windows of up to 64 tokens).
//
// Same contract as attention.cu (timm _attn + WindowAttention minus the two Linear layers; SURVEY section 8a rows
// a5, a6): the cyclic shift, zero padding, window partition / reverse, relative-position bias, shift mask,
// softmax and P@V happen inside the kernel and nothing window-shaped touches HBM.  Differences:
//   * one WARP owns one (window, head) item and is fully autonomous (no block-level barrier): q/k/v (and dO)
//     tiles [64 x 32] bf16 are staged in swizzled shared memory with 16-byte cp.async, every contraction runs on
//     the tensor cores (mma.sync m16n8k16, fp32 accumulate) in 16-token strips, the softmax works on the
//     accumulator fragments (exp2 with log2e folded into the scale and the bias table), and outputs leave
//     through a 1 KB staging strip as 64-byte row segments (full sectors);
//   * forward also writes the per-(token, head) log-sum-exp, so the backward pass needs no max / sum reductions:
//     it walks KEY strips once -- S^T and dP^T strips come straight out of the MMAs in the layout that dV and dK
//     consume as A operands; only dS is transposed (2 KB strip through shared memory + ldmatrix.trans) to
//     accumulate dQ in registers;
//   * the relative-position-bias gradient is binned in shared memory over all the windows a warp processes and
//     flushed with one atomicAdd per bin; the qkv-bias gradient (column sums of dqkv) is accumulated in registers
//     and flushed once per warp, so no separate column-sum pass over dqkv is needed.
// A stand-alone attention kernel is HBM-bound (24.5 FLOP/B at 49 tokens): algorithmic bytes per token are
// 4*C*2 forward (q, k, v in; o out) and 8*C*2 backward (q, k, v, o, dO in; dq, dk, dv out).  The per-window
// contractions are too small (49x32x49) for a tcgen05 tile on their own; they move to tcgen05 when this kernel is
// fused with the qkv / proj GEMMs ("what comes next").
#include "common.cuh"

#define AM_WARPS 4
#define AM_TILE_BYTES 4096      // 64 tokens x 32 channels bf16
#define AM_LOG2E 1.4426950408889634f

__device__ __forceinline__ uint32_t smem_u32_generic(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct AmGeom {
  int B, H, W, C, heads, wh, ww, sh, sw, Hp, Wp, nwx, nwy, N, ntab, lin_stride, lin_off;
  float scale2;           // head_dim^-0.5 * log2(e)
  float scale;            // head_dim^-0.5
  int windows;            // B * nwy * nwx
  int win_per_warp;       // consecutive windows (same head) handled by one CTA
};

__device__ __forceinline__ uint32_t am_off(int row, int chunk) {   // byte offset inside a [64][32] bf16 tile
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// src_bytes = 0 zero-fills the 16 destination bytes (no global access is made)
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// A fragments of the 16-row strip mt of a [64][32] tile: both k-steps (d = 0..15, 16..31)
__device__ __forceinline__ void am_load_a(uint32_t tile, int mt, int lane, uint32_t (&a)[2][4]) {
  const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) ldsm_x4(tile + am_off(row, ks * 2 + (lane >> 4)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// acc[nt] (nt < NTC column tiles of 8 tokens) = A_strip[16 x 32] * T^T where T = tile [64 tokens][32]
template <int NTC>
__device__ __forceinline__ void am_strip_nt(float (&acc)[8][4], const uint32_t (&a)[2][4], uint32_t tile, int lane) {
#pragma unroll
  for (int nt = 0; nt < NTC; ++nt) {
    acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    uint32_t b0, b1, b2, b3;   // (d 0-7, 8-15, 16-23, 24-31) of tokens 8nt..8nt+7
    ldsm_x4(tile + am_off(nt * 8 + (lane & 7), lane >> 3), b0, b1, b2, b3);
    mma_bf16(acc[nt], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
    mma_bf16(acc[nt], a[1][0], a[1][1], a[1][2], a[1][3], b2, b3);
  }
}

// out[nt] (nt = 0..3: 8-channel tiles) = P[16 x 64] * T where P comes as packed bf16 A fragments per 16-token k-step
__device__ __forceinline__ void am_strip_pv(float (&out)[4][4], const uint32_t (&p)[4][4], uint32_t tile, int lane) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) out[nt][0] = out[nt][1] = out[nt][2] = out[nt][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 2; ++np) {   // two channel tiles per ldmatrix.x4.trans
      uint32_t b0, b1, b2, b3;
      const int mi = lane >> 3;
      ldsm_x4_t(tile + am_off(ks * 16 + (mi & 1) * 8 + (lane & 7), np * 2 + (mi >> 1)), b0, b1, b2, b3);
      mma_bf16(out[np * 2], p[ks][0], p[ks][1], p[ks][2], p[ks][3], b0, b1);
      mma_bf16(out[np * 2 + 1], p[ks][0], p[ks][1], p[ks][2], p[ks][3], b2, b3);
    }
  }
}

// staging strip [16 rows][64 bytes], 16-byte chunks swizzled by (row >> 1) & 3
__device__ __forceinline__ uint32_t am_stage_off(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

// writes a [16 x 32] fp32 accumulator strip (times rowmul) to the staging strip, then to global as 64-byte segments.
// colsum (or null): per-thread partial column sums [8] (columns 8*nt + 2*t + {0,1}) of what is written for real tokens.
__device__ __forceinline__ void am_store_strip(const float (&o)[4][4], float mul_lo, float mul_hi, uint32_t stage_s, uint8_t* stage_g,
                                               int lane, int mt, int N, const int* s_src, bf16* base, int64_t row_stride, int col0,
                                               float* bias_grad /* or null */, float* colsum /* or null */) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float v0 = o[nt][0] * mul_lo, v1 = o[nt][1] * mul_lo, v2 = o[nt][2] * mul_hi, v3 = o[nt][3] * mul_hi;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(stage_s + am_stage_off(g, nt) + t * 4), "r"(pack_bf16(v0, v1)) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(stage_s + am_stage_off(g + 8, nt) + t * 4), "r"(pack_bf16(v2, v3)) : "memory");
    if (colsum) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      const bool ok0 = s_src[r0] >= 0, ok1 = s_src[r1] >= 0;
      colsum[nt * 2] += (ok0 ? v0 : 0.f) + (ok1 ? v2 : 0.f);
      colsum[nt * 2 + 1] += (ok0 ? v1 : 0.f) + (ok1 ? v3 : 0.f);
    }
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int r = it * 8 + (lane >> 2), ch = lane & 3;
    const int tok = mt * 16 + r;
    const int src = s_src[tok];                       // -1 for t >= N as well
    const uint4 v = *reinterpret_cast<const uint4*>(stage_g + am_stage_off(r, ch));
    if (src >= 0) *reinterpret_cast<uint4*>(base + (int64_t)src * row_stride + col0 + ch * 8) = v;
    if (bias_grad && tok < N && src < 0) {   // padded token (rare): its k / v are the qkv bias -> the gradient belongs to the bias
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(h[k]);
        atomicAdd(bias_grad + col0 + ch * 8 + 2 * k, f.x);
        atomicAdd(bias_grad + col0 + ch * 8 + 2 * k + 1, f.y);
      }
    }
  }
}

// loads one [N x 32] slice (64-byte row segments) of a [rows, row_stride] bf16 matrix into a swizzled tile; rows of
// absent tokens (padding, t >= N) are zero-filled by the same cp.async (src-size 0), so the loop has no branches
__device__ __forceinline__ void am_load_tile(uint32_t tile_s, const bf16* base, int64_t row_stride, int col0, const int* s_src, int tid) {
#pragma unroll
  for (int it0 = 0; it0 < 64 * 4; it0 += AM_WARPS * 32) {
    const int it = it0 + tid;
    const int tok = it >> 2, ch = it & 3;
    const int src = s_src[tok];
    const bf16* p = base + (int64_t)(src < 0 ? 0 : src) * row_stride + col0 + ch * 8;
    cp_async16_zfill(tile_s + am_off(tok, ch), p, src < 0 ? 0 : 16);
  }
}

// padded tokens of a window (resolution not divisible by the window): timm pads AFTER norm1, so their q / k / v are
// the qkv bias.  Rare path (only windows on the bottom / right border of such maps): overwrite the zero-filled rows.
__device__ __forceinline__ void am_fill_pad_rows(uint8_t* tile_g, const float* pad_bias, int col0, const int* s_src, int N, int tid) {
  for (int it = tid; it < N * 4; it += AM_WARPS * 32) {
    const int tok = it >> 2, ch = it & 3;
    if (s_src[tok] < 0) {
      const float* pb = pad_bias + col0 + ch * 8;
      uint4 v;
      v.x = pack_bf16(__ldg(pb), __ldg(pb + 1)); v.y = pack_bf16(__ldg(pb + 2), __ldg(pb + 3));
      v.z = pack_bf16(__ldg(pb + 4), __ldg(pb + 5)); v.w = pack_bf16(__ldg(pb + 6), __ldg(pb + 7));
      *reinterpret_cast<uint4*>(tile_g + am_off(tok, ch)) = v;
    }
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---- per-warp bookkeeping in shared memory ----------------------------------------------------------------------
// tok tables (once per warp): s_pos[t] = ty | tx << 8 ; s_lin[t] = ty*(2ww-1)+tx ; s_rg[t] = rh3(ty) | rw(tx) << 8
// per window: s_src[t] (source row or -1)
struct AmWin { int b, wy, wx; bool last_row, last_col; };

__device__ __forceinline__ void am_init_tables(const AmGeom& g, int tid, int* s_pos, int* s_lin, int* s_rg) {
  for (int t = tid; t < 64; t += AM_WARPS * 32) {
    int pos = 0, lin = 0, rg = 0;
    if (t < g.N) {
      const int ty = t / g.ww, tx = t - ty * g.ww;
      pos = ty | (tx << 8);
      lin = ty * g.lin_stride + tx;
      // shift-mask regions of the LAST window row / column: slices (-w, -s) -> 1, (-s, end) -> 2
      const int rh = (ty < g.wh - g.sh) ? 1 : 2, rw = (tx < g.ww - g.sw) ? 1 : 2;
      rg = (rh * 3) | (rw << 8);
    }
    s_pos[t] = pos; s_lin[t] = lin; s_rg[t] = rg;
  }
}

__device__ __forceinline__ AmWin am_window(const AmGeom& g, int w) {
  AmWin it;
  it.wx = w % g.nwx; w /= g.nwx; it.wy = w % g.nwy; it.b = w / g.nwy;
  it.last_row = g.sh > 0 && it.wy == g.nwy - 1;
  it.last_col = g.sw > 0 && it.wx == g.nwx - 1;
  return it;
}

__device__ __forceinline__ bool am_has_pad(const AmGeom& g, const AmWin& w) {
  return (w.wy + 1) * g.wh > g.H || (w.wx + 1) * g.ww > g.W;
}
__device__ __forceinline__ void am_sources(const AmGeom& g, const AmWin& w, int tid, const int* s_pos, int* s_src) {
  for (int t = tid; t < 64; t += AM_WARPS * 32) {
    int src = -1;
    if (t < g.N) {
      const int pos = s_pos[t];
      const int py = w.wy * g.wh + (pos & 0xff), px = w.wx * g.ww + (pos >> 8);
      if (py < g.H && px < g.W) {
        int y = py + g.sh, x = px + g.sw;
        if (y >= g.H) y -= g.H;
        if (x >= g.W) x -= g.W;
        src = (w.b * g.H + y) * g.W + x;
      }
    }
    s_src[t] = src;
  }
}

__device__ __forceinline__ int am_region(const AmWin& w, int rg) {
  return (w.last_row ? (rg & 0xff) : 0) + (w.last_col ? (rg >> 8) : 0);
}

// scores of one strip in the log2 domain: s2 = acc*scale2 + tbl2[lin_q - lin_k + off] (+ mask), invalid columns -> -inf.
// NTC column tiles; rows are queries (TRANSPOSED=false) or keys (TRANSPOSED=true).
// lrow0/1: lin of the two rows (+off folded in by the caller for the non-transposed case: see below)
template <int NTC, bool TRANSPOSED, bool MASKED>
__device__ __forceinline__ void am_scores(float (&acc)[8][4], float scale2, const float* s_tbl2, int lin_off, const int (&clin)[16],
                                          const int (&creg)[16], int lr0, int lr1, int rr0, int rr1, int n_valid_cols, int t) {
  const float MASK2 = -100.0f * AM_LOG2E;
#pragma unroll
  for (int nt = 0; nt < NTC; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = nt * 2 + e;
      const bool col_ok = (nt * 8 + 7 < n_valid_cols) || (nt * 8 + 2 * t + e < n_valid_cols);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int lr = h ? lr1 : lr0;
        const int idx = (TRANSPOSED ? (clin[c] - lr) : (lr - clin[c])) + lin_off;
        float s = fmaf(acc[nt][h * 2 + e], scale2, s_tbl2[idx]);
        if (MASKED) { if (creg[c] != (h ? rr1 : rr0)) s += MASK2; }
        acc[nt][h * 2 + e] = col_ok ? s : -INFINITY;
      }
    }
}

// ---- CTA-cooperative kernels: one CTA (4 warps) per (window group, head); warp w owns the 16-token strip w -------------
// shared memory per CTA (bytes):
//   fwd: tiles q,k,v [3][4096] | stage [4][1024] | table | 4 int[64] tables
//   bwd: tiles q,k,v,dO [4][4096] | stage [4][1024] | dS strips [4][2048] | dQ slabs [4][64][AM_SLAB_LD] fp32 | table | bins |
//        4 int[64] tables | delta[64] | lse[64]
#define AM_SLAB_LD 40
#ifndef AM_WAVES
#define AM_WAVES 3
#endif
static __host__ __device__ inline int am_tab_bytes(int ntab) { return (ntab * 4 + 15) & ~15; }
static __host__ __device__ inline int am_cta_bytes(int ntab, bool bwd) {
  int b = (bwd ? 4 : 3) * AM_TILE_BYTES + AM_WARPS * 1024 + (bwd ? 2 : 1) * am_tab_bytes(ntab) + 4 * 256;
  if (bwd) b += AM_WARPS * 2048 + AM_WARPS * 64 * AM_SLAB_LD * 4 + 2 * 256;
  return b;
}

template <int NTC>
__global__ void __launch_bounds__(AM_WARPS * 32, 4) window_attn_mma_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ table,
                                                                               const float* __restrict__ qkv_bias, bf16* __restrict__ out,
                                                                               float* __restrict__ lse, AmGeom g) {
  extern __shared__ __align__(128) uint8_t am_smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % g.heads, wg = blockIdx.x / g.heads;
  uint8_t* tq = am_smem_raw; uint8_t* tk = tq + AM_TILE_BYTES; uint8_t* tv = tk + AM_TILE_BYTES;
  uint8_t* stage = tv + AM_TILE_BYTES + warp * 1024;
  float* s_tbl = reinterpret_cast<float*>(tv + AM_TILE_BYTES + AM_WARPS * 1024);
  int* s_pos = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_tbl) + am_tab_bytes(g.ntab));
  int* s_lin = s_pos + 64; int* s_rg = s_lin + 64; int* s_src = s_rg + 64;
  const uint32_t tq_s = smem_u32_generic(tq), tk_s = smem_u32_generic(tk), tv_s = smem_u32_generic(tv), stage_s = smem_u32_generic(stage);
  for (int t = tid; t < g.ntab; t += AM_WARPS * 32) s_tbl[t] = __ldg(table + t * g.heads + h) * AM_LOG2E;
  am_init_tables(g, tid, s_pos, s_lin, s_rg);
  __syncthreads();
  const int gq = lane >> 2, tq4 = lane & 3;
  const int n_mt = (g.N + 15) >> 4;
  const int mt = warp;                                  // this warp's query strip
  int clin[16], creg[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) { const int j = (c >> 1) * 8 + 2 * tq4 + (c & 1); clin[c] = s_lin[j]; creg[c] = 0; }
  const int r0 = mt * 16 + gq, r1 = r0 + 8;
  const int lr0 = s_lin[r0], lr1 = s_lin[r1];

  for (int wi = 0; wi < g.win_per_warp; ++wi) {
    const int w = wg * g.win_per_warp + wi;
    if (w >= g.windows) break;
    const AmWin win = am_window(g, w);
    const bool masked = win.last_row || win.last_col;
    const bool has_pad = am_has_pad(g, win);
    __syncthreads();                                    // previous window fully consumed
    am_sources(g, win, tid, s_pos, s_src);
    __syncthreads();
    am_load_tile(tq_s, qkv, 3 * g.C, h * 32, s_src, tid);
    am_load_tile(tk_s, qkv, 3 * g.C, g.C + h * 32, s_src, tid);
    am_load_tile(tv_s, qkv, 3 * g.C, 2 * g.C + h * 32, s_src, tid);
    if (masked) {
#pragma unroll
      for (int c = 0; c < 16; ++c) { const int j = (c >> 1) * 8 + 2 * tq4 + (c & 1); creg[c] = am_region(win, s_rg[j]); }
    }
    cp_async_wait_all();
    if (has_pad) {
      __syncthreads();
      am_fill_pad_rows(tq, qkv_bias, h * 32, s_src, g.N, tid);
      am_fill_pad_rows(tk, qkv_bias, g.C + h * 32, s_src, g.N, tid);
      am_fill_pad_rows(tv, qkv_bias, 2 * g.C + h * 32, s_src, g.N, tid);
    }
    __syncthreads();
    if (mt < n_mt) {
      uint32_t a[2][4];
      am_load_a(tq_s, mt, lane, a);
      float acc[8][4];
      am_strip_nt<NTC>(acc, a, tk_s, lane);
      if (masked) am_scores<NTC, false, true>(acc, g.scale2, s_tbl, g.lin_off, clin, creg, lr0, lr1, am_region(win, s_rg[r0]), am_region(win, s_rg[r1]), g.N, tq4);
      else am_scores<NTC, false, false>(acc, g.scale2, s_tbl, g.lin_off, clin, creg, lr0, lr1, 0, 0, g.N, tq4);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) { mx0 = fmaxf(mx0, fmaxf(acc[nt][0], acc[nt][1])); mx1 = fmaxf(mx1, fmaxf(acc[nt][2], acc[nt][3])); }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t p[4][4];
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) {
        const float e0 = ex2f(acc[nt][0] - mx0), e1 = ex2f(acc[nt][1] - mx0);
        const float e2 = ex2f(acc[nt][2] - mx1), e3 = ex2f(acc[nt][3] - mx1);
        sum0 += e0 + e1; sum1 += e2 + e3;
        p[nt >> 1][(nt & 1) * 2] = pack_bf16(e0, e1);
        p[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(e2, e3);
      }
      if (NTC == 7) { p[3][2] = 0u; p[3][3] = 0u; }
      sum0 = quad_sum(sum0); sum1 = quad_sum(sum1);
      if (lse && tq4 == 0) {
        if (s_src[r0] >= 0) lse[(int64_t)s_src[r0] * g.heads + h] = mx0 + log2f(sum0);
        if (s_src[r1] >= 0) lse[(int64_t)s_src[r1] * g.heads + h] = mx1 + log2f(sum1);
      }
      float o[4][4];
      am_strip_pv(o, p, tv_s, lane);
      am_store_strip(o, 1.0f / sum0, 1.0f / sum1, stage_s, stage, lane, mt, g.N, s_src, out, g.C, h * 32, nullptr, nullptr);
    }
  }
}

template <int NTC>
__global__ void __launch_bounds__(AM_WARPS * 32, 3) window_attn_mma_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ qkv,
                                                                               const bf16* __restrict__ outp, const float* __restrict__ lse,
                                                                               const float* __restrict__ table, const float* __restrict__ qkv_bias,
                                                                               bf16* __restrict__ dqkv, float* __restrict__ dtable,
                                                                               float* __restrict__ dqkv_bias, float* __restrict__ dqkv_colsum,
                                                                               AmGeom g) {
  extern __shared__ __align__(128) uint8_t am_smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % g.heads, wg = blockIdx.x / g.heads;
  uint8_t* tq = am_smem_raw; uint8_t* tk = tq + AM_TILE_BYTES; uint8_t* tv = tk + AM_TILE_BYTES; uint8_t* tdo = tv + AM_TILE_BYTES;
  uint8_t* stage = tdo + AM_TILE_BYTES + warp * 1024;
  uint8_t* sds = tdo + AM_TILE_BYTES + AM_WARPS * 1024 + warp * 2048;          // this warp's dS^T strip [16 keys][64 queries] bf16
  float* slabs = reinterpret_cast<float*>(tdo + AM_TILE_BYTES + AM_WARPS * 1024 + AM_WARPS * 2048);
  float* slab = slabs + warp * 64 * AM_SLAB_LD;                                // this warp's partial dQ [64][AM_SLAB_LD]
  float* s_tbl = slabs + AM_WARPS * 64 * AM_SLAB_LD;
  const int tab_bytes = am_tab_bytes(g.ntab);
  float* s_bins = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_tbl) + tab_bytes);
  int* s_pos = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_bins) + tab_bytes);
  int* s_lin = s_pos + 64; int* s_rg = s_lin + 64; int* s_src = s_rg + 64;
  float* s_delta = reinterpret_cast<float*>(s_src + 64);
  float* s_lse = s_delta + 64;
  const uint32_t tq_s = smem_u32_generic(tq), tk_s = smem_u32_generic(tk), tv_s = smem_u32_generic(tv), tdo_s = smem_u32_generic(tdo),
                 stage_s = smem_u32_generic(stage), sds_s = smem_u32_generic(sds);
  for (int t = tid; t < g.ntab; t += AM_WARPS * 32) { s_tbl[t] = __ldg(table + t * g.heads + h) * AM_LOG2E; s_bins[t] = 0.f; }
  am_init_tables(g, tid, s_pos, s_lin, s_rg);
  __syncthreads();
  const int gq = lane >> 2, tq4 = lane & 3;
  const int n_mt = (g.N + 15) >> 4;
  const int jt = warp;                                  // this warp's key strip
  int clin[16], creg[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) { const int j = (c >> 1) * 8 + 2 * tq4 + (c & 1); clin[c] = s_lin[j]; creg[c] = 0; }
  const int r0 = jt * 16 + gq, r1 = r0 + 8;
  const int lr0 = s_lin[r0], lr1 = s_lin[r1];
  const bool row0_ok = r0 < g.N, row1_ok = r1 < g.N;
  float gsum[8][4];                     // dS^T of this warp's strip summed over the CTA's windows (bias-table gradient)
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) gsum[nt][0] = gsum[nt][1] = gsum[nt][2] = gsum[nt][3] = 0.f;
  float cs_k[8], cs_v[8], cs_q[8];      // qkv-bias gradient partials
#pragma unroll
  for (int k = 0; k < 8; ++k) cs_k[k] = cs_v[k] = cs_q[k] = 0.f;
  const float2* lse2 = reinterpret_cast<const float2*>(s_lse) + tq4;
  const float2* del2 = reinterpret_cast<const float2*>(s_delta) + tq4;

  for (int wi = 0; wi < g.win_per_warp; ++wi) {
    const int w = wg * g.win_per_warp + wi;
    if (w >= g.windows) break;
    const AmWin win = am_window(g, w);
    const bool masked = win.last_row || win.last_col;
    const bool has_pad = am_has_pad(g, win);
    __syncthreads();                                    // previous window fully consumed (tiles, slabs, tables)
    am_sources(g, win, tid, s_pos, s_src);
    __syncthreads();
    am_load_tile(tq_s, qkv, 3 * g.C, h * 32, s_src, tid);
    am_load_tile(tk_s, qkv, 3 * g.C, g.C + h * 32, s_src, tid);
    am_load_tile(tv_s, qkv, 3 * g.C, 2 * g.C + h * 32, s_src, tid);
    am_load_tile(tdo_s, dout, g.C, h * 32, s_src, tid);      // padded tokens: dO = 0 (cropped away)
    // delta_i = dO_i . O_i straight from global (4 lanes per token, 16 bytes each); log-sum-exp of the query rows
#pragma unroll
    for (int it0 = 0; it0 < 64 * 4; it0 += AM_WARPS * 32) {
      const int it = it0 + tid;
      const int tok = it >> 2, ch = it & 3;
      float part = 0.f;
      const int src = s_src[tok];
      if (src >= 0) {
        float a[8], b[8];
        IO<bf16>::load8(dout + (int64_t)src * g.C + h * 32 + ch * 8, a);
        IO<bf16>::load8(outp + (int64_t)src * g.C + h * 32 + ch * 8, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) part = fmaf(a[k], b[k], part);
      }
      part = quad_sum(part);
      if (ch == 0) {
        s_delta[tok] = part;
        s_lse[tok] = src >= 0 ? __ldg(lse + (int64_t)src * g.heads + h) : 1e30f;   // padded / absent query: P = 0
      }
    }
    if (masked) {
#pragma unroll
      for (int c = 0; c < 16; ++c) { const int j = (c >> 1) * 8 + 2 * tq4 + (c & 1); creg[c] = am_region(win, s_rg[j]); }
    }
    cp_async_wait_all();
    if (has_pad) {
      __syncthreads();
      am_fill_pad_rows(tq, qkv_bias, h * 32, s_src, g.N, tid);
      am_fill_pad_rows(tk, qkv_bias, g.C + h * 32, s_src, g.N, tid);
      am_fill_pad_rows(tv, qkv_bias, 2 * g.C + h * 32, s_src, g.N, tid);
    }
    __syncthreads();

    if (jt < n_mt) {
      uint32_t a[2][4];
      am_load_a(tk_s, jt, lane, a);
      float acc[8][4];
      am_strip_nt<NTC>(acc, a, tq_s, lane);          // S^T[j][i] = k_j . q_i
      if (masked) am_scores<NTC, true, true>(acc, g.scale2, s_tbl, g.lin_off, clin, creg, lr0, lr1, am_region(win, s_rg[r0]), am_region(win, s_rg[r1]), g.N, tq4);
      else am_scores<NTC, true, false>(acc, g.scale2, s_tbl, g.lin_off, clin, creg, lr0, lr1, 0, 0, g.N, tq4);
      uint32_t pt[4][4];
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) {
        const float2 ls = lse2[nt * 4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = ex2f(acc[nt][e] - ((e & 1) ? ls.y : ls.x));   // -inf / lse 1e30 -> 0
          acc[nt][e] = ((e >> 1) ? row1_ok : row0_ok) ? pv : 0.f;
        }
        pt[nt >> 1][(nt & 1) * 2] = pack_bf16(acc[nt][0], acc[nt][1]);
        pt[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(acc[nt][2], acc[nt][3]);
      }
      if (NTC == 7) { pt[3][2] = 0u; pt[3][3] = 0u; }
      float o[4][4];
      am_strip_pv(o, pt, tdo_s, lane);          // dV[j][d] = sum_i P[i][j] dO[i][d]
      am_store_strip(o, 1.f, 1.f, stage_s, stage, lane, jt, g.N, s_src, dqkv, 3 * g.C, 2 * g.C + h * 32, has_pad ? dqkv_bias : nullptr, dqkv_colsum ? cs_v : nullptr);
      uint32_t av[2][4];
      am_load_a(tv_s, jt, lane, av);
      float dp[8][4];
      am_strip_nt<NTC>(dp, av, tdo_s, lane);    // dP^T[j][i] = v_j . dO_i
      uint32_t dst[4][4];
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) {
        float v[4];
        const float2 dl = del2[nt * 4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[e] = acc[nt][e] * (dp[nt][e] - ((e & 1) ? dl.y : dl.x));     // dS^T[j][i]
          gsum[nt][e] += v[e];
        }
        dst[nt >> 1][(nt & 1) * 2] = pack_bf16(v[0], v[1]);
        dst[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(v[2], v[3]);
      }
      if (NTC == 7) { dst[3][2] = 0u; dst[3][3] = 0u; }
      am_strip_pv(o, dst, tq_s, lane);          // dK[j][d] = scale * sum_i dS[i][j] q[i][d]
      am_store_strip(o, g.scale, g.scale, stage_s, stage, lane, jt, g.N, s_src, dqkv, 3 * g.C, g.C + h * 32, has_pad ? dqkv_bias : nullptr, dqkv_colsum ? cs_k : nullptr);
      // dS^T strip -> shared [16 keys][64 queries] (16-byte chunks swizzled by key row), then partial dQ = dS * K_strip
      __syncwarp();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int c0 = ks * 2, c1 = ks * 2 + 1;   // 16-byte chunk index = query column / 8
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + gq * 128 + ((c0 ^ (gq & 7)) << 4) + tq4 * 4), "r"(dst[ks][0]) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + (gq + 8) * 128 + ((c0 ^ ((gq + 8) & 7)) << 4) + tq4 * 4), "r"(dst[ks][1]) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + gq * 128 + ((c1 ^ (gq & 7)) << 4) + tq4 * 4), "r"(dst[ks][2]) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + (gq + 8) * 128 + ((c1 ^ ((gq + 8) & 7)) << 4) + tq4 * 4), "r"(dst[ks][3]) : "memory");
      }
      __syncwarp();
      uint32_t kb[4][2];            // B fragments: K strip [k = 16 keys][n = 32 channels], transposed loads from the K tile
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        const int mi = lane >> 3;
        ldsm_x4_t(tk_s + am_off(jt * 16 + (mi & 1) * 8 + (lane & 7), np * 2 + (mi >> 1)), kb[np * 2][0], kb[np * 2][1], kb[np * 2 + 1][0], kb[np * 2 + 1][1]);
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {          // query strips
        if (it < n_mt) {
          uint32_t a0, a1, a2, a3;               // A fragment (16 queries x 16 keys) = transpose of the stored [key][query] block
          const int mi = lane >> 3;
          const int key = (mi >> 1) * 8 + (lane & 7), chunk = it * 2 + (mi & 1);
          ldsm_x4_t(sds_s + key * 128 + ((chunk ^ (key & 7)) << 4), a0, a1, a2, a3);
          float dq[4][4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
            mma_bf16(dq[nt], a0, a1, a2, a3, kb[nt][0], kb[nt][1]);
            float* p0 = slab + (it * 16 + gq) * AM_SLAB_LD + nt * 8 + 2 * tq4;
            *reinterpret_cast<float2*>(p0) = make_float2(dq[nt][0], dq[nt][1]);
            *reinterpret_cast<float2*>(p0 + 8 * AM_SLAB_LD) = make_float2(dq[nt][2], dq[nt][3]);
          }
        }
      }
    }
    __syncthreads();
    // dQ = scale * sum over key strips of the partial slabs -> bf16 -> global (64-byte row segments)
#pragma unroll
    for (int it0 = 0; it0 < 64 * 4; it0 += AM_WARPS * 32) {
      const int it = it0 + tid;
      const int tok = it >> 2, ch = it & 3;
      const int src = s_src[tok];
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
      for (int sw = 0; sw < n_mt; ++sw) {
        const float4* p = reinterpret_cast<const float4*>(slabs + (sw * 64 + tok) * AM_SLAB_LD + ch * 8);
        const float4 x = p[0], y = p[1];
        v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w; v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w;
      }
      if (src >= 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] *= g.scale; cs_q[k] += v[k]; }
        IO<bf16>::store8(dqkv + (int64_t)src * 3 * g.C + h * 32 + ch * 8, v);
      }
    }
  }
  // ---- flush: bias-table gradient (register sums -> shared bins -> global), qkv-bias gradient ----
  if (jt < n_mt) {
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * tq4 + (e & 1), row = (e >> 1) ? r1 : r0;
        if (col < g.N && row < g.N) atomicAdd(&s_bins[clin[nt * 2 + (e & 1)] - ((e >> 1) ? lr1 : lr0) + g.lin_off], gsum[nt][e]);
      }
  }
  __syncthreads();
  for (int t = tid; t < g.ntab; t += AM_WARPS * 32) { const float v = s_bins[t]; if (v != 0.f) atomicAdd(dtable + t * g.heads + h, v); }
  if (dqkv_colsum) {
    // k / v partials: columns 8*nt + 2*t + {0,1} per lane, summed over the 8 row groups (lanes with the same t)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float kk = cs_k[k], v = cs_v[k];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) { kk += __shfl_xor_sync(0xffffffffu, kk, o); v += __shfl_xor_sync(0xffffffffu, v, o); }
      if (gq == 0) {
        const int col = h * 32 + (k >> 1) * 8 + 2 * tq4 + (k & 1);
        atomicAdd(dqkv_colsum + g.C + col, kk); atomicAdd(dqkv_colsum + 2 * g.C + col, v);
      }
    }
    // q partials: thread (tok, ch) pattern -> columns ch*8 + k, summed over lanes with the same ch (lane & 3)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float q = cs_q[k];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      if (gq == 0) atomicAdd(dqkv_colsum + h * 32 + tq4 * 8 + k, q);
    }
  }
}

// ---------------------------------------------------------------------------------------------
static int am_geom(AmGeom& g, int B, int H, int W, int C, int heads, int wh, int ww, int sh, int sw) {
  if (B < 0 || H <= 0 || W <= 0 || heads <= 0 || C != heads * 32) return MTUS_ERR_BAD_ARG;
  if (wh <= 0 || ww <= 0 || wh * ww > 64 || sh < 0 || sw < 0 || sh >= wh || sw >= ww) return MTUS_ERR_UNSUPPORTED;
  g.B = B; g.H = H; g.W = W; g.C = C; g.heads = heads; g.wh = wh; g.ww = ww; g.sh = sh; g.sw = sw;
  g.nwy = (H + wh - 1) / wh; g.nwx = (W + ww - 1) / ww;
  g.Hp = g.nwy * wh; g.Wp = g.nwx * ww;
  g.N = wh * ww; g.ntab = (2 * wh - 1) * (2 * ww - 1);
  g.lin_stride = 2 * ww - 1; g.lin_off = (wh - 1) * (2 * ww - 1) + (ww - 1);
  g.scale = 1.0f / sqrtf(32.0f);
  g.scale2 = g.scale * AM_LOG2E;
  g.windows = B * g.nwy * g.nwx;
  return MTUS_OK;
}

bool mtus_window_attn_mma_supported(int wh, int ww, int dtype) { return dtype == MTUS_BF16 && wh * ww <= 64 && wh > 0 && ww > 0; }

// windows per CTA: enough CTAs to fill the machine several times over, otherwise as many windows per CTA as possible
// (table loads, bias-gradient flushes and launch overhead amortise over them)
static int am_plan(AmGeom& g, int target_ctas) {
  int wpw = 1;
  while ((int64_t)((g.windows + wpw * 2 - 1) / (wpw * 2)) * g.heads >= target_ctas && wpw < 32) wpw *= 2;
  g.win_per_warp = wpw;
  return ((g.windows + wpw - 1) / wpw) * g.heads;
}

int mtus_window_attn_mma_fwd(const void* qkv, const float* rel_table, const float* qkv_bias, void* out, float* lse, int B, int H, int W,
                             int C, int heads, int win_h, int win_w, int shift_h, int shift_w, cudaStream_t st) {
  AmGeom g;
  int rc = am_geom(g, B, H, W, C, heads, win_h, win_w, shift_h, shift_w);
  if (rc) return rc;
  if ((g.Hp != H || g.Wp != W) && !qkv_bias) return MTUS_ERR_BAD_ARG;
  if (B == 0) return MTUS_OK;
  const int blocks = am_plan(g, 148 * 4 * AM_WAVES);
  const size_t smem = am_cta_bytes(g.ntab, false);
  if (g.N <= 56) window_attn_mma_fwd_kernel<7><<<blocks, AM_WARPS * 32, smem, st>>>((const bf16*)qkv, rel_table, qkv_bias, (bf16*)out, lse, g);
  else window_attn_mma_fwd_kernel<8><<<blocks, AM_WARPS * 32, smem, st>>>((const bf16*)qkv, rel_table, qkv_bias, (bf16*)out, lse, g);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

int mtus_window_attn_mma_bwd(const void* dout, const void* qkv, const void* out, const float* lse, const float* rel_table,
                             const float* qkv_bias, void* dqkv, float* drel_table, float* dqkv_bias, float* dqkv_colsum, int B, int H,
                             int W, int C, int heads, int win_h, int win_w, int shift_h, int shift_w, cudaStream_t st) {
  AmGeom g;
  int rc = am_geom(g, B, H, W, C, heads, win_h, win_w, shift_h, shift_w);
  if (rc) return rc;
  if (!lse) return MTUS_ERR_BAD_ARG;
  if ((g.Hp != H || g.Wp != W) && !(qkv_bias && dqkv_bias)) return MTUS_ERR_BAD_ARG;
  if (B == 0) return MTUS_OK;
  const int blocks = am_plan(g, 148 * 3 * AM_WAVES);
  const size_t smem = am_cta_bytes(g.ntab, true);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_mma_bwd_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attn_mma_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  if (g.N <= 56)
    window_attn_mma_bwd_kernel<7><<<blocks, AM_WARPS * 32, smem, st>>>((const bf16*)dout, (const bf16*)qkv, (const bf16*)out, lse, rel_table, qkv_bias,
                                                                       (bf16*)dqkv, drel_table, dqkv_bias, dqkv_colsum, g);
  else
    window_attn_mma_bwd_kernel<8><<<blocks, AM_WARPS * 32, smem, st>>>((const bf16*)dout, (const bf16*)qkv, (const bf16*)out, lse, rel_table, qkv_bias,
                                                                       (bf16*)dqkv, drel_table, dqkv_bias, dqkv_colsum, g);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}
