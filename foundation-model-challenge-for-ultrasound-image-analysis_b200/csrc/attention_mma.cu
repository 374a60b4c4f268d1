// Fused shifted-window attention on tensor cores, bf16 storage / fp32 softmax (windows of up to 64 tokens).
//
// Same contract as attention.cu (timm _attn + WindowAttention minus the two Linear layers; SURVEY section 8a rows
// a5, a6): the cyclic shift, zero padding, window partition / reverse, relative-position bias, shift mask,
// softmax and P@V happen inside the kernel and nothing window-shaped touches HBM.  Differences:
//   * one WARP owns one (window, head) item and is fully autonomous (no block-level barrier): q/k/v (and dO)
//     tiles [64 x 32] bf16 are staged in swizzled shared memory with 16-byte cp.async, every contraction runs on
//     the tensor cores (mma.sync m16n8k16, fp32 accumulate) in 16-row strips, the softmax works on the accumulator
//     fragments, and outputs leave through a 1 KB staging strip as 64-byte row segments (full sectors);
//   * the backward pass recomputes S twice (row strips for dQ, column strips for dK / dV) instead of transposing
//     P and dS through shared memory;  the relative-position-bias gradient is binned in shared memory over all
//     the windows a warp processes and flushed with one atomicAdd per bin.
// A stand-alone attention kernel is HBM-bound (24.5 FLOP/B at 49 tokens): algorithmic bytes per token are
// 4*C*2 forward (q, k, v in; o out) and 8*C*2 backward (q, k, v, o, dO in; dq, dk, dv out).  The per-window
// contractions are too small (49x32x49) for a tcgen05 tile on their own; they move to tcgen05 when this kernel is
// fused with the qkv / proj GEMMs ("what comes next").
#include "common.cuh"

#define AM_WARPS 4
#define AM_TILE_BYTES 4096      // 64 tokens x 32 channels bf16

__device__ __forceinline__ uint32_t smem_u32_generic(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct AmGeom {
  int B, H, W, C, heads, wh, ww, sh, sw, Hp, Wp, nwx, nwy, N, ntab, lin_stride, lin_off;
  float scale;
  int windows;            // B * nwy * nwx
  int win_per_warp;       // consecutive windows (same head) handled by one warp
};

__device__ __forceinline__ void am_token(const AmGeom& g, int b, int wy, int wx, int t, int& src, int& reg) {
  const int ty = t / g.ww, tx = t - ty * g.ww;
  const int py = wy * g.wh + ty, px = wx * g.ww + tx;
  int rh = 0, rw = 0;
  if (g.sh > 0) rh = (py < g.Hp - g.wh) ? 0 : ((py < g.Hp - g.sh) ? 1 : 2);
  if (g.sw > 0) rw = (px < g.Wp - g.ww) ? 0 : ((px < g.Wp - g.sw) ? 1 : 2);
  reg = rh * 3 + rw;
  if (py >= g.H || px >= g.W) { src = -1; return; }
  const int y = (py + g.sh) % g.H, x = (px + g.sw) % g.W;
  src = (b * g.H + y) * g.W + x;
}

__device__ __forceinline__ uint32_t am_off(int row, int chunk) {   // byte offset inside a [64][32] bf16 tile
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
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

// A fragments of the 16-row strip mt of a [64][32] tile: both k-steps (d = 0..15, 16..31)
__device__ __forceinline__ void am_load_a(uint32_t tile, int mt, int lane, uint32_t (&a)[2][4]) {
  const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) ldsm_x4(tile + am_off(row, ks * 2 + (lane >> 4)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// acc[nt] (nt = 0..7: 8 column tiles of 8 tokens) = A_strip[16 x 32] * T^T where T = tile [64 tokens][32]
__device__ __forceinline__ void am_strip_nt(float (&acc)[8][4], const uint32_t (&a)[2][4], uint32_t tile, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
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

// writes a [16 x 32] fp32 accumulator strip (times rowmul) to the staging strip, then to global as 64-byte segments
__device__ __forceinline__ void am_store_strip(const float (&o)[4][4], float mul_lo, float mul_hi, uint32_t stage_s, uint8_t* stage_g,
                                               int lane, int mt, int N, const int* s_src, bf16* base, int64_t row_stride, int col0,
                                               float* bias_grad /* or null */) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const uint32_t lo = pack_bf16(o[nt][0] * mul_lo, o[nt][1] * mul_lo), hi = pack_bf16(o[nt][2] * mul_hi, o[nt][3] * mul_hi);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(stage_s + g * 64 + nt * 16 + t * 4), "r"(lo) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(stage_s + (g + 8) * 64 + nt * 16 + t * 4), "r"(hi) : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int r = it * 8 + (lane >> 2), ch = lane & 3;
    const int tok = mt * 16 + r;
    if (tok < N) {
      const uint4 v = *reinterpret_cast<const uint4*>(stage_g + r * 64 + ch * 16);
      const int src = s_src[tok];
      if (src >= 0) *reinterpret_cast<uint4*>(base + (int64_t)src * row_stride + col0 + ch * 8) = v;
      else if (bias_grad) {   // padded token: its k / v are the qkv bias -> the gradient belongs to the bias
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
}

// loads one [N x 32] slice (64-byte row segments) of a [rows, row_stride] bf16 matrix into a swizzled tile
__device__ __forceinline__ void am_load_tile(uint32_t tile_s, uint8_t* tile_g, const bf16* base, int64_t row_stride, int col0,
                                             const int* s_src, int N, int lane, const float* pad_bias /* or null -> zeros */) {
  for (int it = lane; it < 64 * 4; it += 32) {
    const int tok = it >> 2, ch = it & 3;
    const uint32_t off = am_off(tok, ch);
    const int src = tok < N ? s_src[tok] : -1;
    if (src >= 0) cp_async16(tile_s + off, base + (int64_t)src * row_stride + col0 + ch * 8);
    else {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (tok < N && pad_bias) {
        const float* pb = pad_bias + col0 + ch * 8;
        v.x = pack_bf16(__ldg(pb), __ldg(pb + 1)); v.y = pack_bf16(__ldg(pb + 2), __ldg(pb + 3));
        v.z = pack_bf16(__ldg(pb + 4), __ldg(pb + 5)); v.w = pack_bf16(__ldg(pb + 6), __ldg(pb + 7));
      }
      *reinterpret_cast<uint4*>(tile_g + off) = v;
    }
  }
}

struct AmCols {      // per-thread column bookkeeping: 16 columns j = 8*nt + 2*t + {0,1}
  int lin[16];       // relative-position linear coordinate (ty*(2ww-1)+tx) | region << 16 ; -1 for j >= N
};

__device__ __forceinline__ void am_cols(const AmGeom& g, const int* s_reg, int lane, AmCols& c) {
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = nt * 8 + 2 * t + e;
      int v = -1;
      if (j < g.N) { const int ty = j / g.ww, tx = j - ty * g.ww; v = (ty * g.lin_stride + tx) | (s_reg[j] << 16); }
      c.lin[nt * 2 + e] = v;
    }
}

// scores of one strip: s = acc*scale + bias[lin_r - lin_c + off] (+ -100 when the regions differ); invalid columns -> -inf
// rows r0 = 16*mt + g, r1 = r0 + 8 ; (lr0, lr1) their lin|region words (or -1)
template <bool TRANSPOSED>
__device__ __forceinline__ void am_scores(float (&acc)[8][4], const AmGeom& g, const float* s_tbl, const AmCols& c, int lr0, int lr1) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int lc = c.lin[nt * 2 + e];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int lr = h ? lr1 : lr0;
        float s = -INFINITY;
        if (lc >= 0 && lr >= 0) {
          // bias index: (row token as query i, column token as key j); transposed strips have rows = keys
          const int d = TRANSPOSED ? ((lc & 0xffff) - (lr & 0xffff)) : ((lr & 0xffff) - (lc & 0xffff));
          s = acc[nt][h * 2 + e] * g.scale + s_tbl[d + g.lin_off];
          if ((lc >> 16) != (lr >> 16)) s += -100.0f;
        }
        acc[nt][h * 2 + e] = s;
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

struct AmItem { int b, wy, wx; };
__device__ __forceinline__ AmItem am_item(const AmGeom& g, int w) {
  AmItem it; it.wx = w % g.nwx; w /= g.nwx; it.wy = w % g.nwy; it.b = w / g.nwy; return it;
}

// per-warp shared memory layout (bytes): tiles [NT][4096] | stage 1024 | table ntab*4 | bins ntab*4 (bwd) | src 256 | reg 256 | stats 3*256 (bwd)
template <int NT>
__device__ __forceinline__ uint8_t* am_warp_smem(uint8_t* base, int warp, int ntab, bool bwd) {
  const int per = NT * AM_TILE_BYTES + 1024 + (bwd ? 2 : 1) * ((ntab * 4 + 15) & ~15) + 512 + (bwd ? 768 : 0);
  return base + (size_t)warp * ((per + 127) & ~127);
}
static size_t am_smem_bytes(int nt, int ntab, bool bwd) {
  const int per = nt * AM_TILE_BYTES + 1024 + (bwd ? 2 : 1) * ((ntab * 4 + 15) & ~15) + 512 + (bwd ? 768 : 0);
  return (size_t)AM_WARPS * ((per + 127) & ~127) + 128;
}

__global__ void __launch_bounds__(AM_WARPS * 32) window_attn_mma_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ table,
                                                                            const float* __restrict__ qkv_bias, bf16* __restrict__ out,
                                                                            AmGeom g, int n_items) {
  extern __shared__ uint8_t am_smem_raw[];
  uint8_t* sm0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(am_smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * AM_WARPS + warp;       // item = (window group, head)
  if (item >= n_items) return;
  const int h = item % g.heads, wg = item / g.heads;
  uint8_t* sm = am_warp_smem<3>(sm0, warp, g.ntab, false);
  uint8_t* tq = sm; uint8_t* tk = sm + AM_TILE_BYTES; uint8_t* tv = sm + 2 * AM_TILE_BYTES;
  uint8_t* stage = sm + 3 * AM_TILE_BYTES;
  float* s_tbl = reinterpret_cast<float*>(stage + 1024);
  int* s_src = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_tbl) + ((g.ntab * 4 + 15) & ~15));
  int* s_reg = s_src + 64;
  const uint32_t tq_s = smem_u32_generic(tq), tk_s = smem_u32_generic(tk), tv_s = smem_u32_generic(tv), stage_s = smem_u32_generic(stage);
  for (int t = lane; t < g.ntab; t += 32) s_tbl[t] = __ldg(table + t * g.heads + h);
  const int gq = lane >> 2;
  const int n_mt = (g.N + 15) >> 4;
  const float LOG2E = 1.4426950408889634f;

  for (int wi = 0; wi < g.win_per_warp; ++wi) {
    const int w = wg * g.win_per_warp + wi;
    if (w >= g.windows) break;
    const AmItem itw = am_item(g, w);
    __syncwarp();
    for (int t = lane; t < 64; t += 32) {
      int src = -1, reg = 0;
      if (t < g.N) am_token(g, itw.b, itw.wy, itw.wx, t, src, reg);
      s_src[t] = src; s_reg[t] = reg;
    }
    __syncwarp();
    am_load_tile(tq_s, tq, qkv, 3 * g.C, h * 32, s_src, g.N, lane, qkv_bias);
    am_load_tile(tk_s, tk, qkv, 3 * g.C, g.C + h * 32, s_src, g.N, lane, qkv_bias);
    am_load_tile(tv_s, tv, qkv, 3 * g.C, 2 * g.C + h * 32, s_src, g.N, lane, qkv_bias);
    AmCols cols;
    am_cols(g, s_reg, lane, cols);
    cp_async_wait_all();
    __syncwarp();
    for (int mt = 0; mt < n_mt; ++mt) {
      uint32_t a[2][4];
      am_load_a(tq_s, mt, lane, a);
      float acc[8][4];
      am_strip_nt(acc, a, tk_s, lane);
      const int r0 = mt * 16 + gq, r1 = r0 + 8;
      int lr0 = -1, lr1 = -1;
      if (r0 < g.N) { const int ty = r0 / g.ww, tx = r0 - ty * g.ww; lr0 = (ty * g.lin_stride + tx) | (s_reg[r0] << 16); }
      if (r1 < g.N) { const int ty = r1 / g.ww, tx = r1 - ty * g.ww; lr1 = (ty * g.lin_stride + tx) | (s_reg[r1] << 16); }
      am_scores<false>(acc, g, s_tbl, cols, lr0, lr1);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { mx0 = fmaxf(mx0, fmaxf(acc[nt][0], acc[nt][1])); mx1 = fmaxf(mx1, fmaxf(acc[nt][2], acc[nt][3])); }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
      if (mx0 == -INFINITY) mx0 = 0.f;
      if (mx1 == -INFINITY) mx1 = 0.f;
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t p[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float e0 = exp2f((acc[nt][0] - mx0) * LOG2E), e1 = exp2f((acc[nt][1] - mx0) * LOG2E);
        const float e2 = exp2f((acc[nt][2] - mx1) * LOG2E), e3 = exp2f((acc[nt][3] - mx1) * LOG2E);
        sum0 += e0 + e1; sum1 += e2 + e3;
        p[nt >> 1][(nt & 1) * 2] = pack_bf16(e0, e1);
        p[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(e2, e3);
      }
      sum0 = quad_sum(sum0); sum1 = quad_sum(sum1);
      float o[4][4];
      am_strip_pv(o, p, tv_s, lane);
      am_store_strip(o, sum0 > 0.f ? 1.0f / sum0 : 0.f, sum1 > 0.f ? 1.0f / sum1 : 0.f, stage_s, stage, lane, mt, g.N, s_src, out, g.C,
                     h * 32, nullptr);
    }
  }
}

__global__ void __launch_bounds__(AM_WARPS * 32) window_attn_mma_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ qkv,
                                                                            const bf16* __restrict__ outp, const float* __restrict__ table,
                                                                            const float* __restrict__ qkv_bias, bf16* __restrict__ dqkv,
                                                                            float* __restrict__ dtable, float* __restrict__ dqkv_bias,
                                                                            AmGeom g, int n_items) {
  extern __shared__ uint8_t am_smem_raw[];
  uint8_t* sm0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(am_smem_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * AM_WARPS + warp;
  if (item >= n_items) return;
  const int h = item % g.heads, wg = item / g.heads;
  uint8_t* sm = am_warp_smem<4>(sm0, warp, g.ntab, true);
  uint8_t* tq = sm; uint8_t* tk = sm + AM_TILE_BYTES; uint8_t* tv = sm + 2 * AM_TILE_BYTES; uint8_t* tdo = sm + 3 * AM_TILE_BYTES;
  uint8_t* stage = sm + 4 * AM_TILE_BYTES;
  const int tab_bytes = (g.ntab * 4 + 15) & ~15;
  float* s_tbl = reinterpret_cast<float*>(stage + 1024);
  float* s_bins = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_tbl) + tab_bytes);
  int* s_src = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_bins) + tab_bytes);
  int* s_reg = s_src + 64;
  float* s_delta = reinterpret_cast<float*>(s_reg + 64);
  float* s_max = s_delta + 64;
  float* s_inv = s_max + 64;
  const uint32_t tq_s = smem_u32_generic(tq), tk_s = smem_u32_generic(tk), tv_s = smem_u32_generic(tv), tdo_s = smem_u32_generic(tdo),
                 stage_s = smem_u32_generic(stage);
  for (int t = lane; t < g.ntab; t += 32) { s_tbl[t] = __ldg(table + t * g.heads + h); s_bins[t] = 0.f; }
  const int gq = lane >> 2;
  const int n_mt = (g.N + 15) >> 4;
  const float LOG2E = 1.4426950408889634f;

  for (int wi = 0; wi < g.win_per_warp; ++wi) {
    const int w = wg * g.win_per_warp + wi;
    if (w >= g.windows) break;
    const AmItem itw = am_item(g, w);
    __syncwarp();
    for (int t = lane; t < 64; t += 32) {
      int src = -1, reg = 0;
      if (t < g.N) am_token(g, itw.b, itw.wy, itw.wx, t, src, reg);
      s_src[t] = src; s_reg[t] = reg;
    }
    __syncwarp();
    am_load_tile(tq_s, tq, qkv, 3 * g.C, h * 32, s_src, g.N, lane, qkv_bias);
    am_load_tile(tk_s, tk, qkv, 3 * g.C, g.C + h * 32, s_src, g.N, lane, qkv_bias);
    am_load_tile(tv_s, tv, qkv, 3 * g.C, 2 * g.C + h * 32, s_src, g.N, lane, qkv_bias);
    am_load_tile(tdo_s, tdo, dout, g.C, h * 32, s_src, g.N, lane, nullptr);      // padded tokens: dO = 0 (cropped away)
    // delta_i = dO_i . O_i straight from global (4 lanes per token, 16 bytes each)
    for (int it = lane; it < 64 * 4; it += 32) {
      const int tok = it >> 2, ch = it & 3;
      float part = 0.f;
      const int src = tok < g.N ? s_src[tok] : -1;
      if (src >= 0) {
        float a[8], b[8];
        IO<bf16>::load8(dout + (int64_t)src * g.C + h * 32 + ch * 8, a);
        IO<bf16>::load8(outp + (int64_t)src * g.C + h * 32 + ch * 8, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) part = fmaf(a[k], b[k], part);
      }
      part = quad_sum(part);
      if (ch == 0) s_delta[tok] = part;
    }
    AmCols cols;
    am_cols(g, s_reg, lane, cols);
    cp_async_wait_all();
    __syncwarp();

    // ---------------- pass A: row strips (queries): P, dP, dS -> dQ, bias-table gradient ----------------
    for (int mt = 0; mt < n_mt; ++mt) {
      uint32_t a[2][4];
      am_load_a(tq_s, mt, lane, a);
      float acc[8][4];
      am_strip_nt(acc, a, tk_s, lane);
      const int r0 = mt * 16 + gq, r1 = r0 + 8;
      int lr0 = -1, lr1 = -1;
      if (r0 < g.N) { const int ty = r0 / g.ww, tx = r0 - ty * g.ww; lr0 = (ty * g.lin_stride + tx) | (s_reg[r0] << 16); }
      if (r1 < g.N) { const int ty = r1 / g.ww, tx = r1 - ty * g.ww; lr1 = (ty * g.lin_stride + tx) | (s_reg[r1] << 16); }
      am_scores<false>(acc, g, s_tbl, cols, lr0, lr1);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { mx0 = fmaxf(mx0, fmaxf(acc[nt][0], acc[nt][1])); mx1 = fmaxf(mx1, fmaxf(acc[nt][2], acc[nt][3])); }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
      if (mx0 == -INFINITY) mx0 = 0.f;
      if (mx1 == -INFINITY) mx1 = 0.f;
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        acc[nt][0] = exp2f((acc[nt][0] - mx0) * LOG2E); acc[nt][1] = exp2f((acc[nt][1] - mx0) * LOG2E);
        acc[nt][2] = exp2f((acc[nt][2] - mx1) * LOG2E); acc[nt][3] = exp2f((acc[nt][3] - mx1) * LOG2E);
        sum0 += acc[nt][0] + acc[nt][1]; sum1 += acc[nt][2] + acc[nt][3];
      }
      sum0 = quad_sum(sum0); sum1 = quad_sum(sum1);
      const float inv0 = sum0 > 0.f ? 1.0f / sum0 : 0.f, inv1 = sum1 > 0.f ? 1.0f / sum1 : 0.f;
      if ((lane & 3) == 0) { s_max[r0] = mx0; s_inv[r0] = inv0; s_max[r1] = mx1; s_inv[r1] = inv1; }
      // dP = dO_strip * V^T
      uint32_t ad[2][4];
      am_load_a(tdo_s, mt, lane, ad);
      float dp[8][4];
      am_strip_nt(dp, ad, tv_s, lane);
      const float dl0 = s_delta[r0], dl1 = s_delta[r1];
      uint32_t ds[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float v[4];
        v[0] = acc[nt][0] * inv0 * (dp[nt][0] - dl0); v[1] = acc[nt][1] * inv0 * (dp[nt][1] - dl0);
        v[2] = acc[nt][2] * inv1 * (dp[nt][2] - dl1); v[3] = acc[nt][3] * inv1 * (dp[nt][3] - dl1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int lc = cols.lin[nt * 2 + (e & 1)], lr = (e >> 1) ? lr1 : lr0;
          if (lc >= 0 && lr >= 0) atomicAdd(&s_bins[(lr & 0xffff) - (lc & 0xffff) + g.lin_off], v[e]);
        }
        ds[nt >> 1][(nt & 1) * 2] = pack_bf16(v[0], v[1]);
        ds[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(v[2], v[3]);
      }
      float dq[4][4];
      am_strip_pv(dq, ds, tk_s, lane);
      am_store_strip(dq, g.scale, g.scale, stage_s, stage, lane, mt, g.N, s_src, dqkv, 3 * g.C, h * 32, nullptr);
    }
    __syncwarp();

    // ---------------- pass B: column strips (keys): P^T, dP^T, dS^T -> dV, dK ----------------
    float cmax[16], cinv[16], cdel[16];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = nt * 8 + 2 * (lane & 3) + e;
        const bool ok = i < g.N;
        cmax[nt * 2 + e] = ok ? s_max[i] : 0.f; cinv[nt * 2 + e] = ok ? s_inv[i] : 0.f; cdel[nt * 2 + e] = ok ? s_delta[i] : 0.f;
      }
    for (int jt = 0; jt < n_mt; ++jt) {
      uint32_t a[2][4];
      am_load_a(tk_s, jt, lane, a);
      float acc[8][4];
      am_strip_nt(acc, a, tq_s, lane);          // S^T[j][i] = k_j . q_i
      const int r0 = jt * 16 + gq, r1 = r0 + 8;
      int lr0 = -1, lr1 = -1;
      if (r0 < g.N) { const int ty = r0 / g.ww, tx = r0 - ty * g.ww; lr0 = (ty * g.lin_stride + tx) | (s_reg[r0] << 16); }
      if (r1 < g.N) { const int ty = r1 / g.ww, tx = r1 - ty * g.ww; lr1 = (ty * g.lin_stride + tx) | (s_reg[r1] << 16); }
      am_scores<true>(acc, g, s_tbl, cols, lr0, lr1);
      uint32_t pt[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = nt * 2 + (e & 1);
          acc[nt][e] = exp2f((acc[nt][e] - cmax[c]) * LOG2E) * cinv[c];   // -inf -> 0
        }
        pt[nt >> 1][(nt & 1) * 2] = pack_bf16(acc[nt][0], acc[nt][1]);
        pt[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(acc[nt][2], acc[nt][3]);
      }
      float dv[4][4];
      am_strip_pv(dv, pt, tdo_s, lane);          // dV[j][d] = sum_i P[i][j] dO[i][d]
      am_store_strip(dv, 1.f, 1.f, stage_s, stage, lane, jt, g.N, s_src, dqkv, 3 * g.C, 2 * g.C + h * 32, dqkv_bias);
      uint32_t av[2][4];
      am_load_a(tv_s, jt, lane, av);
      float dp[8][4];
      am_strip_nt(dp, av, tdo_s, lane);          // dP^T[j][i] = v_j . dO_i
      uint32_t dst[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = acc[nt][e] * (dp[nt][e] - cdel[nt * 2 + (e & 1)]);
        dst[nt >> 1][(nt & 1) * 2] = pack_bf16(v[0], v[1]);
        dst[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(v[2], v[3]);
      }
      float dk[4][4];
      am_strip_pv(dk, dst, tq_s, lane);          // dK[j][d] = scale * sum_i dS[i][j] q[i][d]
      am_store_strip(dk, g.scale, g.scale, stage_s, stage, lane, jt, g.N, s_src, dqkv, 3 * g.C, g.C + h * 32, dqkv_bias);
    }
  }
  __syncwarp();
  for (int t = lane; t < g.ntab; t += 32) { const float v = s_bins[t]; if (v != 0.f) atomicAdd(dtable + t * g.heads + h, v); }
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
  g.windows = B * g.nwy * g.nwx;
  return MTUS_OK;
}

bool mtus_window_attn_mma_supported(int wh, int ww, int dtype) { return dtype == MTUS_BF16 && wh * ww <= 64 && wh > 0 && ww > 0; }

static int am_plan(AmGeom& g, int& blocks, int& n_items, int target_warps) {
  // enough warps to fill the machine a few times over; more windows per warp = fewer global bin flushes (bwd)
  int wpw = 1;
  while ((int64_t)((g.windows + wpw * 2 - 1) / (wpw * 2)) * g.heads >= target_warps && wpw < 16) wpw *= 2;
  g.win_per_warp = wpw;
  const int groups = (g.windows + wpw - 1) / wpw;
  n_items = groups * g.heads;
  blocks = (n_items + AM_WARPS - 1) / AM_WARPS;
  return MTUS_OK;
}

int mtus_window_attn_mma_fwd(const void* qkv, const float* rel_table, const float* qkv_bias, void* out, int B, int H, int W, int C,
                             int heads, int win_h, int win_w, int shift_h, int shift_w, cudaStream_t st) {
  AmGeom g;
  int rc = am_geom(g, B, H, W, C, heads, win_h, win_w, shift_h, shift_w);
  if (rc) return rc;
  if ((g.Hp != H || g.Wp != W) && !qkv_bias) return MTUS_ERR_BAD_ARG;
  if (B == 0) return MTUS_OK;
  int blocks, n_items;
  am_plan(g, blocks, n_items, 148 * 16 * 2);
  const size_t smem = am_smem_bytes(3, g.ntab, false);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_mma_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  window_attn_mma_fwd_kernel<<<blocks, AM_WARPS * 32, smem, st>>>((const bf16*)qkv, rel_table, qkv_bias, (bf16*)out, g, n_items);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

int mtus_window_attn_mma_bwd(const void* dout, const void* qkv, const void* out, const float* rel_table, const float* qkv_bias,
                             void* dqkv, float* drel_table, float* dqkv_bias, int B, int H, int W, int C, int heads, int win_h,
                             int win_w, int shift_h, int shift_w, cudaStream_t st) {
  AmGeom g;
  int rc = am_geom(g, B, H, W, C, heads, win_h, win_w, shift_h, shift_w);
  if (rc) return rc;
  if ((g.Hp != H || g.Wp != W) && !(qkv_bias && dqkv_bias)) return MTUS_ERR_BAD_ARG;
  if (B == 0) return MTUS_OK;
  int blocks, n_items;
  am_plan(g, blocks, n_items, 148 * 12 * 2);
  const size_t smem = am_smem_bytes(4, g.ntab, true);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_mma_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  window_attn_mma_bwd_kernel<<<blocks, AM_WARPS * 32, smem, st>>>((const bf16*)dout, (const bf16*)qkv, (const bf16*)out, rel_table, qkv_bias,
                                                                  (bf16*)dqkv, drel_table, dqkv_bias, g, n_items);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}
