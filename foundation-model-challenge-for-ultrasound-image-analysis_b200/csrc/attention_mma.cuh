// Device helpers shared by the tensor-core window-attention kernels (attention_mma.cu: windows of up to 64 tokens;
// attention_mma144.cu: up to 144 tokens, i.e. window 12): swizzled [tokens][32] bf16 tiles, cp.async gathers,
// ldmatrix / mma.sync wrappers, window geometry (cyclic shift, padding, mask regions).
#pragma once
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
  int groups;             // G: CTAs per head; CTA (g, h) handles windows g, g + G, g + 2G, ...
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
// source row of token t of window w (cyclic shift undone), or -1 for padding / t >= N
__device__ __forceinline__ int am_source(const AmGeom& g, const AmWin& w, int t, const int* s_pos) {
  if (t >= g.N) return -1;
  const int pos = s_pos[t];
  const int py = w.wy * g.wh + (pos & 0xff), px = w.wx * g.ww + (pos >> 8);
  if (py >= g.H || px >= g.W) return -1;
  int y = py + g.sh, x = px + g.sw;
  if (y >= g.H) y -= g.H;
  if (x >= g.W) x -= g.W;
  return (w.b * g.H + y) * g.W + x;
}

__device__ __forceinline__ int am_region(const AmWin& w, int rg) {
  return (w.last_row ? (rg & 0xff) : 0) + (w.last_col ? (rg >> 8) : 0);
}

