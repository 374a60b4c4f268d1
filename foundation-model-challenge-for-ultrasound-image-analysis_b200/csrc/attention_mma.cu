// Fused shifted-window attention on tensor cores, bf16 storage / fp32 softmax (windows of up to 64 tokens).
//
// Same contract as attention.cu (timm _attn + WindowAttention minus the two Linear layers; SURVEY section 8a rows
// a5, a6): the cyclic shift, zero padding, window partition / reverse, relative-position bias, shift mask,
// softmax and P@V happen inside the kernel and nothing window-shaped touches HBM.
//
// Execution model (v3, persistent): one CTA of 4 warps owns ONE head and walks a strided list of windows
// (grid = heads x G, one wave of CTAs).  Everything that depends only on the head is set up once per CTA:
//   * the relative-position bias is expanded once into per-thread accumulator-fragment order in shared memory
//     (log2 domain, -inf baked in for the columns / rows beyond the window), so a score is one FMA;
//   * the bias-table gradient is accumulated in registers over all windows of the CTA and binned once at the end
//     (a [64 x 64] matrix in shared memory summed along its 169 diagonals -- no shared-memory atomics);
//   * the qkv-bias gradient (column sums of dqkv) is accumulated in registers and flushed once per CTA.
// Per window the q / k / v (/ dO / O) head slices [64 x 32] bf16 are gathered with 16-byte cp.async (zero fill for
// absent tokens) into a 2-stage ring: the loads of window i+1 are in flight while window i is computed.  Warp w
// owns the 16-token strip w; every contraction runs on the tensor cores (mma.sync m16n8k16, fp32 accumulate), the
// softmax works on accumulator fragments with exp2, outputs leave through a 1 KB staging strip per warp as 64-byte
// row segments.  Forward saves the per-(token, head) log-sum-exp; backward walks KEY strips (S^T and dP^T come out
// of the MMAs in the layout dV and dK consume as A operands), parks dS^T (bf16, 8 KB) in shared memory and then
// each warp contracts its own QUERY strip of dS with K for dQ.
//
// A stand-alone attention kernel is HBM-bound (24.5 FLOP/B at 49 tokens): algorithmic bytes per token are
// 4*C*2 forward (q, k, v in; o out) and 8*C*2 backward (q, k, v, o, dO in; dq, dk, dv out).  The per-window
// contractions are too small (49x32x49) for a tcgen05 tile on their own; they move to tcgen05 when this kernel is
// fused with the qkv / proj GEMMs ("what comes next").
#include "common.cuh"

#include "attention_mma.cuh"

// ---- persistent CTA-cooperative kernels ---------------------------------------------------------------------------
// shared memory per CTA (bytes):
//   fwd: ring [2][3 tiles: q,k,v][4096] | stage strips [4][1024] | bias fragments [4][8][32] float4 | int tables
//   bwd: ring [2][5 tiles: q,k,v,dO,O][4096] | stage strips | dS^T [64 keys][64 queries] bf16 | bias fragments | int tables |
//        lse ring [2][64] | delta[64] | column-sum scratch [4][96]
#define AM_FRAG_BYTES (AM_WARPS * 8 * 32 * 16)
#define AM_G_LD 68                                     // row pitch (floats) of the [64][64] dS^T-sum matrix (end of bwd)
#ifndef AM_FWD_OCC
#define AM_FWD_OCC 4
#endif
#ifndef AM_BWD_OCC
#define AM_BWD_OCC 3
#endif
static __host__ __device__ inline int am_cta_bytes(bool bwd) {
  int b = 2 * (bwd ? 5 : 3) * AM_TILE_BYTES + AM_WARPS * 1024 + AM_FRAG_BYTES + 5 * 256;   // s_pos, s_lin, s_rg, s_src[2]
  if (bwd) b += 64 * 128 + 2 * 256 + 256 + AM_WARPS * 96 * 4;
  return b;
}

// Relative-position bias (log2 domain) in accumulator-fragment order: entry (warp, nt, lane) holds the four values of
// acc[nt][0..3] of strip `warp`: rows r0 = 16 warp + lane/4, r1 = r0 + 8; columns 8 nt + 2 (lane % 4) + {0, 1}.
// Columns >= N are -inf (softmax ignores them); rows >= N are -inf in the transposed (backward, rows = keys) form
// so that P^T is exactly zero there, and 0 in the forward form (those query rows are never stored).
template <bool TRANSPOSED>
__device__ __forceinline__ void am_build_bias_frags(float4* s_frag, const float* s_tbl2, const int* s_lin, const AmGeom& g, int tid) {
  for (int e = tid; e < AM_WARPS * 8 * 32; e += AM_WARPS * 32) {
    const int lane = e & 31, nt = (e >> 5) & 7, w = e >> 8;
    const int r0 = w * 16 + (lane >> 2), r1 = r0 + 8, c0 = nt * 8 + 2 * (lane & 3), c1 = c0 + 1;
    auto val = [&](int r, int c) -> float {
      if (c >= g.N) return -INFINITY;
      if (r >= g.N) return TRANSPOSED ? -INFINITY : 0.f;
      return s_tbl2[(TRANSPOSED ? (s_lin[c] - s_lin[r]) : (s_lin[r] - s_lin[c])) + g.lin_off];
    };
    s_frag[e] = make_float4(val(r0, c0), val(r0, c1), val(r1, c0), val(r1, c1));
  }
}

template <int NTC, bool MASKED>
__device__ __forceinline__ void am_scores_frag(float (&acc)[8][4], float scale2, const float4* frag, const int (&creg)[16], int rr0, int rr1) {
  const float MASK2 = -100.0f * AM_LOG2E;
#pragma unroll
  for (int nt = 0; nt < NTC; ++nt) {
    const float4 b = frag[nt * 32];
    acc[nt][0] = fmaf(acc[nt][0], scale2, b.x); acc[nt][1] = fmaf(acc[nt][1], scale2, b.y);
    acc[nt][2] = fmaf(acc[nt][2], scale2, b.z); acc[nt][3] = fmaf(acc[nt][3], scale2, b.w);
    if (MASKED) {
      if (creg[nt * 2] != rr0) acc[nt][0] += MASK2;
      if (creg[nt * 2 + 1] != rr0) acc[nt][1] += MASK2;
      if (creg[nt * 2] != rr1) acc[nt][2] += MASK2;
      if (creg[nt * 2 + 1] != rr1) acc[nt][3] += MASK2;
    }
  }
}

// Issues the gathers of one window into ring stage `ring_s` (NT tiles: q, k, v [, dO, O]); thread (tok, ch) moves the
// 16-byte chunk ch of tokens tok and tok + 32.  Also records the source rows (s_src) and, for the backward pass,
// prefetches the log-sum-exp of the query rows (1e30 for absent queries: P = 0).
template <bool BWD>
__device__ __forceinline__ void am_issue_window(const AmGeom& g, const AmWin& win, int h, uint32_t ring_s, const bf16* qkv, const bf16* dout,
                                                const bf16* outp, const float* lse, const int* s_pos, int* s_src, float* s_lse, int tid) {
  const int ch = tid & 3;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int tok = k * 32 + (tid >> 2);
    const int src = am_source(g, win, tok, s_pos);
    const int bytes = src < 0 ? 0 : 16;
    const int64_t row = src < 0 ? 0 : src;
    const uint32_t off = am_off(tok, ch);
    const bf16* p = qkv + row * (3 * g.C) + h * 32 + ch * 8;
    cp_async16_zfill(ring_s + off, p, bytes);
    cp_async16_zfill(ring_s + AM_TILE_BYTES + off, p + g.C, bytes);
    cp_async16_zfill(ring_s + 2 * AM_TILE_BYTES + off, p + 2 * g.C, bytes);
    if (BWD) {
      cp_async16_zfill(ring_s + 3 * AM_TILE_BYTES + off, dout + row * g.C + h * 32 + ch * 8, bytes);   // padded tokens: dO = 0 (cropped away)
      cp_async16_zfill(ring_s + 4 * AM_TILE_BYTES + off, outp + row * g.C + h * 32 + ch * 8, bytes);
    }
    if (ch == 0) {
      s_src[tok] = src;
      if (BWD) {
        if (src >= 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32_generic(s_lse + tok)), "l"(lse + (int64_t)src * g.heads + h) : "memory");
        else s_lse[tok] = 1e30f;
      }
    }
  }
}

template <int NTC>
__global__ void __launch_bounds__(AM_WARPS * 32, AM_FWD_OCC) window_attn_mma_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ table,
                                                                                        const float* __restrict__ qkv_bias, bf16* __restrict__ out,
                                                                                        float* __restrict__ lse, AmGeom g) {
  extern __shared__ __align__(128) uint8_t am_smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % g.heads, grp = blockIdx.x / g.heads;
  uint8_t* ring = am_smem_raw;                                             // [2][3][4096]
  uint8_t* stage = ring + 2 * 3 * AM_TILE_BYTES + warp * 1024;
  float4* s_frag = reinterpret_cast<float4*>(ring + 2 * 3 * AM_TILE_BYTES + AM_WARPS * 1024);
  int* s_pos = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_frag) + AM_FRAG_BYTES);
  int* s_lin = s_pos + 64; int* s_rg = s_lin + 64; int* s_src2 = s_rg + 64;   // s_src2: [2][64]
  float* s_tbl = reinterpret_cast<float*>(ring + 3 * AM_TILE_BYTES);       // scratch inside ring stage 1 (free until the first prefetch)
  const uint32_t ring_s = smem_u32_generic(ring), stage_s = smem_u32_generic(stage);
  for (int t = tid; t < g.ntab; t += AM_WARPS * 32) s_tbl[t] = __ldg(table + t * g.heads + h) * AM_LOG2E;
  am_init_tables(g, tid, s_pos, s_lin, s_rg);
  pdl_trigger();
  __syncthreads();
  pdl_wait();                                           // table / index set-up overlapped the previous kernel's tail
  int w = grp;
  if (w < g.windows) am_issue_window<false>(g, am_window(g, w), h, ring_s, qkv, nullptr, nullptr, nullptr, s_pos, s_src2, nullptr, tid);
  am_build_bias_frags<false>(s_frag, s_tbl, s_lin, g, tid);
  const float4* frag = s_frag + warp * 8 * 32 + lane;
  const int gq = lane >> 2, tq4 = lane & 3;
  const int n_mt = (g.N + 15) >> 4;
  const int mt = warp;                                  // this warp's query strip
  const int r0 = mt * 16 + gq, r1 = r0 + 8;
  int creg[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) creg[c] = 0;
  int st = 0;
  for (; w < g.windows; w += g.groups, st ^= 1) {
    const AmWin win = am_window(g, w);
    const bool masked = win.last_row || win.last_col;
    const bool has_pad = am_has_pad(g, win);
    cp_async_wait_all();
    __syncthreads();                                    // tiles of this window landed; every warp is done with the previous one
    if (w + g.groups < g.windows)
      am_issue_window<false>(g, am_window(g, w + g.groups), h, ring_s + (st ^ 1) * 3 * AM_TILE_BYTES, qkv, nullptr, nullptr, nullptr, s_pos,
                             s_src2 + (st ^ 1) * 64, nullptr, tid);
    uint8_t* tq = ring + st * 3 * AM_TILE_BYTES;
    const uint32_t tq_s = ring_s + st * 3 * AM_TILE_BYTES, tk_s = tq_s + AM_TILE_BYTES, tv_s = tk_s + AM_TILE_BYTES;
    const int* s_src = s_src2 + st * 64;
    if (has_pad) {
      am_fill_pad_rows(tq, qkv_bias, h * 32, s_src, g.N, tid);
      am_fill_pad_rows(tq + AM_TILE_BYTES, qkv_bias, g.C + h * 32, s_src, g.N, tid);
      am_fill_pad_rows(tq + 2 * AM_TILE_BYTES, qkv_bias, 2 * g.C + h * 32, s_src, g.N, tid);
      __syncthreads();
    }
    if (mt < n_mt) {
      if (masked) {
#pragma unroll
        for (int c = 0; c < 16; ++c) { const int j = (c >> 1) * 8 + 2 * tq4 + (c & 1); creg[c] = am_region(win, s_rg[j]); }
      }
      uint32_t a[2][4];
      am_load_a(tq_s, mt, lane, a);
      float acc[8][4];
      am_strip_nt<NTC>(acc, a, tk_s, lane);
      if (masked) am_scores_frag<NTC, true>(acc, g.scale2, frag, creg, am_region(win, s_rg[r0]), am_region(win, s_rg[r1]));
      else am_scores_frag<NTC, false>(acc, g.scale2, frag, creg, 0, 0);
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
__global__ void __launch_bounds__(AM_WARPS * 32, AM_BWD_OCC) window_attn_mma_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ qkv,
                                                                                        const bf16* __restrict__ outp, const float* __restrict__ lse,
                                                                                        const float* __restrict__ table, const float* __restrict__ qkv_bias,
                                                                                        bf16* __restrict__ dqkv, float* __restrict__ dtable,
                                                                                        float* __restrict__ dqkv_bias, float* __restrict__ dqkv_colsum,
                                                                                        AmGeom g) {
  extern __shared__ __align__(128) uint8_t am_smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % g.heads, grp = blockIdx.x / g.heads;
  uint8_t* ring = am_smem_raw;                                             // [2][5][4096]
  uint8_t* stage = ring + 2 * 5 * AM_TILE_BYTES + warp * 1024;
  uint8_t* sds_all = ring + 2 * 5 * AM_TILE_BYTES + AM_WARPS * 1024;       // dS^T [64 keys][64 queries] bf16, 16-byte chunks swizzled by key row
  float4* s_frag = reinterpret_cast<float4*>(sds_all + 64 * 128);
  int* s_pos = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_frag) + AM_FRAG_BYTES);
  int* s_lin = s_pos + 64; int* s_rg = s_lin + 64; int* s_src2 = s_rg + 64;   // s_src2: [2][64]
  float* s_lse2 = reinterpret_cast<float*>(s_src2 + 128);                  // [2][64]
  float* s_delta = s_lse2 + 128;
  float* s_cs = s_delta + 64;                                              // [4 warps][3][32]
  float* s_tbl = reinterpret_cast<float*>(ring + 5 * AM_TILE_BYTES);       // scratch inside ring stage 1 (free until the first prefetch)
  const uint32_t ring_s = smem_u32_generic(ring), stage_s = smem_u32_generic(stage), sds_all_s = smem_u32_generic(sds_all);
  const uint32_t sds_s = sds_all_s + warp * 2048;                          // this warp's dS^T strip [16 keys][64 queries]
  for (int t = tid; t < g.ntab; t += AM_WARPS * 32) s_tbl[t] = __ldg(table + t * g.heads + h) * AM_LOG2E;
  am_init_tables(g, tid, s_pos, s_lin, s_rg);
  pdl_trigger();
  __syncthreads();
  pdl_wait();                                           // table / index set-up overlapped the previous kernel's tail
  int w = grp;
  if (w < g.windows) am_issue_window<true>(g, am_window(g, w), h, ring_s, qkv, dout, outp, lse, s_pos, s_src2, s_lse2, tid);
  am_build_bias_frags<true>(s_frag, s_tbl, s_lin, g, tid);
  const float4* frag = s_frag + warp * 8 * 32 + lane;
  const int gq = lane >> 2, tq4 = lane & 3;
  const int n_mt = (g.N + 15) >> 4;
  const int jt = warp;                                  // this warp's key strip (and its query strip for dQ)
  const int r0 = jt * 16 + gq, r1 = r0 + 8;
  int creg[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) creg[c] = 0;
  float gsum[8][4];                     // dS^T of this warp's strip summed over the CTA's windows (bias-table gradient)
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) gsum[nt][0] = gsum[nt][1] = gsum[nt][2] = gsum[nt][3] = 0.f;
  float cs_k[8], cs_v[8], cs_q[8];      // qkv-bias gradient partials: columns 8*nt + 2*t + {0,1} of this head
#pragma unroll
  for (int k = 0; k < 8; ++k) cs_k[k] = cs_v[k] = cs_q[k] = 0.f;
  int st = 0;
  for (; w < g.windows; w += g.groups, st ^= 1) {
    const AmWin win = am_window(g, w);
    const bool masked = win.last_row || win.last_col;
    const bool has_pad = am_has_pad(g, win);
    cp_async_wait_all();
    __syncthreads();                                    // S1: tiles of this window landed; every warp is done with the previous one
    if (w + g.groups < g.windows)
      am_issue_window<true>(g, am_window(g, w + g.groups), h, ring_s + (st ^ 1) * 5 * AM_TILE_BYTES, qkv, dout, outp, lse, s_pos,
                            s_src2 + (st ^ 1) * 64, s_lse2 + (st ^ 1) * 64, tid);
    uint8_t* tq = ring + st * 5 * AM_TILE_BYTES;
    const uint32_t tq_s = ring_s + st * 5 * AM_TILE_BYTES, tk_s = tq_s + AM_TILE_BYTES, tv_s = tk_s + AM_TILE_BYTES, tdo_s = tv_s + AM_TILE_BYTES;
    const int* s_src = s_src2 + st * 64;
    const float2* lse2 = reinterpret_cast<const float2*>(s_lse2 + st * 64) + tq4;
    const float2* del2 = reinterpret_cast<const float2*>(s_delta) + tq4;
    if (has_pad) {
      am_fill_pad_rows(tq, qkv_bias, h * 32, s_src, g.N, tid);
      am_fill_pad_rows(tq + AM_TILE_BYTES, qkv_bias, g.C + h * 32, s_src, g.N, tid);
      am_fill_pad_rows(tq + 2 * AM_TILE_BYTES, qkv_bias, 2 * g.C + h * 32, s_src, g.N, tid);
    }
    // delta_i = dO_i . O_i from the shared tiles (4 lanes per token, 16 bytes each)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int tok = k * 32 + (tid >> 2), ch = tid & 3;
      const uint4 a = *reinterpret_cast<const uint4*>(tq + 3 * AM_TILE_BYTES + am_off(tok, ch));
      const uint4 b = *reinterpret_cast<const uint4*>(tq + 4 * AM_TILE_BYTES + am_off(tok, ch));
      const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 fa = __bfloat1622float2(ha[i]), fb = __bfloat1622float2(hb[i]); part = fmaf(fa.x, fb.x, part); part = fmaf(fa.y, fb.y, part); }
      part = quad_sum(part);
      if (ch == 0) s_delta[tok] = part;
    }
    __syncthreads();                                    // S2: delta (and pad rows) visible

    if (jt < n_mt) {
      if (masked) {
#pragma unroll
        for (int c = 0; c < 16; ++c) { const int j = (c >> 1) * 8 + 2 * tq4 + (c & 1); creg[c] = am_region(win, s_rg[j]); }
      }
      uint32_t a[2][4];
      am_load_a(tk_s, jt, lane, a);
      float acc[8][4];
      am_strip_nt<NTC>(acc, a, tq_s, lane);          // S^T[j][i] = k_j . q_i
      if (masked) am_scores_frag<NTC, true>(acc, g.scale2, frag, creg, am_region(win, s_rg[r0]), am_region(win, s_rg[r1]));
      else am_scores_frag<NTC, false>(acc, g.scale2, frag, creg, 0, 0);
      uint32_t pt[4][4];
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) {
        const float2 ls = lse2[nt * 4];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = ex2f(acc[nt][e] - ((e & 1) ? ls.y : ls.x));   // -inf / lse 1e30 -> 0
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
      // dS^T strip -> shared [16 keys][64 queries] (16-byte chunks swizzled by key row)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int c0 = ks * 2, c1 = ks * 2 + 1;   // 16-byte chunk index = query column / 8
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + gq * 128 + ((c0 ^ (gq & 7)) << 4) + tq4 * 4), "r"(dst[ks][0]) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + (gq + 8) * 128 + ((c0 ^ ((gq + 8) & 7)) << 4) + tq4 * 4), "r"(dst[ks][1]) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + gq * 128 + ((c1 ^ (gq & 7)) << 4) + tq4 * 4), "r"(dst[ks][2]) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + (gq + 8) * 128 + ((c1 ^ ((gq + 8) & 7)) << 4) + tq4 * 4), "r"(dst[ks][3]) : "memory");
      }
    }
    __syncthreads();                                    // S3: dS^T complete
    if (jt < n_mt) {
      // dQ strip jt = scale * dS[16 queries][keys] * K: A = transposed loads from dS^T, B = transposed loads from the K tile
      float dq[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
      const int mi = lane >> 3;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (ks < n_mt) {
          uint32_t a0, a1, a2, a3;
          const int key = (mi >> 1) * 8 + (lane & 7), chunk = jt * 2 + (mi & 1);
          ldsm_x4_t(sds_all_s + ks * 2048 + key * 128 + ((chunk ^ (key & 7)) << 4), a0, a1, a2, a3);
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(tk_s + am_off(ks * 16 + (mi & 1) * 8 + (lane & 7), np * 2 + (mi >> 1)), b0, b1, b2, b3);
            mma_bf16(dq[np * 2], a0, a1, a2, a3, b0, b1);
            mma_bf16(dq[np * 2 + 1], a0, a1, a2, a3, b2, b3);
          }
        }
      }
      am_store_strip(dq, g.scale, g.scale, stage_s, stage, lane, jt, g.N, s_src, dqkv, 3 * g.C, h * 32, nullptr, dqkv_colsum ? cs_q : nullptr);
    }
  }
  // ---- flush (once per CTA): bias-table gradient and qkv-bias gradient ----
  __syncthreads();
  float* Gm = reinterpret_cast<float*>(ring);           // [64 keys][AM_G_LD] sums of dS^T over this CTA's windows
  if (jt < n_mt) {
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) {
      *reinterpret_cast<float2*>(Gm + r0 * AM_G_LD + nt * 8 + 2 * tq4) = make_float2(gsum[nt][0], gsum[nt][1]);
      *reinterpret_cast<float2*>(Gm + r1 * AM_G_LD + nt * 8 + 2 * tq4) = make_float2(gsum[nt][2], gsum[nt][3]);
    }
  }
  if (dqkv_colsum) {
    // columns 8*nt + 2*t + {0,1} per lane: sum over the 8 row groups (lanes with the same t), then over the warps
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float q = cs_q[k], kk = cs_k[k], v = cs_v[k];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) { q += __shfl_xor_sync(0xffffffffu, q, o); kk += __shfl_xor_sync(0xffffffffu, kk, o); v += __shfl_xor_sync(0xffffffffu, v, o); }
      if (gq == 0) {
        const int col = (k >> 1) * 8 + 2 * tq4 + (k & 1);
        s_cs[warp * 96 + col] = q; s_cs[warp * 96 + 32 + col] = kk; s_cs[warp * 96 + 64 + col] = v;
      }
    }
  }
  __syncthreads();
  // bin t = (dy + wh - 1) * (2 ww - 1) + (dx + ww - 1) collects dS[i][j] over all pairs with (yi - yj, xi - xj) = (dy, dx)
  for (int t = tid; t < g.ntab; t += AM_WARPS * 32) {
    const int dy = t / g.lin_stride - (g.wh - 1), dx = t % g.lin_stride - (g.ww - 1);
    float v = 0.f;
    for (int yj = max(0, -dy); yj < min(g.wh, g.wh - dy); ++yj)
      for (int xj = max(0, -dx); xj < min(g.ww, g.ww - dx); ++xj)
        v += Gm[(yj * g.ww + xj) * AM_G_LD + (yj + dy) * g.ww + xj + dx];
    if (v != 0.f) atomicAdd(dtable + t * g.heads + h, v);
  }
  if (dqkv_colsum && tid < 96) {
    const float v = s_cs[tid] + s_cs[96 + tid] + s_cs[192 + tid] + s_cs[288 + tid];
    atomicAdd(dqkv_colsum + (tid >> 5) * g.C + h * 32 + (tid & 31), v);
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
  g.groups = 1;
  return MTUS_OK;
}

bool mtus_window_attn_mma_supported(int wh, int ww, int dtype) { return dtype == MTUS_BF16 && wh * ww <= 64 && wh > 0 && ww > 0; }

static int am_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n = v > 0 ? v : 148;
  }
  return n;
}

// one wave of persistent CTAs: G CTAs per head, each walking windows g, g + G, ...; G is the largest count that fits
// the machine, then reduced to the smallest count with the same number of windows per CTA (fewer set-ups / flushes)
static int am_plan(AmGeom& g, int ctas_per_sm) {
  int G = (am_sm_count() * ctas_per_sm) / g.heads;
  if (G < 1) G = 1;
  if (G > g.windows) G = g.windows;
  const int per = (g.windows + G - 1) / G;
  G = (g.windows + per - 1) / per;
  g.groups = G;
  return G * g.heads;
}

int mtus_window_attn_mma_fwd(const void* qkv, const float* rel_table, const float* qkv_bias, void* out, float* lse, int B, int H, int W,
                             int C, int heads, int win_h, int win_w, int shift_h, int shift_w, cudaStream_t st) {
  AmGeom g;
  int rc = am_geom(g, B, H, W, C, heads, win_h, win_w, shift_h, shift_w);
  if (rc) return rc;
  if ((g.Hp != H || g.Wp != W) && !qkv_bias) return MTUS_ERR_BAD_ARG;
  if (B == 0) return MTUS_OK;
  const int blocks = am_plan(g, AM_FWD_OCC);
  const size_t smem = am_cta_bytes(false);
  static mtus_per_device_flag configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_mma_fwd_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attn_mma_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured.set();
  }
  cudaError_t le;
  if (g.N <= 56) le = mtus_launch_pdl(window_attn_mma_fwd_kernel<7>, dim3(blocks), dim3(AM_WARPS * 32), smem, st, (const bf16*)qkv, rel_table, qkv_bias, (bf16*)out, lse, g);
  else le = mtus_launch_pdl(window_attn_mma_fwd_kernel<8>, dim3(blocks), dim3(AM_WARPS * 32), smem, st, (const bf16*)qkv, rel_table, qkv_bias, (bf16*)out, lse, g);
  if (le != cudaSuccess) return (int)le;
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
  const int blocks = am_plan(g, AM_BWD_OCC);
  const size_t smem = am_cta_bytes(true);
  static mtus_per_device_flag configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_mma_bwd_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attn_mma_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured.set();
  }
  cudaError_t le;
  if (g.N <= 56)
    le = mtus_launch_pdl(window_attn_mma_bwd_kernel<7>, dim3(blocks), dim3(AM_WARPS * 32), smem, st, (const bf16*)dout, (const bf16*)qkv, (const bf16*)out, lse,
                         rel_table, qkv_bias, (bf16*)dqkv, drel_table, dqkv_bias, dqkv_colsum, g);
  else
    le = mtus_launch_pdl(window_attn_mma_bwd_kernel<8>, dim3(blocks), dim3(AM_WARPS * 32), smem, st, (const bf16*)dout, (const bf16*)qkv, (const bf16*)out, lse,
                         rel_table, qkv_bias, (bf16*)dqkv, drel_table, dqkv_bias, dqkv_colsum, g);
  if (le != cudaSuccess) return (int)le;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}
