// Fused shifted-window attention on tensor cores for windows of 65..144 tokens (window 12: swin_*_window12_384,
// BASELINE configs[3]), bf16 storage / fp32 softmax.  Same contract and the same building blocks as attention_mma.cu
// (attention_mma.cuh): persistent CTAs that own one head and walk a strided list of windows, 16-byte cp.async gathers
// with the cyclic shift / padding / partition folded into the addresses, mma.sync m16n8k16 contractions, exp2 softmax on
// accumulator fragments, log-sum-exp saved by forward, bias-table gradient accumulated on chip and binned once per CTA.
// What changes with 144 tokens:
//   * 9 warps per CTA (warp w = 16-token strip w), tiles [144][32] bf16 = 9 KB, one CTA per SM;
//   * forward: scores of a strip are 16 x 144 (18 accumulator tiles), bias in fragment order in shared memory (81 KB),
//     2-stage gather ring;
//   * backward: the 144 queries of a key strip are processed in two slabs (80 + 64 columns) so S^T and dP^T stay in
//     registers; dV / dK accumulate across the slabs; dS^T goes to shared memory ([144 keys][144 queries] bf16) for the
//     per-query-strip dQ contraction, and is summed over the CTA's windows into an fp32 [144][144] matrix in shared
//     memory that is folded along its 529 diagonals at the end (bias-table gradient).  No gather ring (shared memory
//     is spent on those two matrices); the relative-position bias is looked up in the table per element.
#include "attention_mma.cuh"

#define A4_NW 9
#define A4_TOK 144
#define A4_THREADS (A4_NW * 32)
#define A4_NTC 18
#define A4_TILE (A4_TOK * 64)          // bytes of one [144][32] bf16 tile
#define A4_FRAG_BYTES (A4_NW * A4_NTC * 32 * 16)
#define A4_DS_PITCH 384                // bytes per key row of dS^T: 18 16-byte chunks, XOR-swizzled within 24
#define A4_G_LD 148                    // floats per row of the dS^T sum

static __host__ __device__ inline int a4_fwd_bytes() { return 2 * 3 * A4_TILE + A4_NW * 1024 + A4_FRAG_BYTES + 5 * A4_TOK * 4; }
static __host__ __device__ inline int a4_bwd_bytes(int ntab) {
  return 5 * A4_TILE + A4_NW * 1024 + A4_TOK * A4_DS_PITCH + A4_TOK * A4_G_LD * 4 + ((ntab * 4 + 15) & ~15) + 6 * A4_TOK * 4 +
         A4_NW * 96 * 4;
}

__device__ __forceinline__ void a4_init_tables(const AmGeom& g, int tid, int* s_pos, int* s_lin, int* s_rg) {
  for (int t = tid; t < A4_TOK; t += A4_THREADS) {
    int pos = 0, lin = 0, rg = 0;
    if (t < g.N) {
      const int ty = t / g.ww, tx = t - ty * g.ww;
      pos = ty | (tx << 8);
      lin = ty * g.lin_stride + tx;
      const int rh = (ty < g.wh - g.sh) ? 1 : 2, rw = (tx < g.ww - g.sw) ? 1 : 2;
      rg = (rh * 3) | (rw << 8);
    }
    s_pos[t] = pos; s_lin[t] = lin; s_rg[t] = rg;
  }
}

// gathers of one window into `tiles_s` (NT tiles: q, k, v [, dO, O]); thread (tok, ch) moves chunk ch of tokens tok, tok + 72
template <bool BWD>
__device__ __forceinline__ void a4_issue_window(const AmGeom& g, const AmWin& win, int h, uint32_t tiles_s, const bf16* qkv, const bf16* dout,
                                                const bf16* outp, const float* lse, const int* s_pos, int* s_src, float* s_lse, int tid) {
  const int ch = tid & 3;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int tok = k * 72 + (tid >> 2);
    const int src = am_source(g, win, tok, s_pos);
    const int bytes = src < 0 ? 0 : 16;
    const int64_t row = src < 0 ? 0 : src;
    const uint32_t off = am_off(tok, ch);
    const bf16* p = qkv + row * (3 * g.C) + h * 32 + ch * 8;
    cp_async16_zfill(tiles_s + off, p, bytes);
    cp_async16_zfill(tiles_s + A4_TILE + off, p + g.C, bytes);
    cp_async16_zfill(tiles_s + 2 * A4_TILE + off, p + 2 * g.C, bytes);
    if (BWD) {
      cp_async16_zfill(tiles_s + 3 * A4_TILE + off, dout + row * g.C + h * 32 + ch * 8, bytes);
      cp_async16_zfill(tiles_s + 4 * A4_TILE + off, outp + row * g.C + h * 32 + ch * 8, bytes);
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

__device__ __forceinline__ void a4_fill_pad_rows(uint8_t* tile_g, const float* pad_bias, int col0, const int* s_src, int N, int tid) {
  for (int it = tid; it < N * 4; it += A4_THREADS) {
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

// acc[nt] = A_strip[16 x 32] * T[(NT0 + nt) * 8 .. + 8][32]^T for nt < NTN
template <int NT0, int NTN>
__device__ __forceinline__ void a4_strip_nt(float (&acc)[NTN][4], const uint32_t (&a)[2][4], uint32_t tile, int lane) {
#pragma unroll
  for (int nt = 0; nt < NTN; ++nt) {
    acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    uint32_t b0, b1, b2, b3;
    ldsm_x4(tile + am_off((NT0 + nt) * 8 + (lane & 7), lane >> 3), b0, b1, b2, b3);
    mma_bf16(acc[nt], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
    mma_bf16(acc[nt], a[1][0], a[1][1], a[1][2], a[1][3], b2, b3);
  }
}

// out[nt] (4 channel tiles) += P[16 x 16 KSN] * T[ROW0 .. ROW0 + 16 KSN][32]; P = packed bf16 A fragments per 16-token k-step
template <int ROW0, int KSN>
__device__ __forceinline__ void a4_strip_pv(float (&out)[4][4], const uint32_t (&p)[KSN][4], uint32_t tile, int lane) {
  const int mi = lane >> 3;
#pragma unroll
  for (int ks = 0; ks < KSN; ++ks) {
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile + am_off(ROW0 + ks * 16 + (mi & 1) * 8 + (lane & 7), np * 2 + (mi >> 1)), b0, b1, b2, b3);
      mma_bf16(out[np * 2], p[ks][0], p[ks][1], p[ks][2], p[ks][3], b0, b1);
      mma_bf16(out[np * 2 + 1], p[ks][0], p[ks][1], p[ks][2], p[ks][3], b2, b3);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(A4_THREADS, 1) window_attn_mma144_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ table,
                                                                                const float* __restrict__ qkv_bias, bf16* __restrict__ out,
                                                                                float* __restrict__ lse, AmGeom g) {
  extern __shared__ __align__(128) uint8_t a4_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % g.heads, grp = blockIdx.x / g.heads;
  uint8_t* ring = a4_smem;                                           // [2][3][A4_TILE]
  uint8_t* stage = ring + 2 * 3 * A4_TILE + warp * 1024;
  float4* s_frag = reinterpret_cast<float4*>(ring + 2 * 3 * A4_TILE + A4_NW * 1024);
  int* s_pos = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_frag) + A4_FRAG_BYTES);
  int* s_lin = s_pos + A4_TOK; int* s_rg = s_lin + A4_TOK; int* s_src2 = s_rg + A4_TOK;   // s_src2: [2][A4_TOK]
  float* s_tbl = reinterpret_cast<float*>(ring + 3 * A4_TILE);       // scratch in ring stage 1 until the first prefetch
  const uint32_t ring_s = smem_u32_generic(ring), stage_s = smem_u32_generic(stage);
  for (int t = tid; t < g.ntab; t += A4_THREADS) s_tbl[t] = __ldg(table + t * g.heads + h) * AM_LOG2E;
  a4_init_tables(g, tid, s_pos, s_lin, s_rg);
  pdl_trigger();
  __syncthreads();
  pdl_wait();
  int w = grp;
  if (w < g.windows) a4_issue_window<false>(g, am_window(g, w), h, ring_s, qkv, nullptr, nullptr, nullptr, s_pos, s_src2, nullptr, tid);
  // bias in accumulator-fragment order: entry (warp, nt, lane) = acc[nt][0..3] of strip `warp` (log2 domain, -inf beyond N)
  for (int e = tid; e < A4_NW * A4_NTC * 32; e += A4_THREADS) {
    const int ln = e & 31, nt = (e >> 5) % A4_NTC, ws = e / (32 * A4_NTC);
    const int r0 = ws * 16 + (ln >> 2), r1 = r0 + 8, c0 = nt * 8 + 2 * (ln & 3), c1 = c0 + 1;
    auto val = [&](int r, int c) -> float {
      if (c >= g.N) return -INFINITY;
      if (r >= g.N) return 0.f;
      return s_tbl[s_lin[r] - s_lin[c] + g.lin_off];
    };
    s_frag[e] = make_float4(val(r0, c0), val(r0, c1), val(r1, c0), val(r1, c1));
  }
  const float4* frag = s_frag + warp * A4_NTC * 32 + lane;
  const int gq = lane >> 2, tq4 = lane & 3;
  const int n_mt = (g.N + 15) >> 4;
  const int mt = warp;
  const int r0 = mt * 16 + gq, r1 = r0 + 8;
  const float MASK2 = -100.0f * AM_LOG2E;
  int st = 0;
  for (; w < g.windows; w += g.groups, st ^= 1) {
    const AmWin win = am_window(g, w);
    const bool masked = win.last_row || win.last_col;
    const bool has_pad = am_has_pad(g, win);
    cp_async_wait_all();
    __syncthreads();
    if (w + g.groups < g.windows)
      a4_issue_window<false>(g, am_window(g, w + g.groups), h, ring_s + (st ^ 1) * 3 * A4_TILE, qkv, nullptr, nullptr, nullptr, s_pos,
                             s_src2 + (st ^ 1) * A4_TOK, nullptr, tid);
    uint8_t* tq = ring + st * 3 * A4_TILE;
    const uint32_t tq_s = ring_s + st * 3 * A4_TILE, tk_s = tq_s + A4_TILE, tv_s = tk_s + A4_TILE;
    const int* s_src = s_src2 + st * A4_TOK;
    if (has_pad) {
      a4_fill_pad_rows(tq, qkv_bias, h * 32, s_src, g.N, tid);
      a4_fill_pad_rows(tq + A4_TILE, qkv_bias, g.C + h * 32, s_src, g.N, tid);
      a4_fill_pad_rows(tq + 2 * A4_TILE, qkv_bias, 2 * g.C + h * 32, s_src, g.N, tid);
      __syncthreads();
    }
    if (mt < n_mt) {
      uint32_t a[2][4];
      am_load_a(tq_s, mt, lane, a);
      float acc[A4_NTC][4];
      a4_strip_nt<0, A4_NTC>(acc, a, tk_s, lane);
      const int rr0 = masked ? am_region(win, s_rg[r0]) : 0, rr1 = masked ? am_region(win, s_rg[r1]) : 0;
#pragma unroll
      for (int nt = 0; nt < A4_NTC; ++nt) {
        const float4 b = frag[nt * 32];
        acc[nt][0] = fmaf(acc[nt][0], g.scale2, b.x); acc[nt][1] = fmaf(acc[nt][1], g.scale2, b.y);
        acc[nt][2] = fmaf(acc[nt][2], g.scale2, b.z); acc[nt][3] = fmaf(acc[nt][3], g.scale2, b.w);
        if (masked) {
          const int cg0 = am_region(win, s_rg[nt * 8 + 2 * tq4]), cg1 = am_region(win, s_rg[nt * 8 + 2 * tq4 + 1]);
          if (cg0 != rr0) acc[nt][0] += MASK2;
          if (cg1 != rr0) acc[nt][1] += MASK2;
          if (cg0 != rr1) acc[nt][2] += MASK2;
          if (cg1 != rr1) acc[nt][3] += MASK2;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < A4_NTC; ++nt) { mx0 = fmaxf(mx0, fmaxf(acc[nt][0], acc[nt][1])); mx1 = fmaxf(mx1, fmaxf(acc[nt][2], acc[nt][3])); }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t p[A4_NTC / 2][4];
#pragma unroll
      for (int nt = 0; nt < A4_NTC; ++nt) {
        const float e0 = ex2f(acc[nt][0] - mx0), e1 = ex2f(acc[nt][1] - mx0);
        const float e2 = ex2f(acc[nt][2] - mx1), e3 = ex2f(acc[nt][3] - mx1);
        sum0 += e0 + e1; sum1 += e2 + e3;
        p[nt >> 1][(nt & 1) * 2] = pack_bf16(e0, e1);
        p[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(e2, e3);
      }
      sum0 = quad_sum(sum0); sum1 = quad_sum(sum1);
      if (lse && tq4 == 0) {
        if (s_src[r0] >= 0) lse[(int64_t)s_src[r0] * g.heads + h] = mx0 + log2f(sum0);
        if (s_src[r1] >= 0) lse[(int64_t)s_src[r1] * g.heads + h] = mx1 + log2f(sum1);
      }
      float o[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
      a4_strip_pv<0, A4_NTC / 2>(o, p, tv_s, lane);
      am_store_strip(o, 1.0f / sum0, 1.0f / sum1, stage_s, stage, lane, mt, g.N, s_src, out, g.C, h * 32, nullptr, nullptr);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------------
// One slab of query columns [NT0 * 8, (NT0 + NTN) * 8) for the key strip of this warp: S^T, P^T, dV += P^T dO, dP^T,
// dS^T (summed into Gm, parked in shared memory), dK += dS^T Q.
template <int NT0, int NTN>
__device__ __forceinline__ void a4_bwd_slab(const AmGeom& g, const AmWin& win, bool masked, int jt, int lane, uint32_t tq_s, uint32_t tk_s,
                                            uint32_t tv_s, uint32_t tdo_s, const float* s_tbl, const int* s_lin, const int* s_rg,
                                            const float* s_lse, const float* s_delta, float* Gm, uint32_t sds_s, float (&dv)[4][4],
                                            float (&dk)[4][4]) {
  const int gq = lane >> 2, tq4 = lane & 3;
  const int r0 = jt * 16 + gq, r1 = r0 + 8;                 // key rows of this thread
  const float MASK2 = -100.0f * AM_LOG2E;
  const bool row0_ok = r0 < g.N, row1_ok = r1 < g.N;
  const int lr0 = s_lin[r0], lr1 = s_lin[r1];
  const int rr0 = masked ? am_region(win, s_rg[r0]) : 0, rr1 = masked ? am_region(win, s_rg[r1]) : 0;
  uint32_t a[2][4];
  am_load_a(tk_s, jt, lane, a);
  float acc[NTN][4];
  a4_strip_nt<NT0, NTN>(acc, a, tq_s, lane);                // S^T[j][i] = k_j . q_i
  uint32_t pt[NTN / 2][4];
#pragma unroll
  for (int nt = 0; nt < NTN; ++nt) {
    const int c0 = (NT0 + nt) * 8 + 2 * tq4, c1 = c0 + 1;  // query columns
    const bool c0_ok = c0 < g.N, c1_ok = c1 < g.N;
    const int lc0 = s_lin[c0], lc1 = s_lin[c1];
    const float ls0 = s_lse[c0], ls1 = s_lse[c1];
    float s00 = fmaf(acc[nt][0], g.scale2, s_tbl[lc0 - lr0 + g.lin_off]);
    float s01 = fmaf(acc[nt][1], g.scale2, s_tbl[lc1 - lr0 + g.lin_off]);
    float s10 = fmaf(acc[nt][2], g.scale2, s_tbl[lc0 - lr1 + g.lin_off]);
    float s11 = fmaf(acc[nt][3], g.scale2, s_tbl[lc1 - lr1 + g.lin_off]);
    if (masked) {
      const int cg0 = am_region(win, s_rg[c0]), cg1 = am_region(win, s_rg[c1]);
      if (cg0 != rr0) s00 += MASK2;
      if (cg1 != rr0) s01 += MASK2;
      if (cg0 != rr1) s10 += MASK2;
      if (cg1 != rr1) s11 += MASK2;
    }
    acc[nt][0] = (row0_ok && c0_ok) ? ex2f(s00 - ls0) : 0.f;   // lse = 1e30 for absent queries -> 0
    acc[nt][1] = (row0_ok && c1_ok) ? ex2f(s01 - ls1) : 0.f;
    acc[nt][2] = (row1_ok && c0_ok) ? ex2f(s10 - ls0) : 0.f;
    acc[nt][3] = (row1_ok && c1_ok) ? ex2f(s11 - ls1) : 0.f;
    pt[nt >> 1][(nt & 1) * 2] = pack_bf16(acc[nt][0], acc[nt][1]);
    pt[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(acc[nt][2], acc[nt][3]);
  }
  a4_strip_pv<NT0 * 8, NTN / 2>(dv, pt, tdo_s, lane);        // dV[j][d] += sum_i P[i][j] dO[i][d]
  uint32_t av[2][4];
  am_load_a(tv_s, jt, lane, av);
  float dp[NTN][4];
  a4_strip_nt<NT0, NTN>(dp, av, tdo_s, lane);               // dP^T[j][i] = v_j . dO_i
  uint32_t dst[NTN / 2][4];
#pragma unroll
  for (int nt = 0; nt < NTN; ++nt) {
    const int c0 = (NT0 + nt) * 8 + 2 * tq4;
    const float d0 = s_delta[c0], d1 = s_delta[c0 + 1];
    const float v0 = acc[nt][0] * (dp[nt][0] - d0), v1 = acc[nt][1] * (dp[nt][1] - d1);
    const float v2 = acc[nt][2] * (dp[nt][2] - d0), v3 = acc[nt][3] * (dp[nt][3] - d1);
    float2* g0 = reinterpret_cast<float2*>(Gm + r0 * A4_G_LD + c0);
    float2* g1 = reinterpret_cast<float2*>(Gm + r1 * A4_G_LD + c0);
    float2 t0 = *g0, t1 = *g1;
    t0.x += v0; t0.y += v1; t1.x += v2; t1.y += v3;
    *g0 = t0; *g1 = t1;
    dst[nt >> 1][(nt & 1) * 2] = pack_bf16(v0, v1);
    dst[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(v2, v3);
  }
  a4_strip_pv<NT0 * 8, NTN / 2>(dk, dst, tq_s, lane);        // dK[j][d] += sum_i dS[i][j] q[i][d]
  // dS^T slab -> shared [key][query] (16-byte chunks XOR-swizzled by key row)
#pragma unroll
  for (int ks = 0; ks < NTN / 2; ++ks) {
    const int c0 = NT0 + ks * 2, c1 = c0 + 1;               // global 16-byte chunk index = query column / 8
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + r0 * A4_DS_PITCH + ((c0 ^ (r0 & 7)) << 4) + tq4 * 4), "r"(dst[ks][0]) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + r1 * A4_DS_PITCH + ((c0 ^ (r1 & 7)) << 4) + tq4 * 4), "r"(dst[ks][1]) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + r0 * A4_DS_PITCH + ((c1 ^ (r0 & 7)) << 4) + tq4 * 4), "r"(dst[ks][2]) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sds_s + r1 * A4_DS_PITCH + ((c1 ^ (r1 & 7)) << 4) + tq4 * 4), "r"(dst[ks][3]) : "memory");
  }
}

__global__ void __launch_bounds__(A4_THREADS, 1) window_attn_mma144_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ qkv,
                                                                                const bf16* __restrict__ outp, const float* __restrict__ lse,
                                                                                const float* __restrict__ table, const float* __restrict__ qkv_bias,
                                                                                bf16* __restrict__ dqkv, float* __restrict__ dtable,
                                                                                float* __restrict__ dqkv_bias, float* __restrict__ dqkv_colsum,
                                                                                AmGeom g) {
  extern __shared__ __align__(128) uint8_t a4_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % g.heads, grp = blockIdx.x / g.heads;
  uint8_t* tiles = a4_smem;                                          // q, k, v, dO, O
  uint8_t* stage = tiles + 5 * A4_TILE + warp * 1024;
  uint8_t* sds = tiles + 5 * A4_TILE + A4_NW * 1024;                 // dS^T [144 keys][A4_DS_PITCH bytes]
  float* Gm = reinterpret_cast<float*>(sds + A4_TOK * A4_DS_PITCH);  // [144][A4_G_LD] sums of dS^T over this CTA's windows
  float* s_tbl = Gm + A4_TOK * A4_G_LD;
  const int tab_bytes = (g.ntab * 4 + 15) & ~15;
  int* s_pos = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s_tbl) + tab_bytes);
  int* s_lin = s_pos + A4_TOK; int* s_rg = s_lin + A4_TOK; int* s_src = s_rg + A4_TOK;
  float* s_lse = reinterpret_cast<float*>(s_src + A4_TOK);
  float* s_delta = s_lse + A4_TOK;
  float* s_cs = s_delta + A4_TOK;                                    // [9 warps][3][32]
  const uint32_t tq_s = smem_u32_generic(tiles), tk_s = tq_s + A4_TILE, tv_s = tk_s + A4_TILE, tdo_s = tv_s + A4_TILE;
  const uint32_t stage_s = smem_u32_generic(stage), sds_s = smem_u32_generic(sds);
  for (int t = tid; t < g.ntab; t += A4_THREADS) s_tbl[t] = __ldg(table + t * g.heads + h) * AM_LOG2E;
  for (int t = tid; t < A4_TOK * A4_G_LD; t += A4_THREADS) Gm[t] = 0.f;
  a4_init_tables(g, tid, s_pos, s_lin, s_rg);
  pdl_trigger();
  __syncthreads();
  pdl_wait();
  const int gq = lane >> 2;
  const int n_mt = (g.N + 15) >> 4;
  const int jt = warp;
  float cs_k[8], cs_v[8], cs_q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) cs_k[k] = cs_v[k] = cs_q[k] = 0.f;

  for (int w = grp; w < g.windows; w += g.groups) {
    const AmWin win = am_window(g, w);
    const bool masked = win.last_row || win.last_col;
    const bool has_pad = am_has_pad(g, win);
    __syncthreads();                                    // previous window fully consumed
    a4_issue_window<true>(g, win, h, tq_s, qkv, dout, outp, lse, s_pos, s_src, s_lse, tid);
    cp_async_wait_all();
    __syncthreads();
    if (has_pad) {
      a4_fill_pad_rows(tiles, qkv_bias, h * 32, s_src, g.N, tid);
      a4_fill_pad_rows(tiles + A4_TILE, qkv_bias, g.C + h * 32, s_src, g.N, tid);
      a4_fill_pad_rows(tiles + 2 * A4_TILE, qkv_bias, 2 * g.C + h * 32, s_src, g.N, tid);
    }
    // delta_i = dO_i . O_i
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int tok = k * 72 + (tid >> 2), ch = tid & 3;
      const uint4 a = *reinterpret_cast<const uint4*>(tiles + 3 * A4_TILE + am_off(tok, ch));
      const uint4 b = *reinterpret_cast<const uint4*>(tiles + 4 * A4_TILE + am_off(tok, ch));
      const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 fa = __bfloat1622float2(ha[i]), fb = __bfloat1622float2(hb[i]); part = fmaf(fa.x, fb.x, part); part = fmaf(fa.y, fb.y, part); }
      part = quad_sum(part);
      if (ch == 0) s_delta[tok] = part;
    }
    __syncthreads();

    if (jt < n_mt) {
      float dv[4][4], dk[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f; dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f; }
      a4_bwd_slab<0, 10>(g, win, masked, jt, lane, tq_s, tk_s, tv_s, tdo_s, s_tbl, s_lin, s_rg, s_lse, s_delta, Gm, sds_s, dv, dk);
      a4_bwd_slab<10, 8>(g, win, masked, jt, lane, tq_s, tk_s, tv_s, tdo_s, s_tbl, s_lin, s_rg, s_lse, s_delta, Gm, sds_s, dv, dk);
      am_store_strip(dv, 1.f, 1.f, stage_s, stage, lane, jt, g.N, s_src, dqkv, 3 * g.C, 2 * g.C + h * 32, has_pad ? dqkv_bias : nullptr, dqkv_colsum ? cs_v : nullptr);
      am_store_strip(dk, g.scale, g.scale, stage_s, stage, lane, jt, g.N, s_src, dqkv, 3 * g.C, g.C + h * 32, has_pad ? dqkv_bias : nullptr, dqkv_colsum ? cs_k : nullptr);
    }
    __syncthreads();                                    // dS^T complete
    if (jt < n_mt) {
      // dQ strip jt = scale * dS[16 queries][keys] * K
      float dq[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
      const int mi = lane >> 3;
#pragma unroll
      for (int ks = 0; ks < A4_NW; ++ks) {
        if (ks < n_mt) {
          uint32_t a0, a1, a2, a3;
          const int key = ks * 16 + (mi >> 1) * 8 + (lane & 7), chunk = jt * 2 + (mi & 1);
          ldsm_x4_t(sds_s + key * A4_DS_PITCH + ((chunk ^ (key & 7)) << 4), a0, a1, a2, a3);
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
  // ---- flush (once per CTA): qkv-bias gradient and bias-table gradient ----
  const int tq4 = lane & 3;
  if (dqkv_colsum) {
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
  for (int t = tid; t < g.ntab; t += A4_THREADS) {
    const int dy = t / g.lin_stride - (g.wh - 1), dx = t % g.lin_stride - (g.ww - 1);
    float v = 0.f;
    for (int yj = max(0, -dy); yj < min(g.wh, g.wh - dy); ++yj)
      for (int xj = max(0, -dx); xj < min(g.ww, g.ww - dx); ++xj)
        v += Gm[(yj * g.ww + xj) * A4_G_LD + (yj + dy) * g.ww + xj + dx];
    if (v != 0.f) atomicAdd(dtable + t * g.heads + h, v);
  }
  if (dqkv_colsum && tid < 96) {
    float v = 0.f;
#pragma unroll
    for (int ws = 0; ws < A4_NW; ++ws) v += s_cs[ws * 96 + tid];
    atomicAdd(dqkv_colsum + (tid >> 5) * g.C + h * 32 + (tid & 31), v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
static int a4_geom(AmGeom& g, int B, int H, int W, int C, int heads, int wh, int ww, int sh, int sw) {
  if (B < 0 || H <= 0 || W <= 0 || heads <= 0 || C != heads * 32) return MTUS_ERR_BAD_ARG;
  if (wh <= 0 || ww <= 0 || wh * ww > A4_TOK || sh < 0 || sw < 0 || sh >= wh || sw >= ww) return MTUS_ERR_UNSUPPORTED;
  g.B = B; g.H = H; g.W = W; g.C = C; g.heads = heads; g.wh = wh; g.ww = ww; g.sh = sh; g.sw = sw;
  g.nwy = (H + wh - 1) / wh; g.nwx = (W + ww - 1) / ww;
  g.Hp = g.nwy * wh; g.Wp = g.nwx * ww;
  g.N = wh * ww; g.ntab = (2 * wh - 1) * (2 * ww - 1);
  g.lin_stride = 2 * ww - 1; g.lin_off = (wh - 1) * (2 * ww - 1) + (ww - 1);
  g.scale = 1.0f / sqrtf(32.0f);
  g.scale2 = g.scale * AM_LOG2E;
  g.windows = B * g.nwy * g.nwx;
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int G = (sms > 0 ? sms : 148) / heads;                // one CTA per SM, one wave
  if (G < 1) G = 1;
  if (G > g.windows) G = g.windows;
  const int per = (g.windows + G - 1) / G;
  g.groups = (g.windows + per - 1) / per;
  return MTUS_OK;
}

bool mtus_window_attn_mma144_supported(int wh, int ww, int dtype) {
  return dtype == MTUS_BF16 && wh > 0 && ww > 0 && wh * ww > 64 && wh * ww <= A4_TOK;
}

int mtus_window_attn_mma144_fwd(const void* qkv, const float* rel_table, const float* qkv_bias, void* out, float* lse, int B, int H, int W,
                                int C, int heads, int win_h, int win_w, int shift_h, int shift_w, cudaStream_t st) {
  AmGeom g;
  int rc = a4_geom(g, B, H, W, C, heads, win_h, win_w, shift_h, shift_w);
  if (rc) return rc;
  if ((g.Hp != H || g.Wp != W) && !qkv_bias) return MTUS_ERR_BAD_ARG;
  if (B == 0) return MTUS_OK;
  const size_t smem = a4_fwd_bytes();
  static mtus_per_device_flag configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_mma144_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured.set();
  }
  cudaError_t le = mtus_launch_pdl(window_attn_mma144_fwd_kernel, dim3(g.groups * heads), dim3(A4_THREADS), smem, st, (const bf16*)qkv, rel_table,
                                   qkv_bias, (bf16*)out, lse, g);
  if (le != cudaSuccess) return (int)le;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

int mtus_window_attn_mma144_bwd(const void* dout, const void* qkv, const void* out, const float* lse, const float* rel_table,
                                const float* qkv_bias, void* dqkv, float* drel_table, float* dqkv_bias, float* dqkv_colsum, int B, int H,
                                int W, int C, int heads, int win_h, int win_w, int shift_h, int shift_w, cudaStream_t st) {
  AmGeom g;
  int rc = a4_geom(g, B, H, W, C, heads, win_h, win_w, shift_h, shift_w);
  if (rc) return rc;
  if (!lse) return MTUS_ERR_BAD_ARG;
  if ((g.Hp != H || g.Wp != W) && !(qkv_bias && dqkv_bias)) return MTUS_ERR_BAD_ARG;
  if (B == 0) return MTUS_OK;
  const size_t smem = a4_bwd_bytes(g.ntab);
  static mtus_per_device_flag configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_mma144_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured.set();
  }
  cudaError_t le = mtus_launch_pdl(window_attn_mma144_bwd_kernel, dim3(g.groups * heads), dim3(A4_THREADS), smem, st, (const bf16*)dout, (const bf16*)qkv,
                                   (const bf16*)out, lse, rel_table, qkv_bias, (bf16*)dqkv, drel_table, dqkv_bias, dqkv_colsum, g);
  if (le != cudaSuccess) return (int)le;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}
