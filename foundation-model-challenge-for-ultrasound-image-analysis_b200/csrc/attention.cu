// Fused shifted-window attention, forward and backward (general engine: any window, fp32 math).
//
// Replaces timm SwinTransformerBlock._attn + window_partition/window_reverse + get_attn_mask +
// WindowAttention.forward minus the two Linear layers (reached from
// /root/reference/code/models/encoders.py:104).  One CTA owns one (image, window, head):
//   * the cyclic shift (roll -s), zero padding to a multiple of the window (timm order: roll THEN
//     pad), window partition, and their inverses are pure index math on the loads / stores --
//     nothing window-shaped is ever written to HBM;
//   * S = (q*scale) k^T + rel_pos_bias[idx(i,j)] + (-100 if region(i) != region(j)), softmax in
//     fp32, O = P v -- all inside shared memory;
//   * backward recomputes P, uses delta_i = dO_i . O_i, and reduces the relative-position-bias
//     gradient into (2w-1)^2 shared-memory bins before touching global atomics.
// Stand-alone this op is HBM-bound (24.5 FLOP/B at N=49): algorithmic bytes per token =
// 4*C*s forward (q,k,v in, o out), 8*C*s backward (q,k,v,o,do in; dq,dk,dv out).
#include "common.cuh"
#include "internal.h"
#include <stdlib.h>
#include <string.h>

#define ATT_D 32          // head_dim is 32 for every Swin variant
#define ATT_LD 33
#define ATT_THREADS 128

struct AttGeom {
  int B, H, W, C, heads, wh, ww, sh, sw, Hp, Wp, nwx, nwy;
  float scale;
};

// token t of window (wy,wx): source row in [B*H*W) or -1 for a padded token; region id for the shift mask
__device__ __forceinline__ void att_token(const AttGeom& g, int b, int wy, int wx, int t, int& src, int& reg) {
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

template <typename T>
__device__ __forceinline__ void att_load_qkv(const AttGeom& g, const T* __restrict__ qkv, const float* __restrict__ qkv_bias,
                                             const int* s_src, int N, int h, float* sq, float* sk, float* sv) {
  // items: (token, part in {q,k,v}, 8-wide chunk)
  for (int it = threadIdx.x; it < N * 12; it += ATT_THREADS) {
    const int t = it / 12, r = it - t * 12, part = r >> 2, c8 = (r & 3) * 8;
    const int col = part * g.C + h * ATT_D + c8;
    float v[8];
    if (s_src[t] >= 0) IO<T>::load8(qkv + (int64_t)s_src[t] * 3 * g.C + col, v);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = qkv_bias ? __ldg(qkv_bias + col + k) : 0.f;
    }
    float* dst = (part == 0 ? sq : (part == 1 ? sk : sv)) + t * ATT_LD + c8;
    const float f = (part == 0) ? g.scale : 1.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[k] = v[k] * f;
  }
}

// S[i][j] -> softmax probabilities, in place
__device__ __forceinline__ void att_scores_softmax(const AttGeom& g, int N, const float* sq, const float* sk,
                                                   const float* s_tbl, const int* s_reg, float* sS) {
  const int LDS = N + 1;
  const bool masked = (g.sh > 0) || (g.sw > 0);
  for (int idx = threadIdx.x; idx < N * N; idx += ATT_THREADS) {
    const int i = idx / N, j = idx - i * N;
    const float* qi = sq + i * ATT_LD; const float* kj = sk + j * ATT_LD;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) s = fmaf(qi[d], kj[d], s);
    const int yi = i / g.ww, xi = i - yi * g.ww, yj = j / g.ww, xj = j - yj * g.ww;
    s += s_tbl[(yi - yj + g.wh - 1) * (2 * g.ww - 1) + (xi - xj + g.ww - 1)];
    if (masked && s_reg[i] != s_reg[j]) s += -100.0f;
    sS[i * LDS + j] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < N; i += ATT_THREADS / 32) {
    float* row = sS + i * LDS;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < N; j += 32) row[j] *= inv;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) window_attn_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ table,
                                                                      const float* __restrict__ qkv_bias, T* __restrict__ out, AttGeom g) {
  extern __shared__ float sm[];
  const int N = g.wh * g.ww, ntab = (2 * g.wh - 1) * (2 * g.ww - 1);
  float* sq = sm; float* sk = sq + N * ATT_LD; float* sv = sk + N * ATT_LD;
  float* sS = sv + N * ATT_LD; float* s_tbl = sS + N * (N + 1);
  int* s_src = reinterpret_cast<int*>(s_tbl + ntab); int* s_reg = s_src + N;
  const int h = blockIdx.x % g.heads;
  int w = blockIdx.x / g.heads;
  const int wx = w % g.nwx; w /= g.nwx; const int wy = w % g.nwy; const int b = w / g.nwy;

  for (int t = threadIdx.x; t < N; t += ATT_THREADS) att_token(g, b, wy, wx, t, s_src[t], s_reg[t]);
  for (int t = threadIdx.x; t < ntab; t += ATT_THREADS) s_tbl[t] = __ldg(table + t * g.heads + h);
  __syncthreads();
  att_load_qkv<T>(g, qkv, qkv_bias, s_src, N, h, sq, sk, sv);
  __syncthreads();
  att_scores_softmax(g, N, sq, sk, s_tbl, s_reg, sS);
  // O = P V ; lanes run over d -> coalesced 32-element stores
  const int LDS = N + 1;
  for (int idx = threadIdx.x; idx < N * ATT_D; idx += ATT_THREADS) {
    const int i = idx >> 5, d = idx & 31;
    if (s_src[i] < 0) continue;
    const float* p = sS + i * LDS;
    float o = 0.f;
    for (int j = 0; j < N; ++j) o = fmaf(p[j], sv[j * ATT_LD + d], o);
    IO<T>::st(out + (int64_t)s_src[i] * g.C + h * ATT_D + d, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) window_attn_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ qkv,
                                                                      const T* __restrict__ outp, const float* __restrict__ table,
                                                                      const float* __restrict__ qkv_bias, T* __restrict__ dqkv,
                                                                      float* __restrict__ dtable, float* __restrict__ dqkv_bias, AttGeom g) {
  extern __shared__ float sm[];
  const int N = g.wh * g.ww, ntab = (2 * g.wh - 1) * (2 * g.ww - 1);
  float* sq = sm; float* sk = sq + N * ATT_LD; float* sv = sk + N * ATT_LD; float* sdo = sv + N * ATT_LD;
  float* sS = sdo + N * ATT_LD; float* s_tbl = sS + N * (N + 1); float* s_bins = s_tbl + ntab; float* s_delta = s_bins + ntab;
  int* s_src = reinterpret_cast<int*>(s_delta + N); int* s_reg = s_src + N;
  const int h = blockIdx.x % g.heads;
  int w = blockIdx.x / g.heads;
  const int wx = w % g.nwx; w /= g.nwx; const int wy = w % g.nwy; const int b = w / g.nwy;
  const int LDS = N + 1;

  for (int t = threadIdx.x; t < N; t += ATT_THREADS) att_token(g, b, wy, wx, t, s_src[t], s_reg[t]);
  for (int t = threadIdx.x; t < ntab; t += ATT_THREADS) { s_tbl[t] = __ldg(table + t * g.heads + h); s_bins[t] = 0.f; }
  __syncthreads();
  att_load_qkv<T>(g, qkv, qkv_bias, s_src, N, h, sq, sk, sv);
  // dO (zero for padded tokens: their outputs are cropped away) and delta_i = dO_i . O_i
  for (int it = threadIdx.x; it < ((N * 4 + 31) & ~31); it += ATT_THREADS) {   // warp-uniform trip count (shuffles below)
    const bool live = it < N * 4;
    const int t = live ? (it >> 2) : 0, c8 = (it & 3) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (live && s_src[t] >= 0) {
      IO<T>::load8(dout + (int64_t)s_src[t] * g.C + h * ATT_D + c8, v);
      IO<T>::load8(outp + (int64_t)s_src[t] * g.C + h * ATT_D + c8, o);
    }
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (live) sdo[t * ATT_LD + c8 + k] = v[k]; part += v[k] * o[k]; }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (live && (it & 3) == 0) s_delta[t] = part;
  }
  __syncthreads();
  att_scores_softmax(g, N, sq, sk, s_tbl, s_reg, sS);   // sS = P

  // dV[j][d] = sum_i P[i][j] dO[i][d]
  for (int idx = threadIdx.x; idx < N * ATT_D; idx += ATT_THREADS) {
    const int j = idx >> 5, d = idx & 31;
    float a = 0.f;
    for (int i = 0; i < N; ++i) a = fmaf(sS[i * LDS + j], sdo[i * ATT_LD + d], a);
    const int col = 2 * g.C + h * ATT_D + d;
    if (s_src[j] >= 0) IO<T>::st(dqkv + (int64_t)s_src[j] * 3 * g.C + col, a);
    else if (dqkv_bias) atomicAdd(dqkv_bias + col, a);
  }
  __syncthreads();
  // dS = P * (dO_i . v_j - delta_i), in place; relative-position-bias gradient into smem bins
  for (int idx = threadIdx.x; idx < N * N; idx += ATT_THREADS) {
    const int i = idx / N, j = idx - i * N;
    const float* doi = sdo + i * ATT_LD; const float* vj = sv + j * ATT_LD;
    float dp = 0.f;
#pragma unroll
    for (int d = 0; d < ATT_D; ++d) dp = fmaf(doi[d], vj[d], dp);
    const float ds = sS[i * LDS + j] * (dp - s_delta[i]);
    sS[i * LDS + j] = ds;
    const int yi = i / g.ww, xi = i - yi * g.ww, yj = j / g.ww, xj = j - yj * g.ww;
    atomicAdd(&s_bins[(yi - yj + g.wh - 1) * (2 * g.ww - 1) + (xi - xj + g.ww - 1)], ds);
  }
  __syncthreads();
  // dq[i][d] = scale * sum_j dS[i][j] k[j][d] ;  dk[j][d] = sum_i dS[i][j] (scale*q)[i][d]
  for (int idx = threadIdx.x; idx < 2 * N * ATT_D; idx += ATT_THREADS) {
    const int which = idx / (N * ATT_D);
    const int r = idx - which * N * ATT_D;
    const int t = r >> 5, d = r & 31;
    float a = 0.f;
    if (which == 0) {
      for (int j = 0; j < N; ++j) a = fmaf(sS[t * LDS + j], sk[j * ATT_LD + d], a);
      a *= g.scale;
    } else {
      for (int i = 0; i < N; ++i) a = fmaf(sS[i * LDS + t], sq[i * ATT_LD + d], a);
    }
    const int col = which * g.C + h * ATT_D + d;
    if (s_src[t] >= 0) IO<T>::st(dqkv + (int64_t)s_src[t] * 3 * g.C + col, a);
    else if (dqkv_bias && which == 1) atomicAdd(dqkv_bias + col, a);
  }
  for (int t = threadIdx.x; t < ntab; t += ATT_THREADS) atomicAdd(dtable + t * g.heads + h, s_bins[t]);
}

// tensor-core engine (attention_mma.cu): bf16, windows of up to 64 tokens.  MTUS_ATTN=simt forces the general engine.
bool mtus_window_attn_mma_supported(int wh, int ww, int dtype);
int mtus_window_attn_mma_fwd(const void* qkv, const float* rel_table, const float* qkv_bias, void* out, float* lse, int B, int H, int W,
                             int C, int heads, int win_h, int win_w, int shift_h, int shift_w, cudaStream_t st);
int mtus_window_attn_mma_bwd(const void* dout, const void* qkv, const void* out, const float* lse, const float* rel_table,
                             const float* qkv_bias, void* dqkv, float* drel_table, float* dqkv_bias, float* dqkv_colsum, int B, int H,
                             int W, int C, int heads, int win_h, int win_w, int shift_h, int shift_w, cudaStream_t st);
// 65..144-token windows (window 12): attention_mma144.cu
bool mtus_window_attn_mma144_supported(int wh, int ww, int dtype);
int mtus_window_attn_mma144_fwd(const void* qkv, const float* rel_table, const float* qkv_bias, void* out, float* lse, int B, int H, int W,
                                int C, int heads, int win_h, int win_w, int shift_h, int shift_w, cudaStream_t st);
int mtus_window_attn_mma144_bwd(const void* dout, const void* qkv, const void* out, const float* lse, const float* rel_table,
                                const float* qkv_bias, void* dqkv, float* drel_table, float* dqkv_bias, float* dqkv_colsum, int B, int H,
                                int W, int C, int heads, int win_h, int win_w, int shift_h, int shift_w, cudaStream_t st);
static bool att_forced_simt() {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("MTUS_ATTN"); forced = (e && !strcmp(e, "simt")) ? 1 : 0; }
  return forced == 1;
}
static bool att_use_mma144(int wh, int ww, int dtype) { return !att_forced_simt() && mtus_window_attn_mma144_supported(wh, ww, dtype); }

static bool att_use_mma(int wh, int ww, int dtype) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("MTUS_ATTN"); forced = (e && !strcmp(e, "simt")) ? 1 : 0; }
  return !forced && mtus_window_attn_mma_supported(wh, ww, dtype);
}

static int att_geom(AttGeom& g, int B, int H, int W, int C, int heads, int wh, int ww, int sh, int sw) {
  if (B < 0 || H <= 0 || W <= 0 || heads <= 0 || C != heads * ATT_D) return MTUS_ERR_BAD_ARG;
  if (wh <= 0 || ww <= 0 || wh > 16 || ww > 16 || sh < 0 || sw < 0 || sh >= wh || sw >= ww) return MTUS_ERR_BAD_ARG;
  g.B = B; g.H = H; g.W = W; g.C = C; g.heads = heads; g.wh = wh; g.ww = ww; g.sh = sh; g.sw = sw;
  g.nwy = (H + wh - 1) / wh; g.nwx = (W + ww - 1) / ww;
  g.Hp = g.nwy * wh; g.Wp = g.nwx * ww;
  g.scale = 1.0f / sqrtf((float)ATT_D);
  return MTUS_OK;
}

extern "C" int mtus_window_attn_fwd(const void* qkv, const float* rel_table, const float* qkv_bias, void* out, float* lse,
                                    int B, int H, int W, int C, int heads, int win_h, int win_w, int shift_h, int shift_w,
                                    int dtype, void* stream) {
  MTUS_CHECK_ARG(qkv && rel_table && out);
  if (mtus_window_attn_tc_eligible(B, H, W, C, heads, win_h, win_w, shift_h, shift_w, dtype))      // tcgen05 / TMEM / TMA engine
    return mtus_window_attn_tc_fwd(qkv, rel_table, out, lse, B, H, W, C, heads, win_h, win_w, shift_h, shift_w, (cudaStream_t)stream);
  if (att_use_mma(win_h, win_w, dtype))
    return mtus_window_attn_mma_fwd(qkv, rel_table, qkv_bias, out, lse, B, H, W, C, heads, win_h, win_w, shift_h, shift_w, (cudaStream_t)stream);
  if (att_use_mma144(win_h, win_w, dtype) && lse)
    return mtus_window_attn_mma144_fwd(qkv, rel_table, qkv_bias, out, lse, B, H, W, C, heads, win_h, win_w, shift_h, shift_w, (cudaStream_t)stream);
  AttGeom g;
  int rc = att_geom(g, B, H, W, C, heads, win_h, win_w, shift_h, shift_w);
  if (rc) return rc;
  if ((g.Hp != H || g.Wp != W) && !qkv_bias) return MTUS_ERR_BAD_ARG;  // padded tokens need the qkv bias
  if (B == 0) return MTUS_OK;
  const int N = win_h * win_w, ntab = (2 * win_h - 1) * (2 * win_w - 1);
  const size_t smem = sizeof(float) * (3 * N * ATT_LD + N * (N + 1) + ntab) + sizeof(int) * 2 * N;
  const int64_t blocks = (int64_t)B * g.nwy * g.nwx * heads;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (dtype == MTUS_F32) {
    e = cudaFuncSetAttribute(window_attn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    window_attn_fwd_kernel<float><<<(unsigned)blocks, ATT_THREADS, smem, st>>>((const float*)qkv, rel_table, qkv_bias, (float*)out, g);
  } else if (dtype == MTUS_BF16) {
    e = cudaFuncSetAttribute(window_attn_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    window_attn_fwd_kernel<bf16><<<(unsigned)blocks, ATT_THREADS, smem, st>>>((const bf16*)qkv, rel_table, qkv_bias, (bf16*)out, g);
  } else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_window_attn_bwd(const void* dout, const void* qkv, const void* out, const float* lse,
                                    const float* rel_table, const float* qkv_bias, void* dqkv, float* drel_table,
                                    float* dqkv_bias, float* dqkv_colsum, int B, int H, int W, int C, int heads, int win_h,
                                    int win_w, int shift_h, int shift_w, int dtype, void* stream) {
  MTUS_CHECK_ARG(dout && qkv && out && rel_table && dqkv && drel_table);
  if (att_use_mma(win_h, win_w, dtype))
    return mtus_window_attn_mma_bwd(dout, qkv, out, lse, rel_table, qkv_bias, dqkv, drel_table, dqkv_bias, dqkv_colsum, B, H, W, C, heads,
                                    win_h, win_w, shift_h, shift_w, (cudaStream_t)stream);
  if (att_use_mma144(win_h, win_w, dtype) && lse)
    return mtus_window_attn_mma144_bwd(dout, qkv, out, lse, rel_table, qkv_bias, dqkv, drel_table, dqkv_bias, dqkv_colsum, B, H, W, C, heads,
                                       win_h, win_w, shift_h, shift_w, (cudaStream_t)stream);
  AttGeom g;
  int rc = att_geom(g, B, H, W, C, heads, win_h, win_w, shift_h, shift_w);
  if (rc) return rc;
  if ((g.Hp != H || g.Wp != W) && !(qkv_bias && dqkv_bias)) return MTUS_ERR_BAD_ARG;
  if (B == 0) return MTUS_OK;
  const int N = win_h * win_w, ntab = (2 * win_h - 1) * (2 * win_w - 1);
  const size_t smem = sizeof(float) * (4 * N * ATT_LD + N * (N + 1) + 2 * ntab + N) + sizeof(int) * 2 * N;
  const int64_t blocks = (int64_t)B * g.nwy * g.nwx * heads;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (dtype == MTUS_F32) {
    e = cudaFuncSetAttribute(window_attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    window_attn_bwd_kernel<float><<<(unsigned)blocks, ATT_THREADS, smem, st>>>((const float*)dout, (const float*)qkv, (const float*)out, rel_table, qkv_bias, (float*)dqkv, drel_table, dqkv_bias, g);
  } else if (dtype == MTUS_BF16) {
    e = cudaFuncSetAttribute(window_attn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    window_attn_bwd_kernel<bf16><<<(unsigned)blocks, ATT_THREADS, smem, st>>>((const bf16*)dout, (const bf16*)qkv, (const bf16*)out, rel_table, qkv_bias, (bf16*)dqkv, drel_table, dqkv_bias, g);
  } else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  if (dqkv_colsum) return mtus_colsum(dqkv, dqkv_colsum, (int64_t)B * H * W, 3 * C, dtype, stream);   // general engine: separate pass
  return MTUS_OK;
}
