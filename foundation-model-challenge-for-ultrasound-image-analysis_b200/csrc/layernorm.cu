// LayerNorm forward / backward (bandwidth-bound; HBM roofline), with an optional fused
// PatchMerging 2x2 gather on the input side.
//
// Replaces: timm LayerNorm call sites in SwinTransformerBlock.norm1/norm2, PatchEmbed.norm and
// PatchMerging.norm (reached from /root/reference/code/models/encoders.py:104); the gather mode
// replaces PatchMerging's reshape/permute/flatten copy.
//
// Layout: rows of C contiguous elements (NHWC tokens).  One warp (WPR=1) or four warps (WPR=4)
// per row; each lane owns NV 8-element vectors (16 B for bf16, 32 B for fp32), statistics in
// fp32 via a two-pass (mean, then centred variance) over registers, warp-shuffle reductions.
// Algorithmic bytes: fwd 2*rows*C*s (+8 B/row stats), bwd 3*rows*C*s (+ dres: 4*rows*C*s).
#include "common.cuh"
#include <stdlib.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

struct MergeGeom { int B, H, W, C, Ho, Wo; };  // input [B,H,W,C] -> rows (b,i,j) of 4C columns

template <typename T, int MODE>
__device__ __forceinline__ const T* ln_src(const T* x, int64_t row, int col, int C, const MergeGeom& g, bool& valid) {
  if (MODE == 0) { valid = true; return x + row * C + col; }
  const int q = col / g.C, cc = col - q * g.C;          // q: 0 h0w0 | 1 h1w0 | 2 h0w1 | 3 h1w1
  const int j = (int)(row % g.Wo); const int64_t t = row / g.Wo;
  const int i = (int)(t % g.Ho); const int64_t b = t / g.Ho;
  const int y = 2 * i + (q & 1), xx = 2 * j + (q >> 1);
  valid = (y < g.H) && (xx < g.W);
  return x + ((b * g.H + y) * (int64_t)g.W + xx) * g.C + cc;
}

template <typename T, int NV, int WPR, int MODE>
__global__ void __launch_bounds__(128) ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, T* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     int64_t rows, int C, float eps, MergeGeom g) {
  __shared__ float red[2][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rpb = 4 / WPR;                                    // rows per block-iteration
  const int sub = (WPR == 1) ? 0 : warp;                      // which quarter of the row this warp owns
  const float invC = 1.0f / (float)C;
  for (int64_t row0 = (int64_t)blockIdx.x * rpb; row0 < rows; row0 += (int64_t)gridDim.x * rpb) {
    const int64_t row = row0 + ((WPR == 1) ? warp : 0);
    const bool active = row < rows;
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = ((sub * NV + i) * 32 + lane) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
      if (active && col < C) {
        bool valid; const T* p = ln_src<T, MODE>(x, row, col, C, g, valid);
        if (valid) IO<T>::load8(p, v[i]);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[i][k];
      }
    }
    s = warp_sum(s);
    if (WPR > 1) {
      if (lane == 0) red[0][warp] = s;
      __syncthreads();
      s = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    }
    const float mean = s * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = ((sub * NV + i) * 32 + lane) * 8;
      if (active && col < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; q += d * d; }
      }
    }
    q = warp_sum(q);
    if (WPR > 1) {
      if (lane == 0) red[1][warp] = q;
      __syncthreads();
      q = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    }
    const float rstd = rsqrtf(q * invC + eps);
    if (active) {
      if (lane == 0 && sub == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = ((sub * NV + i) * 32 + lane) * 8;
        if (col < C) {
          float gm[8], bt[8], o[8];
          IO<float>::load8(gamma + col, gm);
          IO<float>::load8(beta + col, bt);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = (v[i][k] - mean) * rstd * gm[k] + bt[k];
          IO<T>::store8(y + row * C + col, o);
        }
      }
    }
    if (WPR > 1) __syncthreads();
  }
}

template <typename T, int NV, int WPR, int MODE>
__global__ void __launch_bounds__(128) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                     const float* __restrict__ rstd_in, const T* dres,
                                                     T* dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, int64_t rows, int C, MergeGeom g) {
  __shared__ float red[2][4];
  extern __shared__ float acc_smem[];                         // WPR==1: [2][4 warps][C] partial dgamma/dbeta
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rpb = 4 / WPR;
  const int sub = (WPR == 1) ? 0 : warp;
  const float invC = 1.0f / (float)C;
  float adg[NV][8], adb[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) { adg[i][k] = 0.f; adb[i][k] = 0.f; }

  for (int64_t row0 = (int64_t)blockIdx.x * rpb; row0 < rows; row0 += (int64_t)gridDim.x * rpb) {
    const int64_t row = row0 + ((WPR == 1) ? warp : 0);
    const bool active = row < rows;
    const float mean = active ? mean_in[row] : 0.f, rstd = active ? rstd_in[row] : 0.f;
    float xh[NV][8], gy[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = ((sub * NV + i) * 32 + lane) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) { xh[i][k] = 0.f; gy[i][k] = 0.f; }
      if (active && col < C) {
        float xv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dv[8], gm[8];
        bool valid; const T* p = ln_src<T, MODE>(x, row, col, C, g, valid);
        if (valid) IO<T>::load8(p, xv);
        IO<T>::load8(dy + row * C + col, dv);
        IO<float>::load8(gamma + col, gm);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          xh[i][k] = (xv[k] - mean) * rstd;
          gy[i][k] = dv[k] * gm[k];
          s1 += gy[i][k];
          s2 += gy[i][k] * xh[i][k];
          adg[i][k] += dv[k] * xh[i][k];
          adb[i][k] += dv[k];
        }
      }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (WPR > 1) {
      if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
      __syncthreads();
      s1 = red[0][0] + red[0][1] + red[0][2] + red[0][3];
      s2 = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    }
    s1 *= invC; s2 *= invC;
    if (active) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = ((sub * NV + i) * 32 + lane) * 8;
        if (col < C) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = rstd * (gy[i][k] - s1 - xh[i][k] * s2);
          bool valid; const T* p = ln_src<T, MODE>(x, row, col, C, g, valid);
          if (valid) {
            const int64_t off = p - x;
            if (dres) {
              float r[8]; IO<T>::load8(dres + off, r);
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] += r[k];
            }
            IO<T>::store8(dx + off, o);
          }
        }
      }
    }
    if (WPR > 1) __syncthreads();
  }
  // parameter gradients: reduce over the block, then one atomicAdd per column per block
  if (WPR == 1) {
    float* sg = acc_smem; float* sb = acc_smem + 4 * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 8;
      if (col < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { sg[warp * C + col + k] = adg[i][k]; sb[warp * C + col + k] = adb[i][k]; }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      atomicAdd(dgamma + c, sg[c] + sg[C + c] + sg[2 * C + c] + sg[3 * C + c]);
      atomicAdd(dbeta + c, sb[c] + sb[C + c] + sb[2 * C + c] + sb[3 * C + c]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = ((sub * NV + i) * 32 + lane) * 8;
      if (col < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { atomicAdd(dgamma + col + k, adg[i][k]); atomicAdd(dbeta + col + k, adb[i][k]); }
      }
    }
  }
}

// ---- dispatch -------------------------------------------------------------------------------
template <typename T, int MODE>
static int ln_launch(bool fwd, const void* a0, const void* a1, const float* gamma, const float* beta_or_mean,
                     const float* rstd_in, const void* dres, void* out, float* o1, float* o2, int64_t rows, int C,
                     float eps, MergeGeom g, cudaStream_t st) {
  if (rows == 0) return MTUS_OK;
  MTUS_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 4096);
  const int nvec = C / 8;                                      // 8-wide vectors per row
#define LN_CASE(NV_, WPR_)                                                                                   \
  {                                                                                                          \
    const int rpb = 4 / WPR_;                                                                                \
    int64_t blocks = (rows + rpb - 1) / rpb;                                                                 \
    if (fwd) {                                                                                               \
      if (blocks > 148 * 16) blocks = 148 * 16;                                                              \
      ln_fwd_kernel<T, NV_, WPR_, MODE><<<(int)blocks, 128, 0, st>>>((const T*)a0, gamma, beta_or_mean,      \
                                                                    (T*)out, o1, o2, rows, C, eps, g);       \
    } else {                                                                                                 \
      if (blocks > 148 * 4) blocks = 148 * 4;                                                                \
      const size_t sm = (WPR_ == 1) ? (size_t)8 * C * sizeof(float) : 0;                                     \
      ln_bwd_kernel<T, NV_, WPR_, MODE><<<(int)blocks, 128, sm, st>>>((const T*)a0, (const T*)a1, gamma,     \
                                                                     beta_or_mean, rstd_in, (const T*)dres, \
                                                                     (T*)out, o1, o2, rows, C, g);           \
    }                                                                                                        \
  }
  if (nvec <= 32) LN_CASE(1, 1)
  else if (nvec <= 64) LN_CASE(2, 1)
  else if (nvec <= 96) LN_CASE(3, 1)
  else if (nvec <= 128) LN_CASE(1, 4)
  else if (nvec <= 256) LN_CASE(2, 4)
  else if (nvec <= 384) LN_CASE(3, 4)
  else LN_CASE(4, 4)
#undef LN_CASE
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                  float* rstd, int64_t rows, int C, float eps, int dtype, void* stream) {
  MTUS_CHECK_ARG(x && gamma && beta && y && mean && rstd && rows >= 0);
  MergeGeom g{};
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) return ln_launch<float, 0>(true, x, nullptr, gamma, beta, nullptr, nullptr, y, mean, rstd, rows, C, eps, g, st);
  if (dtype == MTUS_BF16) return ln_launch<bf16, 0>(true, x, nullptr, gamma, beta, nullptr, nullptr, y, mean, rstd, rows, C, eps, g, st);
  return MTUS_ERR_UNSUPPORTED;
}

extern "C" int mtus_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                  const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                                  int64_t rows, int C, int dtype, void* stream) {
  MTUS_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && rows >= 0);
  MergeGeom g{};
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) return ln_launch<float, 0>(false, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, C, 0.f, g, st);
  if (dtype == MTUS_BF16) return ln_launch<bf16, 0>(false, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, C, 0.f, g, st);
  return MTUS_ERR_UNSUPPORTED;
}

extern "C" int mtus_patch_merge_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                       float* rstd, int B, int H, int W, int C, float eps, int dtype, void* stream) {
  MTUS_CHECK_ARG(x && gamma && beta && y && mean && rstd && B >= 0 && H > 0 && W > 0 && C % 8 == 0);
  MergeGeom g{B, H, W, C, (H + 1) / 2, (W + 1) / 2};
  const int64_t rows = (int64_t)B * g.Ho * g.Wo;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) return ln_launch<float, 1>(true, x, nullptr, gamma, beta, nullptr, nullptr, y, mean, rstd, rows, 4 * C, eps, g, st);
  if (dtype == MTUS_BF16) return ln_launch<bf16, 1>(true, x, nullptr, gamma, beta, nullptr, nullptr, y, mean, rstd, rows, 4 * C, eps, g, st);
  return MTUS_ERR_UNSUPPORTED;
}

// dx is written at the (un-gathered) [B,H,W,C] positions; dres (same layout as x) is added if given.
extern "C" int mtus_patch_merge_ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                       const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                                       int B, int H, int W, int C, int dtype, void* stream) {
  MTUS_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && B >= 0 && H > 0 && W > 0 && C % 8 == 0);
  MergeGeom g{B, H, W, C, (H + 1) / 2, (W + 1) / 2};
  const int64_t rows = (int64_t)B * g.Ho * g.Wo;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) return ln_launch<float, 1>(false, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, 4 * C, 0.f, g, st);
  if (dtype == MTUS_BF16) return ln_launch<bf16, 1>(false, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, 4 * C, 0.f, g, st);
  return MTUS_ERR_UNSUPPORTED;
}

// =================================================================================================
// Mixed-precision variants for the fp32 residual stream (bf16 mode) -- also used, with every type float, by
// the fp32 mode so the executors have a single schedule.
//   forward : x (TX) -> y (TY); statistics in fp32.
//   backward: dx (fp32, optional) = dres (fp32, optional) + LN'(dy);  dx_lp (T, optional) = rowscale[sample] * dx
//             rounded to the GEMM operand type, together with its column sums (lp_colsum += sum_rows dx_lp): the
//             operand of the next weight-gradient GEMM and the bias gradient of that Linear come out of this pass,
//             so no separate cast / drop-path scale / column-sum kernels run over the gradient stream.
// =================================================================================================
template <typename TX, typename TY, int NV, int WPR, int MODE>
__global__ void __launch_bounds__(128) lnx_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, TY* __restrict__ y,
                                                      float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                      int64_t rows, int C, float eps, MergeGeom g) {
  __shared__ float red[2][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rpb = 4 / WPR;
  const int sub = (WPR == 1) ? 0 : warp;
  const float invC = 1.0f / (float)C;
  for (int64_t row0 = (int64_t)blockIdx.x * rpb; row0 < rows; row0 += (int64_t)gridDim.x * rpb) {
    const int64_t row = row0 + ((WPR == 1) ? warp : 0);
    const bool active = row < rows;
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = ((sub * NV + i) * 32 + lane) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) v[i][k] = 0.f;
      if (active && col < C) {
        bool valid; const TX* p = ln_src<TX, MODE>(x, row, col, C, g, valid);
        if (valid) IO<TX>::load8(p, v[i]);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[i][k];
      }
    }
    s = warp_sum(s);
    if (WPR > 1) {
      if (lane == 0) red[0][warp] = s;
      __syncthreads();
      s = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    }
    const float mean = s * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = ((sub * NV + i) * 32 + lane) * 8;
      if (active && col < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; q += d * d; }
      }
    }
    q = warp_sum(q);
    if (WPR > 1) {
      if (lane == 0) red[1][warp] = q;
      __syncthreads();
      q = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    }
    const float rstd = rsqrtf(q * invC + eps);
    if (active) {
      if (lane == 0 && sub == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = ((sub * NV + i) * 32 + lane) * 8;
        if (col < C) {
          float gm[8], bt[8], o[8];
          IO<float>::load8(gamma + col, gm);
          IO<float>::load8(beta + col, bt);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = (v[i][k] - mean) * rstd * gm[k] + bt[k];
          IO<TY>::store8(y + row * C + col, o);
        }
      }
    }
    if (WPR > 1) __syncthreads();
  }
}

struct LnxBwdArgs {
  const void* dy; const void* x; const float* gamma; const float* mean; const float* rstd;
  const float* dres; float* dx; void* dx_lp; const float* lp_rowscale; int rows_per_sample; float* lp_colsum;
  float* dgamma; float* dbeta; int64_t rows; int C;
};

template <typename T, typename TDY, typename TX, int NV, int WPR, int MODE>
__global__ void __launch_bounds__(128) lnx_bwd_kernel(LnxBwdArgs a, MergeGeom g) {
  __shared__ float red[2][4];
  extern __shared__ float acc_smem[];                         // WPR==1: [3][4 warps][Cx] partial dgamma / dbeta / colsum
  const TDY* __restrict__ dy = reinterpret_cast<const TDY*>(a.dy);
  const TX* __restrict__ x = reinterpret_cast<const TX*>(a.x);
  T* __restrict__ dx_lp = reinterpret_cast<T*>(a.dx_lp);
  const int C = a.C;
  const int64_t rows = a.rows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rpb = 4 / WPR;
  const int sub = (WPR == 1) ? 0 : warp;
  const float invC = 1.0f / (float)C;
  const bool want_cs = a.lp_colsum != nullptr;
  float adg[NV][8], adb[NV][8], acs[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) { adg[i][k] = 0.f; adb[i][k] = 0.f; acs[i][k] = 0.f; }

  for (int64_t row0 = (int64_t)blockIdx.x * rpb; row0 < rows; row0 += (int64_t)gridDim.x * rpb) {
    const int64_t row = row0 + ((WPR == 1) ? warp : 0);
    const bool active = row < rows;
    const float mean = active ? a.mean[row] : 0.f, rstd = active ? a.rstd[row] : 0.f;
    float xh[NV][8], gy[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = ((sub * NV + i) * 32 + lane) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) { xh[i][k] = 0.f; gy[i][k] = 0.f; }
      if (active && col < C) {
        float xv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dv[8], gm[8];
        bool valid; const TX* p = ln_src<TX, MODE>(x, row, col, C, g, valid);
        if (valid) IO<TX>::load8(p, xv);
        IO<TDY>::load8(dy + row * C + col, dv);
        IO<float>::load8(a.gamma + col, gm);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          xh[i][k] = (xv[k] - mean) * rstd;
          gy[i][k] = dv[k] * gm[k];
          s1 += gy[i][k];
          s2 += gy[i][k] * xh[i][k];
          adg[i][k] += dv[k] * xh[i][k];
          adb[i][k] += dv[k];
        }
      }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (WPR > 1) {
      if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
      __syncthreads();
      s1 = red[0][0] + red[0][1] + red[0][2] + red[0][3];
      s2 = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    }
    s1 *= invC; s2 *= invC;
    if (active) {
      float rsc = 1.0f;
      if (a.lp_rowscale) {
        rsc = __ldg(a.lp_rowscale + row / a.rows_per_sample);   // rows_per_sample in units of this kernel's rows
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = ((sub * NV + i) * 32 + lane) * 8;
        if (col < C) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = rstd * (gy[i][k] - s1 - xh[i][k] * s2);
          bool valid; const TX* p = ln_src<TX, MODE>(x, row, col, C, g, valid);
          if (valid) {
            const int64_t off = p - x;
            if (a.dres) {
              float r[8]; IO<float>::load8(a.dres + off, r);
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] += r[k];
            }
            if (a.dx) IO<float>::store8(a.dx + off, o);
            if (dx_lp) {
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] *= rsc;
              IO<T>::store8(dx_lp + off, o);
              if (want_cs) {
                // sum what the GEMM will read: the rounded values
                float rr[8]; IO<T>::load8_reg(o, rr);
#pragma unroll
                for (int k = 0; k < 8; ++k) acs[i][k] += rr[k];
              }
            }
          }
        }
      }
    }
    if (WPR > 1) __syncthreads();
  }
  // parameter gradients / column sums: reduce over the block, then one atomicAdd per column per block
  const int Cc = (MODE == 0) ? C : g.C;                        // columns of the un-gathered tensor (lp_colsum)
  if (WPR == 1) {
    float* sg = acc_smem; float* sb = acc_smem + 4 * C; float* sc = acc_smem + 8 * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 8;
      if (col < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { sg[warp * C + col + k] = adg[i][k]; sb[warp * C + col + k] = adb[i][k]; sc[warp * C + col + k] = acs[i][k]; }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      MTUS_ATOMIC_ADD(a.dgamma + c, sg[c] + sg[C + c] + sg[2 * C + c] + sg[3 * C + c]);
      MTUS_ATOMIC_ADD(a.dbeta + c, sb[c] + sb[C + c] + sb[2 * C + c] + sb[3 * C + c]);
      if (want_cs) MTUS_ATOMIC_ADD(a.lp_colsum + (c % Cc), sc[c] + sc[C + c] + sc[2 * C + c] + sc[3 * C + c]);
    }
  } else {
    // Wide rows (C > 1024, four warps per row): every thread owns its columns, so the CTA's partials go to shared memory
    // as they are ([3][C]); the CTAs of a CLUSTER then add their partials through distributed shared memory -- rank r
    // takes the column slice r -- and issue one coalesced vector reduction per four columns.  The first version issued
    // 48 scalar atomics per thread with a 32-byte lane stride (one L2 sector per lane): 60 us of a 114 us launch at
    // [1568, 2048] with 148 CTAs, and worse with more CTAs (MTUS_DIAG_NOATOM build: 51.7 us -> 19.6 us from 1 to 4 CTAs per SM).
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    float* sg = acc_smem; float* sb = acc_smem + C; float* sc = acc_smem + 2 * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = ((sub * NV + i) * 32 + lane) * 8;
      if (col < C) {
        *reinterpret_cast<float4*>(sg + col) = make_float4(adg[i][0], adg[i][1], adg[i][2], adg[i][3]);
        *reinterpret_cast<float4*>(sg + col + 4) = make_float4(adg[i][4], adg[i][5], adg[i][6], adg[i][7]);
        *reinterpret_cast<float4*>(sb + col) = make_float4(adb[i][0], adb[i][1], adb[i][2], adb[i][3]);
        *reinterpret_cast<float4*>(sb + col + 4) = make_float4(adb[i][4], adb[i][5], adb[i][6], adb[i][7]);
        *reinterpret_cast<float4*>(sc + col) = make_float4(acs[i][0], acs[i][1], acs[i][2], acs[i][3]);
        *reinterpret_cast<float4*>(sc + col + 4) = make_float4(acs[i][4], acs[i][5], acs[i][6], acs[i][7]);
      }
    }
    cluster.sync();
    const int nq = want_cs ? 3 : 2, C4 = C / 4;
    const int per = (C4 + CL - 1) / CL, v0 = rank * per, v1 = min(C4, v0 + per);
    const bool vec_ok = ((((uintptr_t)a.dgamma | (uintptr_t)a.dbeta | (uintptr_t)a.lp_colsum) & 15) == 0) && (Cc % 4 == 0);
    for (int q = 0; q < nq; ++q) {
      float* out = q == 0 ? a.dgamma : (q == 1 ? a.dbeta : a.lp_colsum);
      for (int v = v0 + (int)threadIdx.x; v < v1; v += blockDim.x) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < CL; ++r) {
          const float4 p = *reinterpret_cast<const float4*>(cluster.map_shared_rank(acc_smem, r) + q * C + v * 4);
          t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
        }
        const int c = q == 2 ? (v * 4) % Cc : v * 4;
        if (vec_ok) {
#ifdef MTUS_DIAG_NOATOM
          if (t.x == 1.2345e-30f)
#endif
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
        } else {
          MTUS_ATOMIC_ADD(out + (q == 2 ? (v * 4) % Cc : v * 4), t.x);
          MTUS_ATOMIC_ADD(out + (q == 2 ? (v * 4 + 1) % Cc : v * 4 + 1), t.y);
          MTUS_ATOMIC_ADD(out + (q == 2 ? (v * 4 + 2) % Cc : v * 4 + 2), t.z);
          MTUS_ATOMIC_ADD(out + (q == 2 ? (v * 4 + 3) % Cc : v * 4 + 3), t.w);
        }
      }
    }
    cluster.sync();                                       // no CTA leaves while a peer still reads its partials
  }
}

// -------------------------------------------------------------------------------------------------
// v2 kernels for C <= 1024: LPR lanes per row (a warp works on 32/LPR rows at once), NV 8-element vectors per lane,
// U row groups in flight per iteration -- every lane is busy for narrow rows (C = 96..256) and each lane has
// several 16/32-byte loads outstanding, which is what an HBM-bound kernel needs.
// -------------------------------------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TX, typename TY, int LPR, int NV, int U, int MODE>
__global__ void __launch_bounds__(128) lnv2_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, TY* __restrict__ y,
                                                       float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                       int64_t rows, int C, float eps, MergeGeom g) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  const int64_t gwarp = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 4;
  const float invC = 1.0f / (float)C;
  float gm[NV][8], bt[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * LPR + sub) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { gm[i][k] = 0.f; bt[i][k] = 0.f; }
    if (col < C) { IO<float>::load8(gamma + col, gm[i]); IO<float>::load8(beta + col, bt[i]); }
  }
  pdl_trigger();
  pdl_wait();                               // gamma / beta (parameters) were fetched while the previous kernel drained
  for (int64_t row0 = gwarp * (RPW * U); row0 < rows; row0 += nwarps * (RPW * U)) {
    float v[U][NV][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + u * RPW + grp;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * LPR + sub) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) v[u][i][k] = 0.f;
        if (row < rows && col < C) {
          bool valid; const TX* p = ln_src<TX, MODE>(x, row, col, C, g, valid);
          if (valid) IO<TX>::load8(p, v[u][i]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + u * RPW + grp;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[u][i][k];
      const float mean = group_sum<LPR>(s) * invC;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * LPR + sub) * 8;
        if (col < C) {
#pragma unroll
          for (int k = 0; k < 8; ++k) { const float d = v[u][i][k] - mean; q += d * d; }
        }
      }
      const float rstd = rsqrtf(group_sum<LPR>(q) * invC + eps);
      if (row < rows) {
        if (sub == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int col = (i * LPR + sub) * 8;
          if (col < C) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = (v[u][i][k] - mean) * rstd * gm[i][k] + bt[i][k];
            IO<TY>::store8(y + row * C + col, o);
          }
        }
      }
    }
  }
}

template <typename T, typename TDY, typename TX, int LPR, int NV, int U, int MODE>
__global__ void __launch_bounds__(128) lnv2_bwd_kernel(LnxBwdArgs a, MergeGeom g) {
  constexpr int RPW = 32 / LPR;
  extern __shared__ __align__(16) float acc_smem[];           // [4 warps][C] partials of one of dgamma / dbeta / colsum at a time
  const TDY* __restrict__ dy = reinterpret_cast<const TDY*>(a.dy);
  const TX* __restrict__ x = reinterpret_cast<const TX*>(a.x);
  T* __restrict__ dx_lp = reinterpret_cast<T*>(a.dx_lp);
  const int C = a.C;
  const int64_t rows = a.rows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  const int64_t gwarp = (int64_t)blockIdx.x * 4 + warp, nwarps = (int64_t)gridDim.x * 4;
  const float invC = 1.0f / (float)C;
  const bool want_cs = a.lp_colsum != nullptr;
  float gm[NV][8], adg[NV][8], adb[NV][8], acs[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * LPR + sub) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { gm[i][k] = 0.f; adg[i][k] = 0.f; adb[i][k] = 0.f; acs[i][k] = 0.f; }
    if (col < C) IO<float>::load8(a.gamma + col, gm[i]);
  }
  pdl_trigger();
  pdl_wait();
  // U rows per warp iteration: all the loads of the U rows (x, dy and the incoming stream gradient) are issued before
  // the first reduction, so one HBM latency is paid per iteration, not two per row
  for (int64_t row0 = gwarp * (RPW * U); row0 < rows; row0 += nwarps * (RPW * U)) {
    float xh[U][NV][8], gy[U][NV][8], rs[U][NV][8];
    int64_t off[U][NV];
    float mean[U], rstd[U], s1[U], s2[U];
    bool active[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + u * RPW + grp;
      active[u] = row < rows;
      mean[u] = active[u] ? a.mean[row] : 0.f; rstd[u] = active[u] ? a.rstd[row] : 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * LPR + sub) * 8;
        off[u][i] = -1;
#pragma unroll
        for (int k = 0; k < 8; ++k) { xh[u][i][k] = 0.f; gy[u][i][k] = 0.f; rs[u][i][k] = 0.f; }
        if (active[u] && col < C) {
          bool valid; const TX* p = ln_src<TX, MODE>(x, row, col, C, g, valid);
          if (valid) { IO<TX>::load8(p, xh[u][i]); off[u][i] = p - x; if (a.dres) IO<float>::load8(a.dres + off[u][i], rs[u][i]); }
          IO<TDY>::load8(dy + row * C + col, gy[u][i]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s1[u] = 0.f; s2[u] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * LPR + sub) * 8;
        if (active[u] && col < C) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float dv = gy[u][i][k];
            const float xn = off[u][i] >= 0 ? (xh[u][i][k] - mean[u]) * rstd[u] : (0.f - mean[u]) * rstd[u];
            xh[u][i][k] = xn;
            gy[u][i][k] = dv * gm[i][k];
            s1[u] += gy[u][i][k];
            s2[u] += gy[u][i][k] * xn;
            adg[i][k] += dv * xn;
            adb[i][k] += dv;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { s1[u] = group_sum<LPR>(s1[u]) * invC; s2[u] = group_sum<LPR>(s2[u]) * invC; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (active[u]) {
        const int64_t row = row0 + u * RPW + grp;
        const float rsc = a.lp_rowscale ? __ldg(a.lp_rowscale + row / a.rows_per_sample) : 1.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (off[u][i] >= 0) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = rstd[u] * (gy[u][i][k] - s1[u] - xh[u][i][k] * s2[u]) + rs[u][i][k];
            if (a.dx) IO<float>::store8(a.dx + off[u][i], o);
            if (dx_lp) {
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] *= rsc;
              IO<T>::store8(dx_lp + off[u][i], o);
              if (want_cs) {
                float rr[8]; IO<T>::load8_reg(o, rr);
#pragma unroll
                for (int k = 0; k < 8; ++k) acs[i][k] += rr[k];
              }
            }
          }
        }
      }
    }
  }
  // reduce the per-lane partials over the row groups of the warp, the warps of the block, then one atomic per column
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        adg[i][k] += __shfl_xor_sync(0xffffffffu, adg[i][k], o);
        adb[i][k] += __shfl_xor_sync(0xffffffffu, adb[i][k], o);
        acs[i][k] += __shfl_xor_sync(0xffffffffu, acs[i][k], o);
      }
    }
  // one quantity at a time through a [4 warps][C] buffer (16 C bytes instead of 48 C): with 8 KB per CTA at C = 512 three
  // of these CTAs still fit next to a resident persistent-GEMM CTA (200 KB) when the weight-gradient stream overlaps
  const int Cc = (MODE == 0) ? C : g.C;
#pragma unroll
  for (int qn = 0; qn < 3; ++qn) {
    if (qn == 2 && !want_cs) break;
    if (qn > 0) __syncthreads();
    if (grp == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * LPR + sub) * 8;
        if (col < C) {
          float* dst = acc_smem + warp * C + col;
          const float* src = qn == 0 ? adg[i] : (qn == 1 ? adb[i] : acs[i]);
          *reinterpret_cast<float4*>(dst) = make_float4(src[0], src[1], src[2], src[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(src[4], src[5], src[6], src[7]);
        }
      }
    }
    __syncthreads();
    float* out = qn == 0 ? a.dgamma : (qn == 1 ? a.dbeta : a.lp_colsum);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float v = acc_smem[c] + acc_smem[C + c] + acc_smem[2 * C + c] + acc_smem[3 * C + c];
      MTUS_ATOMIC_ADD(out + (qn == 2 ? (c % Cc) : c), v);
    }
  }
}

template <typename TX, typename TY, int MODE>
static int lnv2_fwd_launch(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int64_t rows, int C,
                           float eps, MergeGeom g, cudaStream_t st) {
  const int nvec = C / 8;
#define V2F(LPR_, NV_, U_)                                                                                               \
  {                                                                                                                      \
    const int rpi = (32 / LPR_) * U_ * 4;                                                                                \
    int64_t blocks = (rows + rpi - 1) / rpi;                                                                             \
    if (blocks > 148 * 8) blocks = 148 * 8;                                                                              \
    cudaError_t le = mtus_launch_pdl(lnv2_fwd_kernel<TX, TY, LPR_, NV_, U_, MODE>, dim3((int)blocks), dim3(128), 0, st, (const TX*)x, gamma, beta, (TY*)y, mean, rstd, rows, C, eps, g); \
    if (le != cudaSuccess) return (int)le;                                                                               \
  }
  if (nvec <= 4) V2F(4, 1, 4)
  else if (nvec <= 8) V2F(8, 1, 4)
  else if (nvec <= 16) V2F(16, 1, 4)
  else if (nvec <= 32) V2F(32, 1, 4)
  else if (nvec <= 64) V2F(32, 2, 2)
  else if (nvec <= 96) V2F(32, 3, 1)
  else V2F(32, 4, 1)
#undef V2F
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

static int ln_bwd_bps_override() {     // MTUS_LN_BWD_BPS: CTAs per SM of the LayerNorm backward (tuning sweeps)
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTUS_LN_BWD_BPS"); v = e ? atoi(e) : 0; }
  return v;
}

template <typename T, typename TDY, typename TX, int MODE>
static int lnv2_bwd_launch(const LnxBwdArgs& a, MergeGeom g, cudaStream_t st) {
  const int C = a.C, nvec = C / 8;
  const size_t sm = (size_t)4 * C * sizeof(float);
#define V2B(LPR_, NV_, U_, BPS_)                                                                             \
  {                                                                                                          \
    const int rpi = (32 / LPR_) * U_ * 4;                                                                    \
    int64_t blocks = (a.rows + rpi - 1) / rpi;                                                               \
    const int bps = ln_bwd_bps_override() > 0 ? ln_bwd_bps_override() : BPS_;                                \
    if (blocks > 148 * bps) blocks = 148 * bps;                                                              \
    auto kern = lnv2_bwd_kernel<T, TDY, TX, LPR_, NV_, U_, MODE>;                                            \
    if (sm > 48 * 1024) {                                                                                    \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);      \
      if (e != cudaSuccess) return (int)e;                                                                   \
    }                                                                                                        \
    cudaError_t le = mtus_launch_pdl(kern, dim3((int)blocks), dim3(128), sm, st, a, g);                      \
    if (le != cudaSuccess) return (int)le;                                                                   \
  }
  // 138 registers x 128 threads: three CTAs per SM are resident, so three per SM is exactly one wave (six left a second wave
  // that pays the prologue / column-sum epilogue again: 41.8 -> 40.4 us at [100352, 128], 24.4 -> 23.2 us at [25088, 256])
  if (nvec <= 4) V2B(4, 1, 2, 3)
  else if (nvec <= 8) V2B(8, 1, 2, 3)
  else if (nvec <= 16) V2B(16, 1, 2, 3)
  else if (nvec <= 32) V2B(32, 1, 2, 3)
  else if (nvec <= 64) {
    // one row per warp iteration, 162 registers, 3 CTAs per SM: 13.1 us at [6272, 512] against 15.5 us for the two-row form
    // (214 registers, 2 CTAs per SM), which stays selectable with MTUS_LN_BWD_U=2
    static int u2 = -1;
    if (u2 < 0) { const char* e = getenv("MTUS_LN_BWD_U"); u2 = (e && atoi(e) == 2) ? 1 : 0; }
    if (u2) V2B(32, 2, 2, 2)
    else V2B(32, 2, 1, 3)
  }
  else if (nvec <= 96) V2B(32, 3, 1, 2)      // 206-255 registers: two CTAs per SM are resident, a third would wait for a second wave
  else V2B(32, 4, 1, 2)
#undef V2B
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

template <typename TX, typename TY, int MODE>
static int lnx_fwd_launch(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int64_t rows, int C,
                          float eps, MergeGeom g, cudaStream_t st) {
  if (rows == 0) return MTUS_OK;
  MTUS_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 4096);
  if (C <= 1024) return lnv2_fwd_launch<TX, TY, MODE>(x, gamma, beta, y, mean, rstd, rows, C, eps, g, st);
  const int nvec = C / 8;
#define LNX_CASE(NV_, WPR_)                                                                                                \
  {                                                                                                                        \
    const int rpb = 4 / WPR_;                                                                                              \
    int64_t blocks = (rows + rpb - 1) / rpb;                                                                               \
    if (blocks > 148 * 16) blocks = 148 * 16;                                                                              \
    lnx_fwd_kernel<TX, TY, NV_, WPR_, MODE><<<(int)blocks, 128, 0, st>>>((const TX*)x, gamma, beta, (TY*)y, mean, rstd, rows, C, eps, g); \
  }
  if (nvec <= 32) LNX_CASE(1, 1)
  else if (nvec <= 64) LNX_CASE(2, 1)
  else if (nvec <= 96) LNX_CASE(3, 1)
  else if (nvec <= 128) LNX_CASE(1, 4)
  else if (nvec <= 256) LNX_CASE(2, 4)
  else if (nvec <= 384) LNX_CASE(3, 4)
  else LNX_CASE(4, 4)
#undef LNX_CASE
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

static int lnx_bwd_wide_bps() {        // MTUS_LNX_BWD_BPS: CTAs per SM of the wide-row (C > 1024) LayerNorm backward
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTUS_LNX_BWD_BPS"); v = e ? atoi(e) : 4; if (v < 1) v = 1; }
  return v;
}

template <typename T, typename TDY, typename TX, int MODE>
static int lnx_bwd_launch(const LnxBwdArgs& a, MergeGeom g, cudaStream_t st) {
  if (a.rows == 0) return MTUS_OK;
  const int C = a.C;
  MTUS_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 4096);
  if (C <= 1024) return lnv2_bwd_launch<T, TDY, TX, MODE>(a, g, st);
  const int nvec = C / 8;
#define LNX_CASE(NV_, WPR_)                                                                                  \
  {                                                                                                          \
    const int rpb = 4 / WPR_;                                                                                \
    int64_t blocks = (a.rows + rpb - 1) / rpb;                                                               \
    if (blocks > (WPR_ == 1 ? 148 * 4 : 148 * lnx_bwd_wide_bps())) blocks = (WPR_ == 1 ? 148 * 4 : 148 * lnx_bwd_wide_bps()); \
    const int cl = (WPR_ == 1) ? 1 : (blocks >= 8 ? 8 : 1);        /* wide rows: clusters of 8 CTAs add their partials */ \
    blocks = (blocks + cl - 1) / cl * cl;                                                                    \
    const size_t sm = (WPR_ == 1) ? (size_t)12 * C * sizeof(float) : (size_t)3 * C * sizeof(float);          \
    auto kern = lnx_bwd_kernel<T, TDY, TX, NV_, WPR_, MODE>;                                                 \
    if (sm > 48 * 1024) {                                                                                    \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);      \
      if (e != cudaSuccess) return (int)e;                                                                   \
    }                                                                                                        \
    cudaLaunchConfig_t cfg = {};                                                                             \
    cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = sm; cfg.stream = st; \
    cudaLaunchAttribute attr[1];                                                                             \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                        \
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;               \
    cfg.attrs = attr; cfg.numAttrs = 1;                                                                      \
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, a, g);                                                   \
    if (le != cudaSuccess) return (int)le;                                                                   \
  }
  if (nvec <= 32) LNX_CASE(1, 1)
  else if (nvec <= 64) LNX_CASE(2, 1)
  else if (nvec <= 96) LNX_CASE(3, 1)
  else if (nvec <= 128) LNX_CASE(1, 4)
  else if (nvec <= 256) LNX_CASE(2, 4)
  else if (nvec <= 384) LNX_CASE(3, 4)
  else LNX_CASE(4, 4)
#undef LNX_CASE
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_layernorm_fwd_mixed(const void* x, int x_f32, const float* gamma, const float* beta, void* y, int y_f32,
                                        float* mean, float* rstd, int64_t rows, int C, float eps, int dtype, void* stream) {
  MTUS_CHECK_ARG(x && gamma && beta && y && mean && rstd && rows >= 0);
  MergeGeom g{};
  cudaStream_t st = (cudaStream_t)stream;
  const bool xf = x_f32 || dtype == MTUS_F32, yf = y_f32 || dtype == MTUS_F32;
  if (xf && yf) return lnx_fwd_launch<float, float, 0>(x, gamma, beta, y, mean, rstd, rows, C, eps, g, st);
  if (dtype != MTUS_BF16) return MTUS_ERR_UNSUPPORTED;
  if (xf && !yf) return lnx_fwd_launch<float, bf16, 0>(x, gamma, beta, y, mean, rstd, rows, C, eps, g, st);
  if (!xf && yf) return lnx_fwd_launch<bf16, float, 0>(x, gamma, beta, y, mean, rstd, rows, C, eps, g, st);
  return mtus_layernorm_fwd(x, gamma, beta, y, mean, rstd, rows, C, eps, dtype, stream);
}

// x: fp32 [B,H,W,C] (the residual stream); y: dtype [B,Ho,Wo,4C]
extern "C" int mtus_patch_merge_ln_fwd_mixed(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                             int B, int H, int W, int C, float eps, int dtype, void* stream) {
  MTUS_CHECK_ARG(x && gamma && beta && y && mean && rstd && B >= 0 && H > 0 && W > 0 && C % 8 == 0);
  MergeGeom g{B, H, W, C, (H + 1) / 2, (W + 1) / 2};
  const int64_t rows = (int64_t)B * g.Ho * g.Wo;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) return lnx_fwd_launch<float, float, 1>(x, gamma, beta, y, mean, rstd, rows, 4 * C, eps, g, st);
  if (dtype == MTUS_BF16) return lnx_fwd_launch<float, bf16, 1>(x, gamma, beta, y, mean, rstd, rows, 4 * C, eps, g, st);
  return MTUS_ERR_UNSUPPORTED;
}

extern "C" int mtus_layernorm_bwd_mixed(const void* dy, int dy_f32, const void* x, int x_f32, const float* gamma, const float* mean,
                                        const float* rstd, const float* dres, float* dx, void* dx_lp, const float* lp_rowscale,
                                        int rows_per_sample, float* lp_colsum, float* dgamma, float* dbeta, int64_t rows, int C,
                                        int dtype, void* stream) {
  MTUS_CHECK_ARG(dy && x && gamma && mean && rstd && (dx || dx_lp) && dgamma && dbeta && rows >= 0);
  MTUS_CHECK_ARG(!lp_colsum || dx_lp);
  LnxBwdArgs a{dy, x, gamma, mean, rstd, dres, dx, dx_lp, lp_rowscale, rows_per_sample > 0 ? rows_per_sample : 1, lp_colsum,
               dgamma, dbeta, rows, C};
  MergeGeom g{};
  cudaStream_t st = (cudaStream_t)stream;
  const bool df = dy_f32 || dtype == MTUS_F32, xf = x_f32 || dtype == MTUS_F32;
  if (dtype == MTUS_F32) return lnx_bwd_launch<float, float, float, 0>(a, g, st);
  if (dtype != MTUS_BF16) return MTUS_ERR_UNSUPPORTED;
  if (!df && xf) return lnx_bwd_launch<bf16, bf16, float, 0>(a, g, st);
  if (df && !xf) return lnx_bwd_launch<bf16, float, bf16, 0>(a, g, st);
  return MTUS_ERR_UNSUPPORTED;
}

// dy: dtype [B,Ho,Wo,4C]; x / dres / dx: fp32 [B,H,W,C]; dx_lp: dtype [B,H,W,C]; lp_colsum: [C]
extern "C" int mtus_patch_merge_ln_bwd_mixed(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                             const float* dres, float* dx, void* dx_lp, const float* lp_rowscale, int rows_per_sample,
                                             float* lp_colsum, float* dgamma, float* dbeta, int B, int H, int W, int C, int dtype,
                                             void* stream) {
  MTUS_CHECK_ARG(dy && x && gamma && mean && rstd && (dx || dx_lp) && dgamma && dbeta && B >= 0 && H > 0 && W > 0 && C % 8 == 0);
  MergeGeom g{B, H, W, C, (H + 1) / 2, (W + 1) / 2};
  (void)rows_per_sample;                                      // one sample = Ho*Wo gathered rows
  LnxBwdArgs a{dy, x, gamma, mean, rstd, dres, dx, dx_lp, lp_rowscale, g.Ho * g.Wo, lp_colsum,
               dgamma, dbeta, (int64_t)B * g.Ho * g.Wo, 4 * C};
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) return lnx_bwd_launch<float, float, float, 1>(a, g, st);
  if (dtype == MTUS_BF16) return lnx_bwd_launch<bf16, bf16, float, 1>(a, g, st);
  return MTUS_ERR_UNSUPPORTED;
}
