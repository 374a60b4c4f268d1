// FPN decoder bandwidth-bound pieces, all on NHWC maps (smp FPNDecoder):
//   * nearest-x2 top-down add (FPNBlock) and its backward (2x2 sum),
//   * GroupNorm(32) statistics / normalise+ReLU / backward (Conv3x3GNReLU),
//   * bilinear x2 with align_corners=True and its (gather-form, atomic-free) backward,
//   * cat/add merge fused with the Dropout2d channel scale and the NHWC->NCHW transpose,
//   * 3x3 weight repacking for the implicit-GEMM convolutions.
// Replaces aten upsample_nearest2d/add, native_group_norm(+backward), relu, upsample_bilinear2d,
// cat, feature_dropout_ that eager PyTorch runs for smp (reached from
// /root/reference/code/models/decoders.py:42-49).  HBM roofline; 16-byte vectorised, coalesced.
#include "common.cuh"

static inline int grid_for(int64_t work_items, int threads, int max_blocks = 148 * 16) {
  int64_t b = (work_items + threads - 1) / threads;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- nearest x2 + add ----------------------------------------------------------------------------
template <typename T>
__global__ void upadd_fwd_kernel(const T* __restrict__ skip, const T* __restrict__ top, T* __restrict__ y, int B, int H,
                                 int W, int C8) {
  const int64_t total = (int64_t)B * H * W * C8, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C8); int64_t p = i / C8;
    const int x = (int)(p % W); p /= W; const int yy = (int)(p % H); const int64_t b = p / H;
    const int64_t tp = ((b * (H / 2) + yy / 2) * (W / 2) + x / 2) * C8 + c;
    float a[8], t[8];
    IO<T>::load8(skip + i * 8, a);
    IO<T>::load8(top + tp * 8, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += t[k];
    IO<T>::store8(y + i * 8, a);
  }
}

template <typename T>
__global__ void upadd_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dtop, int acc, int B, int H, int W, int C8) {
  const int Ht = H / 2, Wt = W / 2;
  const int64_t total = (int64_t)B * Ht * Wt * C8, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C8); int64_t p = i / C8;
    const int x = (int)(p % Wt); p /= Wt; const int yy = (int)(p % Ht); const int64_t b = p / Ht;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (acc) IO<T>::load8(dtop + i * 8, s);
#pragma unroll
    for (int dyy = 0; dyy < 2; ++dyy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float v[8];
        IO<T>::load8(dy + (((b * H + 2 * yy + dyy) * W + 2 * x + dx) * (int64_t)C8 + c) * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] += v[k];
      }
    IO<T>::store8(dtop + i * 8, s);
  }
}

extern "C" int mtus_upsample_add_fwd(const void* skip, const void* top, void* y, int B, int H, int W, int C, int dtype,
                                     void* stream) {
  MTUS_CHECK_ARG(skip && top && y && H % 2 == 0 && W % 2 == 0 && C % 8 == 0);
  const int64_t n = (int64_t)B * H * W * (C / 8);
  if (n == 0) return MTUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) upadd_fwd_kernel<float><<<grid_for(n, 256), 256, 0, st>>>((const float*)skip, (const float*)top, (float*)y, B, H, W, C / 8);
  else if (dtype == MTUS_BF16) upadd_fwd_kernel<bf16><<<grid_for(n, 256), 256, 0, st>>>((const bf16*)skip, (const bf16*)top, (bf16*)y, B, H, W, C / 8);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_upsample_add_bwd(const void* dy, void* dtop, int accumulate, int B, int H, int W, int C, int dtype,
                                     void* stream) {
  MTUS_CHECK_ARG(dy && dtop && H % 2 == 0 && W % 2 == 0 && C % 8 == 0);
  const int64_t n = (int64_t)B * (H / 2) * (W / 2) * (C / 8);
  if (n == 0) return MTUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) upadd_bwd_kernel<float><<<grid_for(n, 256), 256, 0, st>>>((const float*)dy, (float*)dtop, accumulate, B, H, W, C / 8);
  else if (dtype == MTUS_BF16) upadd_bwd_kernel<bf16><<<grid_for(n, 256), 256, 0, st>>>((const bf16*)dy, (bf16*)dtop, accumulate, B, H, W, C / 8);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// ---- GroupNorm ---------------------------------------------------------------------------------
// Block = 256 threads = (C/8 vector lanes) x (256/(C/8) pixel lanes); grid (chunks, B).
// PASS 0: sum -> mean_acc[b,g];  PASS 1: centred sum of squares (mean = mean_acc/n) -> var_acc[b,g].
template <typename T, int PASS>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, float* __restrict__ mean_acc,
                                                       float* __restrict__ var_acc, int HW, int C, int G, int pix_per_chunk) {
  extern __shared__ float sred[];  // [C]
  const int C8 = C / 8, cpg = C / G;
  const int v = threadIdx.x % C8, pl = threadIdx.x / C8, npl = 256 / C8;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_chunk, p1 = min(HW, p0 + pix_per_chunk);
  const float inv_n = 1.0f / ((float)HW * cpg);
  float mu[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) mu[k] = (PASS == 1) ? mean_acc[b * G + (v * 8 + k) / cpg] * inv_n : 0.f;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pl < npl) {
    for (int p = p0 + pl; p < p1; p += npl) {
      float a[8];
      IO<T>::load8(x + ((int64_t)b * HW + p) * C + v * 8, a);
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = a[k] - mu[k]; acc[k] += (PASS == 1) ? d * d : a[k]; }
    }
  }
  for (int c = threadIdx.x; c < C; c += 256) sred[c] = 0.f;
  __syncthreads();
  if (pl < npl) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&sred[v * 8 + k], acc[k]);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += 256) {
    float s = 0.f;
    for (int k = 0; k < cpg; ++k) s += sred[g * cpg + k];
    atomicAdd((PASS == 1 ? var_acc : mean_acc) + b * G + g, s);
  }
}

__global__ void gn_finalize_kernel(float* mean, float* rstd, int n, float inv_cnt, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { mean[i] *= inv_cnt; rstd[i] = rsqrtf(rstd[i] * inv_cnt + eps); }
}

static int gn_chunks(int B, int HW, int C, int& ppc) {
  const int npl = 256 / (C / 8);
  int chunks = (148 * 4 + B - 1) / (B > 0 ? B : 1);
  const int maxc = (HW + npl * 4 - 1) / (npl * 4);
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  ppc = (HW + chunks - 1) / chunks;
  return (HW + ppc - 1) / ppc;
}

extern "C" int mtus_groupnorm_stats(const void* x, float* mean, float* rstd, int B, int HW, int C, int G, float eps,
                                    int dtype, void* stream) {
  MTUS_CHECK_ARG(x && mean && rstd && C % 8 == 0 && C / 8 <= 256 && G > 0 && C % G == 0 && B <= 65535);
  if (B == 0) return MTUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(mean, 0, sizeof(float) * B * G, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(rstd, 0, sizeof(float) * B * G, st);
  if (e != cudaSuccess) return (int)e;
  int ppc; const int chunks = gn_chunks(B, HW, C, ppc);
  dim3 grid(chunks, B);
  const size_t sm = sizeof(float) * C;
  if (dtype == MTUS_F32) {
    gn_stats_kernel<float, 0><<<grid, 256, sm, st>>>((const float*)x, mean, rstd, HW, C, G, ppc);
    gn_stats_kernel<float, 1><<<grid, 256, sm, st>>>((const float*)x, mean, rstd, HW, C, G, ppc);
  } else if (dtype == MTUS_BF16) {
    gn_stats_kernel<bf16, 0><<<grid, 256, sm, st>>>((const bf16*)x, mean, rstd, HW, C, G, ppc);
    gn_stats_kernel<bf16, 1><<<grid, 256, sm, st>>>((const bf16*)x, mean, rstd, HW, C, G, ppc);
  } else return MTUS_ERR_UNSUPPORTED;
  gn_finalize_kernel<<<ceil_div(B * G, 256), 256, 0, st>>>(mean, rstd, B * G, 1.0f / ((float)HW * (C / G)), eps);
  MTUS_LAUNCH_STATUS_N(3);
  return MTUS_OK;
}

// ACT 0: ReLU (smp Conv3x3GNReLU) | ACT 1: SiLU (the reference's segmentation head, code/models/heads.py:16-42)
__device__ __forceinline__ float gn_sigmoid(float z) { return 1.0f / (1.0f + __expf(-z)); }

template <typename T, int ACT>
__global__ void gn_relu_fwd_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ y,
                                   int64_t total, int HW, int C8, int G, int cpg) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = (int)(i % C8);
    const int64_t b = i / ((int64_t)C8 * HW);
    float a[8], gm[8], bt[8];
    IO<T>::load8(x + i * 8, a);
    IO<float>::load8(gamma + v * 8, gm);
    IO<float>::load8(beta + v * 8, bt);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int g = (v * 8 + k) / cpg;
      const float m = __ldg(mean + b * G + g), r = __ldg(rstd + b * G + g);
      const float z = (a[k] - m) * r * gm[k] + bt[k];
      a[k] = ACT == 0 ? fmaxf(z, 0.f) : z * gn_sigmoid(z);
    }
    IO<T>::store8(y + i * 8, a);
  }
}

extern "C" int mtus_groupnorm_act_fwd(const void* x, const float* mean, const float* rstd, const float* gamma,
                                      const float* beta, void* y, int B, int HW, int C, int G, int act, int dtype, void* stream) {
  MTUS_CHECK_ARG(x && mean && rstd && gamma && beta && y && C % 8 == 0 && G > 0 && C % G == 0 && (act == 0 || act == 1));
  const int64_t total = (int64_t)B * HW * (C / 8);
  if (total == 0) return MTUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define GN_FWD(T_, ACT_) gn_relu_fwd_kernel<T_, ACT_><<<grid_for(total, 256), 256, 0, st>>>((const T_*)x, mean, rstd, gamma, beta, (T_*)y, total, HW, C / 8, G, C / G)
  if (dtype == MTUS_F32) { if (act == 0) GN_FWD(float, 0); else GN_FWD(float, 1); }
  else if (dtype == MTUS_BF16) { if (act == 0) GN_FWD(bf16, 0); else GN_FWD(bf16, 1); }
  else return MTUS_ERR_UNSUPPORTED;
#undef GN_FWD
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_groupnorm_relu_fwd(const void* x, const float* mean, const float* rstd, const float* gamma,
                                       const float* beta, void* y, int B, int HW, int C, int G, int dtype, void* stream) {
  return mtus_groupnorm_act_fwd(x, mean, rstd, gamma, beta, y, B, HW, C, G, 0, dtype, stream);
}

// d act(z) / dz times dy.  ReLU: mask from the saved output y; SiLU: z recomputed from x (sigma(z) (1 + z (1 - sigma(z)))).
template <int ACT>
__device__ __forceinline__ float gn_act_grad(float dy, float yv, float xh, float gm, float bt) {
  if (ACT == 0) return yv > 0.f ? dy : 0.f;
  const float z = xh * gm + bt, sg = gn_sigmoid(z);
  return dy * sg * (1.0f + z * (1.0f - sg));
}

// backward pass 1: per (b,g) s1 = sum dyr*gamma, s2 = sum dyr*gamma*xhat (ws[0..BG), ws[BG..2BG));
//                  per channel dgamma += sum dyr*xhat, dbeta += sum dyr     (dyr = dy * (y > 0))
template <typename T, int ACT>
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ ws,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, int B, int HW,
                                                            int C, int G, int pix_per_chunk) {
  extern __shared__ float sred[];  // [4][C]: s1, s2 (per channel, folded to groups later), dgamma, dbeta
  const int C8 = C / 8, cpg = C / G;
  const int v = threadIdx.x % C8, pl = threadIdx.x / C8, npl = 256 / C8;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_chunk, p1 = min(HW, p0 + pix_per_chunk);
  float mu[8], rs[8], gm[8], bt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int g = (v * 8 + k) / cpg;
    mu[k] = mean[b * G + g]; rs[k] = rstd[b * G + g]; gm[k] = gamma[v * 8 + k]; bt[k] = ACT == 0 ? 0.f : beta[v * 8 + k];
  }
  float a1[8], a2[8], ag[8], ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a1[k] = a2[k] = ag[k] = ab[k] = 0.f;
  if (pl < npl) {
    for (int p = p0 + pl; p < p1; p += npl) {
      const int64_t off = ((int64_t)b * HW + p) * C + v * 8;
      float d[8], xv[8], yv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      IO<T>::load8(dy + off, d); IO<T>::load8(x + off, xv);
      if (ACT == 0) IO<T>::load8(y + off, yv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (xv[k] - mu[k]) * rs[k];
        const float dr = gn_act_grad<ACT>(d[k], yv[k], xh, gm[k], bt[k]);
        a1[k] += dr * gm[k]; a2[k] += dr * gm[k] * xh; ag[k] += dr * xh; ab[k] += dr;
      }
    }
  }
  for (int c = threadIdx.x; c < 4 * C; c += 256) sred[c] = 0.f;
  __syncthreads();
  if (pl < npl) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&sred[v * 8 + k], a1[k]); atomicAdd(&sred[C + v * 8 + k], a2[k]);
      atomicAdd(&sred[2 * C + v * 8 + k], ag[k]); atomicAdd(&sred[3 * C + v * 8 + k], ab[k]);
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += 256) {
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < cpg; ++k) { s1 += sred[g * cpg + k]; s2 += sred[C + g * cpg + k]; }
    atomicAdd(ws + b * G + g, s1);
    atomicAdd(ws + B * G + b * G + g, s2);
  }
  for (int c = threadIdx.x; c < C; c += 256) { atomicAdd(dgamma + c, sred[2 * C + c]); atomicAdd(dbeta + c, sred[3 * C + c]); }
}

template <typename T, int ACT>
__global__ void gn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ ws, T* __restrict__ dx, int64_t total, int B, int HW, int C8, int G,
                                    int cpg) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float inv_n = 1.0f / ((float)HW * cpg);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = (int)(i % C8);
    const int64_t b = i / ((int64_t)C8 * HW);
    float d[8], xv[8], yv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gm[8], bt[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, o[8];
    IO<T>::load8(dy + i * 8, d); IO<T>::load8(x + i * 8, xv);
    if (ACT == 0) IO<T>::load8(y + i * 8, yv); else IO<float>::load8(beta + v * 8, bt);
    IO<float>::load8(gamma + v * 8, gm);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int g = (v * 8 + k) / cpg;
      const float m = __ldg(mean + b * G + g), r = __ldg(rstd + b * G + g);
      const float s1 = __ldg(ws + b * G + g) * inv_n, s2 = __ldg(ws + (int64_t)B * G + b * G + g) * inv_n;
      const float xh = (xv[k] - m) * r;
      const float dr = gn_act_grad<ACT>(d[k], yv[k], xh, gm[k], bt[k]);
      o[k] = r * (dr * gm[k] - s1 - xh * s2);
    }
    IO<T>::store8(dx + i * 8, o);
  }
}

extern "C" int mtus_groupnorm_act_bwd(const void* dy, const void* x, const void* y, const float* mean, const float* rstd,
                                      const float* gamma, const float* beta, void* dx, float* dgamma, float* dbeta, float* ws,
                                      int B, int HW, int C, int G, int act, int dtype, void* stream) {
  MTUS_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && ws && (act == 0 || act == 1));
  MTUS_CHECK_ARG(act == 0 ? y != nullptr : beta != nullptr);
  MTUS_CHECK_ARG(C % 8 == 0 && C / 8 <= 256 && G > 0 && C % G == 0 && B <= 65535);
  if (B == 0) return MTUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(float) * 2 * B * G, st);
  if (e != cudaSuccess) return (int)e;
  int ppc; const int chunks = gn_chunks(B, HW, C, ppc);
  dim3 grid(chunks, B);
  const size_t sm = sizeof(float) * 4 * C;
  const int64_t total = (int64_t)B * HW * (C / 8);
#define GN_BWD(T_, ACT_)                                                                                                                  \
  {                                                                                                                                       \
    gn_bwd_reduce_kernel<T_, ACT_><<<grid, 256, sm, st>>>((const T_*)dy, (const T_*)x, (const T_*)y, mean, rstd, gamma, beta, ws, dgamma, dbeta, B, HW, C, G, ppc); \
    gn_bwd_apply_kernel<T_, ACT_><<<grid_for(total, 256), 256, 0, st>>>((const T_*)dy, (const T_*)x, (const T_*)y, mean, rstd, gamma, beta, ws, (T_*)dx, total, B, HW, C / 8, G, C / G); \
  }
  if (dtype == MTUS_F32) { if (act == 0) GN_BWD(float, 0) else GN_BWD(float, 1) }
  else if (dtype == MTUS_BF16) { if (act == 0) GN_BWD(bf16, 0) else GN_BWD(bf16, 1) }
  else return MTUS_ERR_UNSUPPORTED;
#undef GN_BWD
  MTUS_LAUNCH_STATUS_N(2);
  return MTUS_OK;
}

// ---- BatchNorm2d + activation over NHWC rows [M = B*H*W, C] (the reference's baseline detection head stacks
// Conv3x3 -> BatchNorm2d -> ReLU twice, code/models/heads.py:404-428) ---------------------------------
// Per-channel statistics over all rows are GroupNorm statistics with one "sample" of M pixels and one group per channel,
// so the GroupNorm kernels above serve unchanged: stats (two-pass, centred), fused normalise + activation, and the
// two-kernel backward.  training = 0 (running statistics): the statistics are constants, so the backward's mean terms
// (ws) are zeroed between the reduce and the apply kernel; dgamma / dbeta are the same sums either way.
extern "C" int mtus_batchnorm_stats(const void* x, float* mean, float* rstd, int64_t M, int C, float eps, int dtype, void* stream) {
  MTUS_CHECK_ARG(M >= 0 && M < (1ll << 31));
  return mtus_groupnorm_stats(x, mean, rstd, 1, (int)M, C, C, eps, dtype, stream);
}

extern "C" int mtus_batchnorm_act_fwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                      void* y, int64_t M, int C, int act, int dtype, void* stream) {
  MTUS_CHECK_ARG(M >= 0 && M < (1ll << 31));
  return mtus_groupnorm_act_fwd(x, mean, rstd, gamma, beta, y, 1, (int)M, C, C, act, dtype, stream);
}

extern "C" int mtus_batchnorm_act_bwd(const void* dy, const void* x, const void* y, const float* mean, const float* rstd,
                                      const float* gamma, const float* beta, void* dx, float* dgamma, float* dbeta, float* ws,
                                      int64_t M, int C, int act, int training, int dtype, void* stream) {
  MTUS_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && ws && (act == 0 || act == 1));
  MTUS_CHECK_ARG(act == 0 ? y != nullptr : beta != nullptr);
  MTUS_CHECK_ARG(C % 8 == 0 && C / 8 <= 256 && M >= 0 && M < (1ll << 31));
  if (M == 0) return MTUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int B = 1, HW = (int)M, G = C;
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(float) * 2 * B * G, st);
  if (e != cudaSuccess) return (int)e;
  int ppc; const int chunks = gn_chunks(B, HW, C, ppc);
  dim3 grid(chunks, B);
  const size_t sm = sizeof(float) * 4 * C;
  const int64_t total = (int64_t)HW * (C / 8);
#define BN_BWD(T_, ACT_)                                                                                                                  \
  {                                                                                                                                       \
    gn_bwd_reduce_kernel<T_, ACT_><<<grid, 256, sm, st>>>((const T_*)dy, (const T_*)x, (const T_*)y, mean, rstd, gamma, beta, ws, dgamma, dbeta, B, HW, C, G, ppc); \
    if (!training) { e = cudaMemsetAsync(ws, 0, sizeof(float) * 2 * B * G, st); if (e != cudaSuccess) return (int)e; }                   \
    gn_bwd_apply_kernel<T_, ACT_><<<grid_for(total, 256), 256, 0, st>>>((const T_*)dy, (const T_*)x, (const T_*)y, mean, rstd, gamma, beta, ws, (T_*)dx, total, B, HW, C / 8, G, C / G); \
  }
  if (dtype == MTUS_F32) { if (act == 0) BN_BWD(float, 0) else BN_BWD(float, 1) }
  else if (dtype == MTUS_BF16) { if (act == 0) BN_BWD(bf16, 0) else BN_BWD(bf16, 1) }
  else return MTUS_ERR_UNSUPPORTED;
#undef BN_BWD
  MTUS_LAUNCH_STATUS_N(2);
  return MTUS_OK;
}

extern "C" int mtus_groupnorm_relu_bwd(const void* dy, const void* x, const void* y, const float* mean, const float* rstd,
                                       const float* gamma, void* dx, float* dgamma, float* dbeta, float* ws, int B, int HW,
                                       int C, int G, int dtype, void* stream) {
  MTUS_CHECK_ARG(y);
  return mtus_groupnorm_act_bwd(dy, x, y, mean, rstd, gamma, nullptr, dx, dgamma, dbeta, ws, B, HW, C, G, 0, dtype, stream);
}

// ---- bilinear x2, align_corners=True (aten upsample_bilinear2d semantics) ------------------------
__device__ __forceinline__ void bil_src(int o, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  const float r = scale * (float)o;
  i0 = (int)r;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = r - (float)i0;
  l0 = 1.0f - l1;
}

template <typename T>
__global__ void bilinear_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C8, float sy, float sx) {
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t total = (int64_t)B * Ho * Wo * C8, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C8); int64_t p = i / C8;
    const int ox = (int)(p % Wo); p /= Wo; const int oy = (int)(p % Ho); const int64_t b = p / Ho;
    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
    bil_src(oy, sy, H, y0, y1, ly0, ly1);
    bil_src(ox, sx, W, x0, x1, lx0, lx1);
    const T* base = x + b * H * (int64_t)W * C8 * 8 + c * 8;
    float a[8], bq[8], cc[8], d[8], o[8];
    IO<T>::load8(base + ((int64_t)y0 * W + x0) * C8 * 8, a);
    IO<T>::load8(base + ((int64_t)y0 * W + x1) * C8 * 8, bq);
    IO<T>::load8(base + ((int64_t)y1 * W + x0) * C8 * 8, cc);
    IO<T>::load8(base + ((int64_t)y1 * W + x1) * C8 * 8, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = ly0 * (lx0 * a[k] + lx1 * bq[k]) + ly1 * (lx0 * cc[k] + lx1 * d[k]);
    IO<T>::store8(y + i * 8, o);
  }
}

// gather-form backward: each input pixel collects from the output pixels whose 4 taps touch it
template <typename T>
__global__ void bilinear_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int H, int W, int C8, float sy, float sx) {
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t total = (int64_t)B * H * W * C8, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C8); int64_t p = i / C8;
    const int ix = (int)(p % W); p /= W; const int iy = (int)(p % H); const int64_t b = p / H;
    // candidate output range: r = s*o in (i-1, i+1)  =>  o in ((i-1)/s, (i+1)/s); widened by 1 for rounding
    int oy_lo = (sy > 0.f) ? (int)floorf((float)(iy - 1) / sy) - 1 : 0, oy_hi = (sy > 0.f) ? (int)ceilf((float)(iy + 1) / sy) + 1 : Ho - 1;
    int ox_lo = (sx > 0.f) ? (int)floorf((float)(ix - 1) / sx) - 1 : 0, ox_hi = (sx > 0.f) ? (int)ceilf((float)(ix + 1) / sx) + 1 : Wo - 1;
    oy_lo = max(oy_lo, 0); oy_hi = min(oy_hi, Ho - 1); ox_lo = max(ox_lo, 0); ox_hi = min(ox_hi, Wo - 1);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1; float ly0, ly1;
      bil_src(oy, sy, H, y0, y1, ly0, ly1);
      float wy = 0.f;
      if (y0 == iy) wy += ly0;
      if (y1 == iy) wy += ly1;
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1; float lx0, lx1;
        bil_src(ox, sx, W, x0, x1, lx0, lx1);
        float wx = 0.f;
        if (x0 == ix) wx += lx0;
        if (x1 == ix) wx += lx1;
        if (wx == 0.f) continue;
        float v[8];
        IO<T>::load8(dy + (((b * Ho + oy) * (int64_t)Wo + ox) * C8 + c) * 8, v);
        const float wgt = wy * wx;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += wgt * v[k];
      }
    }
    IO<T>::store8(dx + i * 8, acc);
  }
}

static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

extern "C" int mtus_bilinear2x_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
  MTUS_CHECK_ARG(x && y && H > 0 && W > 0 && C % 8 == 0);
  const int64_t n = (int64_t)B * 4 * H * W * (C / 8);
  if (n == 0) return MTUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float sy = ac_scale(H, 2 * H), sx = ac_scale(W, 2 * W);
  if (dtype == MTUS_F32) bilinear_fwd_kernel<float><<<grid_for(n, 256), 256, 0, st>>>((const float*)x, (float*)y, B, H, W, C / 8, sy, sx);
  else if (dtype == MTUS_BF16) bilinear_fwd_kernel<bf16><<<grid_for(n, 256), 256, 0, st>>>((const bf16*)x, (bf16*)y, B, H, W, C / 8, sy, sx);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_bilinear2x_bwd(const void* dy, void* dx, int B, int H, int W, int C, int dtype, void* stream) {
  MTUS_CHECK_ARG(dy && dx && H > 0 && W > 0 && C % 8 == 0);
  const int64_t n = (int64_t)B * H * W * (C / 8);
  if (n == 0) return MTUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float sy = ac_scale(H, 2 * H), sx = ac_scale(W, 2 * W);
  if (dtype == MTUS_F32) bilinear_bwd_kernel<float><<<grid_for(n, 256), 256, 0, st>>>((const float*)dy, (float*)dx, B, H, W, C / 8, sy, sx);
  else if (dtype == MTUS_BF16) bilinear_bwd_kernel<bf16><<<grid_for(n, 256), 256, 0, st>>>((const bf16*)dy, (bf16*)dx, B, H, W, C / 8, sy, sx);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// ---- merge (cat / add) + Dropout2d scale + NHWC -> NCHW --------------------------------------------
struct MergePtrs { const void* p[4]; };
struct MergeOutPtrs { void* p[4]; };

template <typename T, typename TO>
__global__ void __launch_bounds__(256) merge_fwd_kernel(MergePtrs src, int nsrc, int cat, const float* __restrict__ chanscale,
                                                        const float* __restrict__ chanshift, TO* __restrict__ out, int HW, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int Cout = cat ? nsrc * C : C;
  const int r0 = blockIdx.y * 32;           // pixel tile
  const int c0g = blockIdx.x * 32;          // output-channel tile (global over the concatenation)
  const int s = cat ? c0g / C : 0, c0 = cat ? c0g - s * C : c0g;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    float v = 0.f;
    if (r < HW && c < C) {
      if (cat) v = IO<T>::ld(reinterpret_cast<const T*>(src.p[s]) + ((int64_t)b * HW + r) * C + c);
      else for (int k = 0; k < nsrc; ++k) v += IO<T>::ld(reinterpret_cast<const T*>(src.p[k]) + ((int64_t)b * HW + r) * C + c);
    }
    tile[ty + i * 8][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, r = r0 + tx;
    if (r < HW && c < C) {
      const int cg = cat ? s * C + c : c;
      const float f = chanscale ? __ldg(chanscale + (int64_t)b * Cout + cg) : 1.0f;
      const float sh = chanshift ? __ldg(chanshift + cg) : 0.0f;
      IO<TO>::st(out + ((int64_t)b * Cout + cg) * HW + r, fmaf(tile[tx][ty + i * 8], f, sh));
    }
  }
}

template <typename T, typename TI>
__global__ void __launch_bounds__(256) merge_bwd_kernel(const TI* __restrict__ dout, int nsrc, int cat, const float* __restrict__ chanscale,
                                                        MergeOutPtrs dst, int HW, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int Cout = cat ? nsrc * C : C;
  const int r0 = blockIdx.y * 32, c0g = blockIdx.x * 32;
  const int s = cat ? c0g / C : 0, c0 = cat ? c0g - s * C : c0g;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, r = r0 + tx;
    float v = 0.f;
    if (r < HW && c < C) {
      const int cg = cat ? s * C + c : c;
      const float f = chanscale ? __ldg(chanscale + (int64_t)b * Cout + cg) : 1.0f;
      v = IO<TI>::ld(dout + ((int64_t)b * Cout + cg) * HW + r) * f;
    }
    tile[ty + i * 8][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    if (r < HW && c < C) {
      const float v = tile[tx][ty + i * 8];
      if (cat) IO<T>::st(reinterpret_cast<T*>(dst.p[s]) + ((int64_t)b * HW + r) * C + c, v);
      else for (int k = 0; k < nsrc; ++k) IO<T>::st(reinterpret_cast<T*>(dst.p[k]) + ((int64_t)b * HW + r) * C + c, v);
    }
  }
}

// channels-last output: out[b, r, cg] = scale[b, cg] * src[s][b, r, c]  (8-wide vectors, coalesced on both sides)
template <typename T, typename TO>
__global__ void merge_nhwc_fwd_kernel(MergePtrs src, int nsrc, int cat, const float* __restrict__ chanscale,
                                      const float* __restrict__ chanshift, TO* __restrict__ out, int64_t total, int HW, int C) {
  const int Cout = cat ? nsrc * C : C, C8 = Cout / 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int cg = (int)(i % C8) * 8;
    const int64_t pix = i / C8;                       // b*HW + r
    const int64_t b = pix / HW;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cat) {
      const int s = cg / C, c = cg - s * C;
      IO<T>::load8(reinterpret_cast<const T*>(src.p[s]) + pix * C + c, v);
    } else {
      for (int k = 0; k < nsrc; ++k) {
        float t[8]; IO<T>::load8(reinterpret_cast<const T*>(src.p[k]) + pix * C + cg, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
      }
    }
    if (chanscale) {
      float f[8]; IO<float>::load8(chanscale + b * Cout + cg, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= f[j];
    }
    if (chanshift) {                                  // FiLM shift beta[c] (the scale gamma[c] is folded into chanscale)
      float sh[8]; IO<float>::load8(chanshift + cg, sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += sh[j];
    }
    IO<TO>::store8(out + pix * Cout + cg, v);
  }
}

template <typename T, typename TI>
__global__ void merge_nhwc_bwd_kernel(const TI* __restrict__ dout, int nsrc, int cat, const float* __restrict__ chanscale, MergeOutPtrs dst,
                                      int64_t total, int HW, int C) {
  const int Cout = cat ? nsrc * C : C, C8 = Cout / 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int cg = (int)(i % C8) * 8;
    const int64_t pix = i / C8;
    const int64_t b = pix / HW;
    float v[8];
    IO<TI>::load8(dout + pix * Cout + cg, v);
    if (chanscale) {
      float f[8]; IO<float>::load8(chanscale + b * Cout + cg, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= f[j];
    }
    if (cat) {
      const int s = cg / C, c = cg - s * C;
      IO<T>::store8(reinterpret_cast<T*>(dst.p[s]) + pix * C + c, v);
    } else {
      for (int k = 0; k < nsrc; ++k) IO<T>::store8(reinterpret_cast<T*>(dst.p[k]) + pix * C + cg, v);
    }
  }
}

extern "C" int mtus_fpn_merge_film_fwd(const void* const* srcs, int nsrc, int policy_cat, const float* chanscale,
                                       const float* chanshift, void* out, int B, int HW, int C, int dtype, int out_f32, int out_nhwc, void* stream) {
  MTUS_CHECK_ARG(srcs && out && nsrc >= 1 && nsrc <= 4 && C % 32 == 0 && B <= 65535);
  if (B == 0) return MTUS_OK;
  MergePtrs mp{};
  for (int i = 0; i < nsrc; ++i) { MTUS_CHECK_ARG(srcs[i]); mp.p[i] = srcs[i]; }
  dim3 grid((policy_cat ? nsrc * C : C) / 32, ceil_div(HW, 32), B);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_nhwc) {
    const int64_t total = (int64_t)B * HW * ((policy_cat ? nsrc * C : C) / 8);
    const int g1 = grid_for(total, 256);
    if (dtype == MTUS_F32) merge_nhwc_fwd_kernel<float, float><<<g1, 256, 0, st>>>(mp, nsrc, policy_cat, chanscale, chanshift, (float*)out, total, HW, C);
    else if (dtype == MTUS_BF16 && out_f32) merge_nhwc_fwd_kernel<bf16, float><<<g1, 256, 0, st>>>(mp, nsrc, policy_cat, chanscale, chanshift, (float*)out, total, HW, C);
    else if (dtype == MTUS_BF16) merge_nhwc_fwd_kernel<bf16, bf16><<<g1, 256, 0, st>>>(mp, nsrc, policy_cat, chanscale, chanshift, (bf16*)out, total, HW, C);
    else return MTUS_ERR_UNSUPPORTED;
    MTUS_LAUNCH_STATUS();
    return MTUS_OK;
  }
  if (dtype == MTUS_F32) merge_fwd_kernel<float, float><<<grid, 256, 0, st>>>(mp, nsrc, policy_cat, chanscale, chanshift, (float*)out, HW, C);
  else if (dtype == MTUS_BF16) {
    if (out_f32) merge_fwd_kernel<bf16, float><<<grid, 256, 0, st>>>(mp, nsrc, policy_cat, chanscale, chanshift, (float*)out, HW, C);
    else merge_fwd_kernel<bf16, bf16><<<grid, 256, 0, st>>>(mp, nsrc, policy_cat, chanscale, chanshift, (bf16*)out, HW, C);
  } else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_fpn_merge_fwd(const void* const* srcs, int nsrc, int policy_cat, const float* chanscale, void* out,
                                  int B, int HW, int C, int dtype, int out_f32, int out_nhwc, void* stream) {
  return mtus_fpn_merge_film_fwd(srcs, nsrc, policy_cat, chanscale, nullptr, out, B, HW, C, dtype, out_f32, out_nhwc, stream);
}

// FiLM parameter gradients for out = gamma[c] * (drop[b,c] * merged[b,r,c]) + beta[c] (code/models/film_layer.py:94-99 applied
// to the decoder output, multitask_model.py:214-216), one pass over dout and the merge sources (both NHWC):
//   dbeta[c] += sum_{b,r} dout[b,r,c]        dgamma[c] += sum_{b,r} dout[b,r,c] * drop[b,c] * merged[b,r,c]
// Block = (Cout/8 vector lanes) x (256 / (Cout/8) pixel lanes); per-thread accumulators, shared-memory fold, one atomic per
// channel per block.
template <typename T, typename TI>
__global__ void __launch_bounds__(256) film_grad_kernel(const TI* __restrict__ dout, MergePtrs src, int nsrc, int cat,
                                                        const float* __restrict__ dropscale, float* __restrict__ dgamma,
                                                        float* __restrict__ dbeta, int64_t npix, int HW, int C, int pix_per_block) {
  extern __shared__ float sred[];   // [2][Cout]
  const int Cout = cat ? nsrc * C : C, C8 = Cout / 8;
  const int v = threadIdx.x % C8, pl = threadIdx.x / C8, npl = 256 / C8;
  const int cg = v * 8;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block, p1 = min(npix, p0 + pix_per_block);
  float ag[8], ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) ag[k] = ab[k] = 0.f;
  if (pl < npl) {
    for (int64_t pix = p0 + pl; pix < p1; pix += npl) {
      float d[8], x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      IO<TI>::load8(dout + pix * Cout + cg, d);
      if (cat) {
        const int s_ = cg / C, c = cg - s_ * C;
        IO<T>::load8(reinterpret_cast<const T*>(src.p[s_]) + pix * C + c, x);
      } else {
        for (int k = 0; k < nsrc; ++k) {
          float t[8]; IO<T>::load8(reinterpret_cast<const T*>(src.p[k]) + pix * C + cg, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] += t[j];
        }
      }
      if (dropscale) {
        float f[8]; IO<float>::load8(dropscale + (pix / HW) * Cout + cg, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] *= f[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { ag[j] = fmaf(d[j], x[j], ag[j]); ab[j] += d[j]; }
    }
  }
  for (int c = threadIdx.x; c < 2 * Cout; c += 256) sred[c] = 0.f;
  __syncthreads();
  if (pl < npl) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(&sred[cg + j], ag[j]); atomicAdd(&sred[Cout + cg + j], ab[j]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cout; c += 256) { atomicAdd(dgamma + c, sred[c]); atomicAdd(dbeta + c, sred[Cout + c]); }
}

extern "C" int mtus_film_grad(const void* dout, int dout_f32, const void* const* srcs, int nsrc, int policy_cat,
                              const float* dropscale, float* dgamma, float* dbeta, int B, int HW, int C, int dtype, void* stream) {
  MTUS_CHECK_ARG(dout && srcs && dgamma && dbeta && nsrc >= 1 && nsrc <= 4 && C % 8 == 0);
  const int Cout = policy_cat ? nsrc * C : C;
  MTUS_CHECK_ARG(Cout / 8 <= 256);
  if (B == 0) return MTUS_OK;
  MergePtrs mp{};
  for (int i = 0; i < nsrc; ++i) { MTUS_CHECK_ARG(srcs[i]); mp.p[i] = srcs[i]; }
  const int64_t npix = (int64_t)B * HW;
  const int npl = 256 / (Cout / 8);
  int64_t blocks = 148 * 4;
  const int64_t maxb = (npix + npl * 4 - 1) / (npl * 4);
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  const int ppb = (int)((npix + blocks - 1) / blocks);
  const int grid = (int)((npix + ppb - 1) / ppb);
  const size_t sm = sizeof(float) * 2 * Cout;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) film_grad_kernel<float, float><<<grid, 256, sm, st>>>((const float*)dout, mp, nsrc, policy_cat, dropscale, dgamma, dbeta, npix, HW, C, ppb);
  else if (dtype == MTUS_BF16 && dout_f32) film_grad_kernel<bf16, float><<<grid, 256, sm, st>>>((const float*)dout, mp, nsrc, policy_cat, dropscale, dgamma, dbeta, npix, HW, C, ppb);
  else if (dtype == MTUS_BF16) film_grad_kernel<bf16, bf16><<<grid, 256, sm, st>>>((const bf16*)dout, mp, nsrc, policy_cat, dropscale, dgamma, dbeta, npix, HW, C, ppb);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_fpn_merge_bwd(const void* dout, int nsrc, int policy_cat, const float* chanscale, void* const* dsrcs,
                                  int B, int HW, int C, int dtype, int in_f32, int in_nhwc, void* stream) {
  MTUS_CHECK_ARG(dout && dsrcs && nsrc >= 1 && nsrc <= 4 && C % 32 == 0 && B <= 65535);
  if (B == 0) return MTUS_OK;
  MergeOutPtrs mp{};
  for (int i = 0; i < nsrc; ++i) { MTUS_CHECK_ARG(dsrcs[i]); mp.p[i] = dsrcs[i]; }
  dim3 grid((policy_cat ? nsrc * C : C) / 32, ceil_div(HW, 32), B);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_nhwc) {
    const int64_t total = (int64_t)B * HW * ((policy_cat ? nsrc * C : C) / 8);
    const int g1 = grid_for(total, 256);
    if (dtype == MTUS_F32) merge_nhwc_bwd_kernel<float, float><<<g1, 256, 0, st>>>((const float*)dout, nsrc, policy_cat, chanscale, mp, total, HW, C);
    else if (dtype == MTUS_BF16 && in_f32) merge_nhwc_bwd_kernel<bf16, float><<<g1, 256, 0, st>>>((const float*)dout, nsrc, policy_cat, chanscale, mp, total, HW, C);
    else if (dtype == MTUS_BF16) merge_nhwc_bwd_kernel<bf16, bf16><<<g1, 256, 0, st>>>((const bf16*)dout, nsrc, policy_cat, chanscale, mp, total, HW, C);
    else return MTUS_ERR_UNSUPPORTED;
    MTUS_LAUNCH_STATUS();
    return MTUS_OK;
  }
  if (dtype == MTUS_F32) merge_bwd_kernel<float, float><<<grid, 256, 0, st>>>((const float*)dout, nsrc, policy_cat, chanscale, mp, HW, C);
  else if (dtype == MTUS_BF16) {
    if (in_f32) merge_bwd_kernel<bf16, float><<<grid, 256, 0, st>>>((const float*)dout, nsrc, policy_cat, chanscale, mp, HW, C);
    else merge_bwd_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)dout, nsrc, policy_cat, chanscale, mp, HW, C);
  } else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// ---- 3x3 weight repack ---------------------------------------------------------------------------
template <typename T>
__global__ void conv_repack_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int Cout, int Cin) {
  const int total = Cout * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9; const int c = (i / 9) % Cin; const int co = i / (9 * Cin);
    const float v = w[i];
    if (wf) IO<T>::st(wf + ((int64_t)co * 9 + tap) * Cin + c, v);
    if (wd) IO<T>::st(wd + ((int64_t)c * 9 + (8 - tap)) * Cout + co, v);
  }
}

extern "C" int mtus_conv3x3_repack(const float* w, void* w_fwd, void* w_dgrad, int Cout, int Cin, int dtype, void* stream) {
  MTUS_CHECK_ARG(w && (w_fwd || w_dgrad) && Cout > 0 && Cin > 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for((int64_t)Cout * Cin * 9, 256);
  if (dtype == MTUS_F32) conv_repack_kernel<float><<<g, 256, 0, st>>>(w, (float*)w_fwd, (float*)w_dgrad, Cout, Cin);
  else if (dtype == MTUS_BF16) conv_repack_kernel<bf16><<<g, 256, 0, st>>>(w, (bf16*)w_fwd, (bf16*)w_dgrad, Cout, Cin);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

__global__ void conv_unpack_grad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int Cin) {
  const int total = Cout * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9; const int c = (i / 9) % Cin; const int co = i / (9 * Cin);
    dw[i] += dwp[((int64_t)co * 9 + tap) * Cin + c];
  }
}

extern "C" int mtus_conv3x3_unpack_grad(const float* dw_packed, float* dw, int Cout, int Cin, void* stream) {
  MTUS_CHECK_ARG(dw_packed && dw && Cout > 0 && Cin > 0);
  conv_unpack_grad_kernel<<<grid_for((int64_t)Cout * Cin * 9, 256), 256, 0, (cudaStream_t)stream>>>(dw_packed, dw, Cout, Cin);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}
