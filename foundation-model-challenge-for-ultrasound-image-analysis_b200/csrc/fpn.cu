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
#include <stdlib.h>

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
// sigma(z) = 1 / (1 + 2^(-z log2 e)) with ex2.approx + rcp.approx (2 MUFU + 2 FMA; relative error ~2e-7).  The IEEE division of
// the first version cost ~10 extra instructions per element and made the SiLU kernels issue-bound at 8 warps per SM.
__device__ __forceinline__ float gn_sigmoid(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
// z = (x - mean) * rstd * gamma + beta evaluated as ONE fma x * a + s with a = rstd * gamma, s = beta - mean * a.  Every kernel
// that needs z (forward, fused forward, the backward's ReLU-gate recomputation) uses this form, so the gate is bit-identical.
__device__ __forceinline__ void gn_affine(float mean, float rstd, float gamma, float beta, float& a, float& s) {
  a = rstd * gamma;
  s = fmaf(-mean, a, beta);
}

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
      float ga, gs;
      gn_affine(m, r, gm[k], bt[k], ga, gs);
      const float z = fmaf(a[k], ga, gs);
      a[k] = ACT == 0 ? fmaxf(z, 0.f) : z * gn_sigmoid(z);
    }
    IO<T>::store8(y + i * 8, a);
  }
}

// ---- fused GroupNorm + activation forward: ONE kernel, one HBM read + one HBM write (the algorithmic minimum) ------------------
// A thread-block CLUSTER of CL CTAs owns one sample: each CTA pulls its slice of the sample's pixels into shared memory once
// (cp.async, every load of the CTA in flight together), the per-group sums are exchanged through DISTRIBUTED SHARED MEMORY
// (cluster.map_shared_rank), and mean -> centred variance -> normalise + activation all run out of shared memory.  Replaces
// memset x2 + gn_stats x2 + gn_finalize + gn_relu_fwd (6 launches, 3 reads + 1 write of the tensor).  mean / rstd are still
// written for the backward.  Used when a sample's slice fits (bf16 [56,56,128] = 98 KB per CTA at CL = 8); everything else
// (fp32 at the largest map, BatchNorm's single "sample" of all rows) stays on the two-pass kernels above.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <typename T, int ACT>
__global__ void __launch_bounds__(256) gn_fused_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int HW, int C,
                                                           int G, int ppc, float eps) {
  extern __shared__ __align__(16) uint8_t gf_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int b = blockIdx.y, C8 = C / 8, cpg = C / G, tid = threadIdx.x;
  const int p0 = min(HW, rank * ppc), p1 = min(HW, p0 + ppc), npix = p1 - p0, nvec = npix * C8;
  T* sdata = reinterpret_cast<T*>(gf_smem);                                   // [ppc][C]
  float* sred = reinterpret_cast<float*>(gf_smem + (size_t)ppc * C * sizeof(T));   // [C] per-channel partial sums
  float* sgA = sred + C;        // [G] this CTA's group sums (read by the peers)
  float* sgB = sgA + G;         // [G] this CTA's centred group sums of squares
  float* smu = sgB + G;         // [G]
  float* srs = smu + G;         // [G]
  const T* src = x + ((int64_t)b * HW + p0) * C;
  constexpr int PIECES = sizeof(T) * 8 / 16;                                  // 16-byte pieces per 8-element vector
  for (int i = tid; i < nvec * PIECES; i += 256) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<uint8_t*>(sdata) + (size_t)i * 16);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(reinterpret_cast<const uint8_t*>(src) + (size_t)i * 16) : "memory");
  }
  for (int c = tid; c < C; c += 256) sred[c] = 0.f;
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // this thread's vectors: i = tid + 256 j -> channel vector v = tid % C8 for every j (256 % C8 == 0)
  const int v = tid % C8;
  auto ld8 = [&](int i, float (&a)[8]) {
    if (sizeof(T) == 2) {
      const uint4 r = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(sdata) + (size_t)i * 16);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); a[2 * k] = f.x; a[2 * k + 1] = f.y; }
    } else {
      const float4 lo = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(sdata) + (size_t)i * 32);
      const float4 hi = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(sdata) + (size_t)i * 32 + 16);
      a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
    }
  };
  // block reduction of 8 per-thread channel sums: lanes that hold the same channel vector first (C8 < 32), then shared atomics
  auto block_reduce = [&](float (&acc)[8]) {
    for (int off = C8; off < 32; off <<= 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
    }
    if ((tid & 31) < C8 || C8 >= 32) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&sred[v * 8 + k], acc[k]);
    }
  };
  // cluster-wide group totals of the per-channel sums in sred, through this CTA's slot sg (read remotely by the peers)
  auto cluster_total = [&](float* sg, float* dst, bool is_var) {
    __syncthreads();
    for (int g = tid; g < G; g += 256) {
      float t = 0.f;
      for (int k = 0; k < cpg; ++k) t += sred[g * cpg + k];
      sg[g] = t;
    }
    cluster.sync();
    const float inv_n = 1.0f / ((float)HW * (float)cpg);
    for (int g = tid; g < G; g += 256) {
      float t = 0.f;
      for (int r = 0; r < CL; ++r) t += cluster.map_shared_rank(sg, r)[g];
      dst[g] = is_var ? rsqrtf(t * inv_n + eps) : t * inv_n;
    }
    __syncthreads();
  };
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int i = tid; i < nvec; i += 256) {
    float a[8];
    ld8(i, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += a[k];
  }
  block_reduce(acc);
  cluster_total(sgA, smu, false);
  float mu[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) mu[k] = smu[(v * 8 + k) / cpg];
  for (int c = tid; c < C; c += 256) sred[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int i = tid; i < nvec; i += 256) {
    float a[8];
    ld8(i, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = a[k] - mu[k]; acc[k] = fmaf(d, d, acc[k]); }
  }
  block_reduce(acc);
  cluster_total(sgB, srs, true);
  cluster.barrier_arrive();                             // this CTA has read its peers' slots; it may not exit before they have read its own
  if (rank == 0) {
    for (int g = tid; g < G; g += 256) { mean_out[b * G + g] = smu[g]; rstd_out[b * G + g] = srs[g]; }
  }
  float ga[8], gs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = v * 8 + k;
    gn_affine(mu[k], srs[c / cpg], __ldg(gamma + c), __ldg(beta + c), ga[k], gs[k]);
  }
  T* dstp = y + ((int64_t)b * HW + p0) * C;
  for (int i = tid; i < nvec; i += 256) {
    float a[8];
    ld8(i, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float z = fmaf(a[k], ga[k], gs[k]);
      a[k] = ACT == 0 ? fmaxf(z, 0.f) : z * gn_sigmoid(z);
    }
    IO<T>::store8(dstp + (size_t)i * 8, a);
  }
  cluster.barrier_wait();
}

// cluster size and pixels per CTA of the fused kernel; 0 = not eligible
static int gn_fused_plan(int B, int HW, int C, int G, int esize, int& ppc, size_t& smem) {
  const int C8 = C / 8;
  if (C % 8 || C8 > 256 || 256 % C8 || G <= 0 || C % G || B > 65535 || HW <= 0) return 0;
  int CL = 8;
  while (CL > 1 && (int64_t)B * CL > 2 * 148) CL >>= 1;                     // about one wave of CTAs at two per SM
  while (CL < 8 && (int64_t)((HW + CL - 1) / CL) * C * esize > 100 * 1024) CL <<= 1;
  ppc = (HW + CL - 1) / CL;
  smem = (size_t)ppc * C * esize + sizeof(float) * (C + 4 * G) + 16;
  if (smem > 200 * 1024) return 0;
  return CL;
}

template <typename T, int ACT>
static int gn_fused_launch(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int B, int HW, int C, int G,
                           float eps, int CL, int ppc, size_t smem, cudaStream_t st) {
  auto kern = gn_fused_fwd_kernel<T, ACT>;
  static mtus_per_device_flag configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured.set();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, B); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (const T*)x, gamma, beta, (T*)y, mean, rstd, HW, C, G, ppc, eps);
  if (e != cudaSuccess) return (int)e;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
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

// statistics + normalise + activation in one call: the fused cluster kernel when a sample's slice fits in shared memory,
// else mtus_groupnorm_stats followed by mtus_groupnorm_act_fwd.  mean / rstd [B, G] are outputs (saved for the backward).
extern "C" int mtus_groupnorm_act_fused_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int B,
                                            int HW, int C, int G, float eps, int act, int dtype, void* stream) {
  MTUS_CHECK_ARG(x && gamma && beta && y && mean && rstd && (act == 0 || act == 1));
  MTUS_CHECK_ARG(dtype == MTUS_F32 || dtype == MTUS_BF16);
  if (B == 0 || HW == 0) return MTUS_OK;
  static int use_fused = -1;
  if (use_fused < 0) { const char* e = getenv("MTUS_GN_FUSED"); use_fused = e ? atoi(e) : 1; }
  int ppc = 0; size_t smem = 0;
  const int CL = use_fused ? gn_fused_plan(B, HW, C, G, dtype == MTUS_F32 ? 4 : 2, ppc, smem) : 0;
  if (CL > 0) {
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MTUS_F32) return act == 0 ? gn_fused_launch<float, 0>(x, gamma, beta, y, mean, rstd, B, HW, C, G, eps, CL, ppc, smem, st)
                                           : gn_fused_launch<float, 1>(x, gamma, beta, y, mean, rstd, B, HW, C, G, eps, CL, ppc, smem, st);
    return act == 0 ? gn_fused_launch<bf16, 0>(x, gamma, beta, y, mean, rstd, B, HW, C, G, eps, CL, ppc, smem, st)
                    : gn_fused_launch<bf16, 1>(x, gamma, beta, y, mean, rstd, B, HW, C, G, eps, CL, ppc, smem, st);
  }
  const int rc = mtus_groupnorm_stats(x, mean, rstd, B, HW, C, G, eps, dtype, stream);
  if (rc) return rc;
  return mtus_groupnorm_act_fwd(x, mean, rstd, gamma, beta, y, B, HW, C, G, act, dtype, stream);
}

// d act(z) / dz times dy.  ReLU: gate from the saved output y (USEY) or from z recomputed as x * a + s (the forward's own
// expression, gn_affine: one read of the tensor less); SiLU: z recomputed (sigma(z) (1 + z (1 - sigma(z)))).
template <int ACT, bool USEY>
__device__ __forceinline__ float gn_act_grad(float dy, float yv, float xv, float ga, float gs) {
  if (ACT == 0 && USEY) return yv > 0.f ? dy : 0.f;
  const float z = fmaf(xv, ga, gs);
  if (ACT == 0) return z > 0.f ? dy : 0.f;
  const float sg = gn_sigmoid(z);
  return dy * sg * (1.0f + z * (1.0f - sg));
}

// backward pass 1: per (b,g) s1 = sum dyr*gamma, s2 = sum dyr*gamma*xhat (ws[0..BG), ws[BG..2BG));
//                  per channel dgamma += sum dyr*xhat, dbeta += sum dyr     (dyr = dy * (y > 0))
template <typename T, int ACT, bool USEY>
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
  float mu[8], rs[8], gm[8], ga[8], gs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int g = (v * 8 + k) / cpg;
    mu[k] = mean[b * G + g]; rs[k] = rstd[b * G + g]; gm[k] = gamma[v * 8 + k];
    gn_affine(mu[k], rs[k], gm[k], (ACT == 0 && USEY) ? 0.f : beta[v * 8 + k], ga[k], gs[k]);
  }
  float a1[8], a2[8], ag[8], ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a1[k] = a2[k] = ag[k] = ab[k] = 0.f;
  if (pl < npl) {
    for (int p = p0 + pl; p < p1; p += npl) {
      const int64_t off = ((int64_t)b * HW + p) * C + v * 8;
      float d[8], xv[8], yv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      IO<T>::load8(dy + off, d); IO<T>::load8(x + off, xv);
      if (ACT == 0 && USEY) IO<T>::load8(y + off, yv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (xv[k] - mu[k]) * rs[k];
        const float dr = gn_act_grad<ACT, USEY>(d[k], yv[k], xv[k], ga[k], gs[k]);
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
    MTUS_ATOMIC_ADD(ws + b * G + g, s1);
    MTUS_ATOMIC_ADD(ws + B * G + b * G + g, s2);
  }
  for (int c = threadIdx.x; c < C; c += 256) { MTUS_ATOMIC_ADD(dgamma + c, sred[2 * C + c]); MTUS_ATOMIC_ADD(dbeta + c, sred[3 * C + c]); }
}

template <typename T, int ACT, bool USEY>
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
    if (ACT == 0 && USEY) IO<T>::load8(y + i * 8, yv); else IO<float>::load8(beta + v * 8, bt);
    IO<float>::load8(gamma + v * 8, gm);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int g = (v * 8 + k) / cpg;
      const float m = __ldg(mean + b * G + g), r = __ldg(rstd + b * G + g);
      const float s1 = __ldg(ws + b * G + g) * inv_n, s2 = __ldg(ws + (int64_t)B * G + b * G + g) * inv_n;
      const float xh = (xv[k] - m) * r;
      float ga, gs;
      gn_affine(m, r, gm[k], bt[k], ga, gs);
      const float dr = gn_act_grad<ACT, USEY>(d[k], yv[k], xv[k], ga, gs);
      o[k] = r * (dr * gm[k] - s1 - xh * s2);
    }
    IO<T>::store8(dx + i * 8, o);
  }
}

// ---- fused GroupNorm + activation backward: ONE cluster kernel, two reads + one write (the algorithmic minimum) ---------------
// Same ownership as gn_fused_fwd_kernel: a cluster of CL CTAs owns one sample and each CTA pulls its slice of x AND dy into
// shared memory once.  Pass 1 accumulates per channel  ag = sum dr xhat,  ab = sum dr  (dr = dy act'(z), gate / sigmoid
// recomputed from x with the forward's own fma); dgamma / dbeta are those sums, and the two group means the input gradient
// needs are  s1 = sum_c gamma_c ab_c,  s2 = sum_c gamma_c ag_c  over the group's channels, exchanged across the cluster
// through distributed shared memory.  Pass 2 writes dx = rstd (dr gamma - s1/n - xhat s2/n) out of shared memory.  Replaces
// memset + gn_bwd_reduce + gn_bwd_apply (2 x 2 reads + 1 write, and a reduce kernel that kept only 16 KB per SM in flight:
// 80 us at bf16 [32,56,56,128] against 23 us of traffic).
template <typename T, int ACT, bool DY_SMEM>
__global__ void __launch_bounds__(256) gn_fused_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, T* __restrict__ dx, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int HW, int C, int G, int ppc, int vec_ok) {
  extern __shared__ __align__(16) uint8_t gf_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int b = blockIdx.y, C8 = C / 8, cpg = C / G, tid = threadIdx.x;
  const int p0 = min(HW, rank * ppc), p1 = min(HW, p0 + ppc), npix = p1 - p0, nvec = npix * C8;
  const size_t slice = (size_t)ppc * C * sizeof(T);
  uint8_t* sx = gf_smem;                                                      // [ppc][C] x
  uint8_t* sdy = gf_smem + slice;                                             // [ppc][C] dy (DY_SMEM only)
  float* sred = reinterpret_cast<float*>(gf_smem + (DY_SMEM ? 2 : 1) * slice);   // [2][C]: ag, ab per channel (this CTA)
  float* sg1 = sred + 2 * C;    // [G] this CTA's partial of s1 (read by the peers)
  float* sg2 = sg1 + G;         // [G] this CTA's partial of s2
  float* st1 = sg2 + G;         // [G] s1 / n of the sample
  float* st2 = st1 + G;         // [G] s2 / n
  const size_t goff = ((size_t)b * HW + p0) * C * sizeof(T);
  constexpr int PIECES = sizeof(T) * 8 / 16;
  for (int i = tid; i < nvec * PIECES; i += 256) {
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(sx + (size_t)i * 16), d1 = (uint32_t)__cvta_generic_to_shared(sdy + (size_t)i * 16);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(reinterpret_cast<const uint8_t*>(x) + goff + (size_t)i * 16) : "memory");
    if (DY_SMEM) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(reinterpret_cast<const uint8_t*>(dy) + goff + (size_t)i * 16) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int c = tid; c < 2 * C; c += 256) sred[c] = 0.f;
  const int v = tid % C8;                                                     // this thread's channel vector (256 % C8 == 0)
  float mu[8], rs[8], gm[8], ga[8], gs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = v * 8 + k, g = c / cpg;
    mu[k] = __ldg(mean + b * G + g); rs[k] = __ldg(rstd + b * G + g); gm[k] = __ldg(gamma + c);
    gn_affine(mu[k], rs[k], gm[k], __ldg(beta + c), ga[k], gs[k]);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  auto ld8 = [&](const uint8_t* base, int i, float (&a)[8]) {
    if (sizeof(T) == 2) {
      const uint4 r = *reinterpret_cast<const uint4*>(base + (size_t)i * 16);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); a[2 * k] = f.x; a[2 * k + 1] = f.y; }
    } else {
      const float4 lo = *reinterpret_cast<const float4*>(base + (size_t)i * 32);
      const float4 hi = *reinterpret_cast<const float4*>(base + (size_t)i * 32 + 16);
      a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
    }
  };
  auto st8 = [&](uint8_t* base, int i, const float (&a)[8]) {
    if (sizeof(T) == 2) {
      uint4 r;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(a[2 * k], a[2 * k + 1]);
      *reinterpret_cast<uint4*>(base + (size_t)i * 16) = r;
    } else {
      *reinterpret_cast<float4*>(base + (size_t)i * 32) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(base + (size_t)i * 32 + 16) = make_float4(a[4], a[5], a[6], a[7]);
    }
  };
  float ag[8], ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) ag[k] = ab[k] = 0.f;
  const T* dyg = dy + ((int64_t)b * HW + p0) * C;      // this CTA's slice of dy in global memory (!DY_SMEM: read twice, 2nd time from L2)
  if (DY_SMEM) {
    for (int i = tid; i < nvec; i += 256) {
      float xv[8], d[8];
      ld8(sx, i, xv); ld8(sdy, i, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        d[k] = gn_act_grad<ACT, false>(d[k], 0.f, xv[k], ga[k], gs[k]);          // dr = dy act'(z)
        ag[k] = fmaf(d[k], (xv[k] - mu[k]) * rs[k], ag[k]);
        ab[k] += d[k];
      }
      if (ACT != 0) st8(sdy, i, d);     // SiLU: park dr over dy (own vector only) so pass 2 does not evaluate the sigmoid again
    }
  } else {
    for (int i0 = tid; i0 < nvec; i0 += 4 * 256) {      // four independent 16/32-byte global loads of dy in flight per thread
      float d[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 256;
        if (i < nvec) IO<T>::load8(dyg + (size_t)i * 8, d[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 256;
        if (i < nvec) {
          float xv[8];
          ld8(sx, i, xv);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float dr = gn_act_grad<ACT, false>(d[u][k], 0.f, xv[k], ga[k], gs[k]);
            ag[k] = fmaf(dr, (xv[k] - mu[k]) * rs[k], ag[k]);
            ab[k] += dr;
          }
        }
      }
    }
  }
  for (int off = C8; off < 32; off <<= 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { ag[k] += __shfl_xor_sync(0xffffffffu, ag[k], off); ab[k] += __shfl_xor_sync(0xffffffffu, ab[k], off); }
  }
  if ((tid & 31) < C8 || C8 >= 32) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&sred[v * 8 + k], ag[k]); atomicAdd(&sred[C + v * 8 + k], ab[k]); }
  }
  __syncthreads();
  for (int g = tid; g < G; g += 256) {
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < cpg; ++k) { const float gmc = __ldg(gamma + g * cpg + k); s1 = fmaf(gmc, sred[C + g * cpg + k], s1); s2 = fmaf(gmc, sred[g * cpg + k], s2); }
    sg1[g] = s1; sg2[g] = s2;
  }
  if (npix > 0) {
    if (vec_ok) {
      for (int c4 = tid; c4 < C / 2; c4 += 256) {                              // C / 4 vectors of dgamma, then C / 4 of dbeta
        const int which = c4 >= C / 4, c = (c4 - which * (C / 4)) * 4;
        const float* sp = sred + which * C + c;
        float* dst = (which ? dbeta : dgamma) + c;
#ifdef MTUS_DIAG_NOATOM
        if (sp[0] == 1.2345e-30f)
#endif
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(sp[0]), "f"(sp[1]), "f"(sp[2]), "f"(sp[3]) : "memory");
      }
    } else {
      for (int c = tid; c < C; c += 256) { MTUS_ATOMIC_ADD(dgamma + c, sred[c]); MTUS_ATOMIC_ADD(dbeta + c, sred[C + c]); }
    }
  }
  cluster.sync();
  const float inv_n = 1.0f / ((float)HW * (float)cpg);
  for (int g = tid; g < G; g += 256) {
    float t1 = 0.f, t2 = 0.f;
    for (int r = 0; r < CL; ++r) { t1 += cluster.map_shared_rank(sg1, r)[g]; t2 += cluster.map_shared_rank(sg2, r)[g]; }
    st1[g] = t1 * inv_n; st2[g] = t2 * inv_n;
  }
  __syncthreads();
  cluster.barrier_arrive();                             // this CTA has read its peers' slots; it may not exit before they have read its own
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { const int g = (v * 8 + k) / cpg; s1[k] = st1[g]; s2[k] = st2[g]; }
  T* dstp = dx + ((int64_t)b * HW + p0) * C;
  if (DY_SMEM) {
    for (int i = tid; i < nvec; i += 256) {
      float xv[8], d[8], o[8];
      ld8(sx, i, xv); ld8(sdy, i, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float dr = ACT != 0 ? d[k] : gn_act_grad<ACT, false>(d[k], 0.f, xv[k], ga[k], gs[k]);
        const float xh = (xv[k] - mu[k]) * rs[k];
        o[k] = rs[k] * (dr * gm[k] - s1[k] - xh * s2[k]);
      }
      IO<T>::store8(dstp + (size_t)i * 8, o);
    }
  } else {
    for (int i0 = tid; i0 < nvec; i0 += 4 * 256) {
      float d[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 256;
        if (i < nvec) IO<T>::load8(dyg + (size_t)i * 8, d[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 256;
        if (i < nvec) {
          float xv[8], o[8];
          ld8(sx, i, xv);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float dr = gn_act_grad<ACT, false>(d[u][k], 0.f, xv[k], ga[k], gs[k]);
            const float xh = (xv[k] - mu[k]) * rs[k];
            o[k] = rs[k] * (dr * gm[k] - s1[k] - xh * s2[k]);
          }
          IO<T>::store8(dstp + (size_t)i * 8, o);
        }
      }
    }
  }
  cluster.barrier_wait();
}

#define GN_FUSED_BWD_MAX_SMEM (216 * 1024)
// cluster size and pixels per CTA of the fused backward; 0 = not eligible (the two-kernel path below runs)
// dy_smem: x AND dy staged (small maps: everything fits at two or more CTAs per SM); otherwise only x is staged and dy is read
// from global memory in both passes (the second read hits the L2) -- 100 KB instead of 200 KB per CTA at bf16 [56,56,128], so
// two CTAs per SM are resident and the 32 clusters of a batch run as ONE wave (with 200 KB: two waves, SMs idle 44 % of the time).
static int gn_fused_bwd_plan(int B, int HW, int C, int G, int esize, int& ppc, size_t& smem, bool& dy_smem) {
  const int C8 = C / 8;
  if (C % 8 || C8 > 256 || 256 % C8 || G <= 0 || C % G || B > 65535 || HW <= 0) return 0;
  int CL = 8;
  while (CL > 1 && HW < 16 * CL) CL >>= 1;
  ppc = (HW + CL - 1) / CL;
  const size_t slice = (size_t)ppc * C * esize, tail = sizeof(float) * (2 * C + 4 * G) + 16;
  static int force = -1;                 // MTUS_GN_BWD_DY_SMEM=1: always stage dy too when it fits (A/B)
  if (force < 0) { const char* e = getenv("MTUS_GN_BWD_DY_SMEM"); force = e ? atoi(e) : 0; }
  dy_smem = 2 * slice + tail <= (size_t)(force ? GN_FUSED_BWD_MAX_SMEM : 100 * 1024);
  smem = (dy_smem ? 2 : 1) * slice + tail;
  if (smem > GN_FUSED_BWD_MAX_SMEM) return 0;
  return CL;
}

template <typename T, int ACT, bool DY_SMEM>
static int gn_fused_bwd_launch(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta, void* dx,
                               float* dgamma, float* dbeta, int B, int HW, int C, int G, int CL, int ppc, size_t smem, cudaStream_t st) {
  auto kern = gn_fused_bwd_kernel<T, ACT, DY_SMEM>;
  static mtus_per_device_flag configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GN_FUSED_BWD_MAX_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured.set();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, B); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const int vec_ok = (((uintptr_t)dgamma | (uintptr_t)dbeta) & 15) == 0 && C % 4 == 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (const T*)dy, (const T*)x, mean, rstd, gamma, beta, (T*)dx, dgamma, dbeta, HW, C, G, ppc, vec_ok);
  if (e != cudaSuccess) return (int)e;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_groupnorm_act_bwd(const void* dy, const void* x, const void* y, const float* mean, const float* rstd,
                                      const float* gamma, const float* beta, void* dx, float* dgamma, float* dbeta, float* ws,
                                      int B, int HW, int C, int G, int act, int dtype, void* stream) {
  MTUS_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && ws && (act == 0 || act == 1));
  MTUS_CHECK_ARG(act == 0 ? (y != nullptr || beta != nullptr) : beta != nullptr);   // ReLU: gate from y, or recomputed when beta is given
  MTUS_CHECK_ARG(C % 8 == 0 && C / 8 <= 256 && G > 0 && C % G == 0 && B <= 65535);
  if (B == 0) return MTUS_OK;
  const bool usey = act == 0 && beta == nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (!usey && (dtype == MTUS_F32 || dtype == MTUS_BF16)) {
    static int use_fused = -1;
    if (use_fused < 0) {
      const char* ev = getenv("MTUS_GN_FUSED"); use_fused = ev ? atoi(ev) : 1;
      const char* eb = getenv("MTUS_GN_FUSED_BWD"); if (eb) use_fused = atoi(eb);
    }
    int ppc = 0; size_t smem = 0; bool dys = true;
    const int CL = use_fused ? gn_fused_bwd_plan(B, HW, C, G, dtype == MTUS_F32 ? 4 : 2, ppc, smem, dys) : 0;
    if (CL > 0) {
#define GN_FB(T_, A_, D_) gn_fused_bwd_launch<T_, A_, D_>(dy, x, mean, rstd, gamma, beta, dx, dgamma, dbeta, B, HW, C, G, CL, ppc, smem, st)
      if (dtype == MTUS_F32) return act == 0 ? (dys ? GN_FB(float, 0, true) : GN_FB(float, 0, false)) : (dys ? GN_FB(float, 1, true) : GN_FB(float, 1, false));
      return act == 0 ? (dys ? GN_FB(bf16, 0, true) : GN_FB(bf16, 0, false)) : (dys ? GN_FB(bf16, 1, true) : GN_FB(bf16, 1, false));
#undef GN_FB
    }
  }
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(float) * 2 * B * G, st);
  if (e != cudaSuccess) return (int)e;
  int ppc; const int chunks = gn_chunks(B, HW, C, ppc);
  dim3 grid(chunks, B);
  const size_t sm = sizeof(float) * 4 * C;
  const int64_t total = (int64_t)B * HW * (C / 8);
#define GN_BWD(T_, ACT_, UY_)                                                                                                             \
  {                                                                                                                                       \
    gn_bwd_reduce_kernel<T_, ACT_, UY_><<<grid, 256, sm, st>>>((const T_*)dy, (const T_*)x, (const T_*)y, mean, rstd, gamma, beta, ws, dgamma, dbeta, B, HW, C, G, ppc); \
    gn_bwd_apply_kernel<T_, ACT_, UY_><<<grid_for(total, 256), 256, 0, st>>>((const T_*)dy, (const T_*)x, (const T_*)y, mean, rstd, gamma, beta, ws, (T_*)dx, total, B, HW, C / 8, G, C / G); \
  }
  if (dtype == MTUS_F32) { if (act == 1) GN_BWD(float, 1, false) else if (usey) GN_BWD(float, 0, true) else GN_BWD(float, 0, false) }
  else if (dtype == MTUS_BF16) { if (act == 1) GN_BWD(bf16, 1, false) else if (usey) GN_BWD(bf16, 0, true) else GN_BWD(bf16, 0, false) }
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
#define BN_BWD(T_, ACT_, UY_)                                                                                                             \
  {                                                                                                                                       \
    gn_bwd_reduce_kernel<T_, ACT_, UY_><<<grid, 256, sm, st>>>((const T_*)dy, (const T_*)x, (const T_*)y, mean, rstd, gamma, beta, ws, dgamma, dbeta, B, HW, C, G, ppc); \
    if (!training) { e = cudaMemsetAsync(ws, 0, sizeof(float) * 2 * B * G, st); if (e != cudaSuccess) return (int)e; }                   \
    gn_bwd_apply_kernel<T_, ACT_, UY_><<<grid_for(total, 256), 256, 0, st>>>((const T_*)dy, (const T_*)x, (const T_*)y, mean, rstd, gamma, beta, ws, (T_*)dx, total, B, HW, C / 8, G, C / G); \
  }
  if (dtype == MTUS_F32) { if (act == 0) BN_BWD(float, 0, true) else BN_BWD(float, 1, false) }
  else if (dtype == MTUS_BF16) { if (act == 0) BN_BWD(bf16, 0, true) else BN_BWD(bf16, 1, false) }
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

// grid (ceil(Wo * C8 / 256), ceil(Ho / 4), B): a thread produces the same (ox, channel vector) of FOUR consecutive output rows;
// all their taps (at most 3 input rows x 2 columns are distinct) are requested before the first is used, 32-bit index math only
#define BIL_ROWS 4
template <typename T>
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C8, float sy, float sx) {
  const int Wo = 2 * W, Ho = 2 * H, oy0 = blockIdx.y * BIL_ROWS, b = blockIdx.z;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= Wo * C8) return;
  const int c = j % C8, ox = j / C8;
  int x0, x1; float lx0, lx1;
  bil_src(ox, sx, W, x0, x1, lx0, lx1);
  const T* base = x + ((int64_t)b * H * W * C8 + c) * 8;
  float a[BIL_ROWS][8], bq[BIL_ROWS][8], cc[BIL_ROWS][8], d[BIL_ROWS][8], ly0[BIL_ROWS], ly1[BIL_ROWS];
#pragma unroll
  for (int r = 0; r < BIL_ROWS; ++r) {
    const int oy = min(oy0 + r, Ho - 1);
    int y0, y1;
    bil_src(oy, sy, H, y0, y1, ly0[r], ly1[r]);
    IO<T>::load8(base + (int64_t)(y0 * W + x0) * C8 * 8, a[r]);
    IO<T>::load8(base + (int64_t)(y0 * W + x1) * C8 * 8, bq[r]);
    IO<T>::load8(base + (int64_t)(y1 * W + x0) * C8 * 8, cc[r]);
    IO<T>::load8(base + (int64_t)(y1 * W + x1) * C8 * 8, d[r]);
  }
#pragma unroll
  for (int r = 0; r < BIL_ROWS; ++r) {
    if (oy0 + r >= Ho) break;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = ly0[r] * (lx0 * a[r][k] + lx1 * bq[r][k]) + ly1[r] * (lx0 * cc[r][k] + lx1 * d[r][k]);
    IO<T>::store8(y + ((((int64_t)b * Ho + oy0 + r) * Wo + ox) * C8 + c) * 8, o);
  }
}

// Weights with which input index i receives from the (at most 8) output indices o_lo .. o_lo + 7 along one axis: output o
// touches i through its tap y0 (weight l0) and / or y1 (weight l1).  r = s * o lies in (i - 1, i + 1) for every contributor,
// so o_lo = floor((i - 1) / s) - 1 (clamped) starts early enough and 2 / s <= 6 candidates plus the rounding margin fit in 8.
__device__ __forceinline__ int bil_gather_weights(int i, float s, int in_size, int out_size, float (&w)[8]) {
  int o_lo = (s > 0.f) ? (int)floorf((float)(i - 1) / s) - 1 : 0;
  o_lo = max(o_lo, 0);
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int o = o_lo + t;
    int i0, i1; float l0, l1;
    bil_src(o, s, in_size, i0, i1, l0, l1);
    float wt = 0.f;
    if (i0 == i) wt += l0;
    if (i1 == i) wt += l1;
    w[t] = (o < out_size) ? wt : 0.f;
  }
  return o_lo;
}

// gather-form backward (atomic-free): each input pixel collects from the output pixels whose taps touch it.
// grid (ceil(W * C8 / 256), H, B)
template <typename T>
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int H, int W, int C8, float sy, float sx) {
  const int Wo = 2 * W, Ho = 2 * H, iy = blockIdx.y, b = blockIdx.z;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= W * C8) return;
  const int c = j % C8, ix = j / C8;
  float wy[8], wx[8];
  const int oy_lo = bil_gather_weights(iy, sy, H, Ho, wy);
  const int ox_lo = bil_gather_weights(ix, sx, W, Wo, wx);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const T* base = dy + ((int64_t)b * Ho * Wo * C8 + c) * 8;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    if (wy[t] == 0.f) continue;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float wgt = wy[t] * wx[u];
      if (wgt == 0.f) continue;
      float v[8];
      IO<T>::load8(base + (int64_t)((oy_lo + t) * Wo + ox_lo + u) * C8 * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, v[k], acc[k]);
    }
  }
  IO<T>::store8(dx + ((((int64_t)b * H + iy) * W + ix) * C8 + c) * 8, acc);
}

static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

extern "C" int mtus_bilinear2x_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
  MTUS_CHECK_ARG(x && y && H > 0 && W > 0 && C % 8 == 0);
  const int64_t n = (int64_t)B * 4 * H * W * (C / 8);
  if (n == 0) return MTUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float sy = ac_scale(H, 2 * H), sx = ac_scale(W, 2 * W);
  MTUS_CHECK_ARG(2 * H <= 65535 && B <= 65535);
  const dim3 grid(ceil_div(2 * W * (C / 8), 256), ceil_div(2 * H, BIL_ROWS), B);
  if (dtype == MTUS_F32) bilinear_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, H, W, C / 8, sy, sx);
  else if (dtype == MTUS_BF16) bilinear_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)y, H, W, C / 8, sy, sx);
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
  MTUS_CHECK_ARG(H <= 65535 && B <= 65535);
  const dim3 grid(ceil_div(W * (C / 8), 256), H, B);
  if (dtype == MTUS_F32) bilinear_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, (float*)dx, H, W, C / 8, sy, sx);
  else if (dtype == MTUS_BF16) bilinear_bwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dy, (bf16*)dx, H, W, C / 8, sy, sx);
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

// ---- pointwise (1x1) convolution with a handful of output channels: the tails of the task heads -----------------------------
// SegmentationHead ends in Conv2d(128 -> num_classes, 1) (code/models/heads.py:16-42 via smp's SegmentationHead) and the baseline
// detection head in Conv2d(128 -> 4 + num_classes, 1) (heads.py:404-428): N <= 8 outputs over [B*H*W, K] channels-last rows.  As
// GEMMs they are degenerate (cuBLAS: 27 us forward, and 19 + 92 + 8 us of data / weight gradient GEMMs plus a 49 us two-block bias
// reduction backward at [100352, 128]); as streaming kernels they are one read of x forward and one read + one write backward.
// Layouts: x [B, HW, K] (T), w [N, K] fp32, y / dy [B, N, HW] fp32 (NCHW planes, what the upsampling / loss behind them reads).
// LPR = K / 8 lanes share a pixel (8 channels each), so a warp covers 32 / LPR pixels per iteration; K / 8 must be a power of two <= 32.
#define PW_MAXN 8
template <typename T, int N>
__global__ void __launch_bounds__(256) pointwise_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ y, int64_t rows, int HW, int K) {
  const int LPR = K / 8, sub = threadIdx.x % LPR, ppb = 256 / LPR;
  float wr[N][8];
#pragma unroll
  for (int j = 0; j < N; ++j) IO<float>::load8(w + (size_t)j * K + sub * 8, wr[j]);
  for (int64_t p0 = (int64_t)blockIdx.x * ppb * 2; p0 < rows; p0 += (int64_t)gridDim.x * ppb * 2) {
    float xv[2][8];
    int64_t p[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {                       // two pixels in flight per lane group
      p[u] = p0 + u * ppb + threadIdx.x / LPR;
#pragma unroll
      for (int k = 0; k < 8; ++k) xv[u][k] = 0.f;
      if (p[u] < rows) IO<T>::load8(x + p[u] * K + sub * 8, xv[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {                       // no early exit: the shuffles below need every lane of the warp
      float acc[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        acc[j] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[j] = fmaf(xv[u][k], wr[j][k], acc[j]);
      }
      for (int off = 1; off < LPR; off <<= 1) {
#pragma unroll
        for (int j = 0; j < N; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off, 32);
      }
      if (sub == 0 && p[u] < rows) {
        const int64_t b = p[u] / HW, q = p[u] - b * HW;
#pragma unroll
        for (int j = 0; j < N; ++j) y[(b * N + j) * HW + q] = acc[j] + (bias ? __ldg(bias + j) : 0.f);
      }
    }
  }
}

// dx[p, :] = sum_j dy[j, p] w[j, :];  dw[j, :] += sum_p dy[j, p] x[p, :];  dbias[j] += sum_p dy[j, p]
template <typename T, int N>
__global__ void __launch_bounds__(256) pointwise_bwd_kernel(const float* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ w,
                                                            T* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dbias, int64_t rows,
                                                            int HW, int K) {
  extern __shared__ float pw_red[];                     // [N][K] + [N]
  const int LPR = K / 8, sub = threadIdx.x % LPR, ppb = 256 / LPR;
  float wr[N][8], aw[N][8], ab[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    IO<float>::load8(w + (size_t)j * K + sub * 8, wr[j]);
    ab[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) aw[j][k] = 0.f;
  }
  for (int i = threadIdx.x; i < N * K + N; i += 256) pw_red[i] = 0.f;
  for (int64_t p0 = (int64_t)blockIdx.x * ppb * 2; p0 < rows; p0 += (int64_t)gridDim.x * ppb * 2) {
    float xv[2][8], g[2][N];
    int64_t p[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      p[u] = p0 + u * ppb + threadIdx.x / LPR;
      if (p[u] < rows) {
        IO<T>::load8(x + p[u] * K + sub * 8, xv[u]);
        const int64_t b = p[u] / HW, q = p[u] - b * HW;
#pragma unroll
        for (int j = 0; j < N; ++j) g[u][j] = __ldg(dy + (b * N + j) * HW + q);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (p[u] >= rows) continue;
      float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { o[k] = fmaf(g[u][j], wr[j][k], o[k]); aw[j][k] = fmaf(g[u][j], xv[u][k], aw[j][k]); }
        if (sub == 0) ab[j] += g[u][j];
      }
      if (dx) IO<T>::store8(dx + p[u] * K + sub * 8, o);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < N; ++j) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = aw[j][k];
      for (int off = LPR; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off, 32);      // lanes of the warp with the same channels
      if ((threadIdx.x & 31) < LPR || LPR >= 32) atomicAdd(&pw_red[j * K + sub * 8 + k], v);
    }
    if (sub == 0) atomicAdd(&pw_red[N * K + j], ab[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N * K; i += 256) MTUS_ATOMIC_ADD(dw + i, pw_red[i]);
  if (threadIdx.x < N) MTUS_ATOMIC_ADD(dbias + threadIdx.x, pw_red[N * K + threadIdx.x]);
}

static bool pw_ok(int K, int N) {
  const int l = K / 8;
  return K % 8 == 0 && l >= 1 && l <= 32 && (l & (l - 1)) == 0 && N >= 1 && N <= PW_MAXN;
}

extern "C" int mtus_pointwise_conv_fwd(const void* x, const float* w, const float* bias, float* y, int B, int HW, int K, int N, int dtype,
                                       void* stream) {
  MTUS_CHECK_ARG(x && w && y && B >= 0 && HW > 0);
  if (!pw_ok(K, N)) return MTUS_ERR_UNSUPPORTED;
  MTUS_CHECK_ARG(dtype == MTUS_F32 || dtype == MTUS_BF16);
  const int64_t rows = (int64_t)B * HW;
  if (rows == 0) return MTUS_OK;
  const int ppb = 256 / (K / 8);
  const int grid = (int)std::min<int64_t>((rows + 2 * ppb - 1) / (2 * ppb), 148 * 8);
  cudaStream_t st = (cudaStream_t)stream;
#define PW_F(T_, N_) case N_: pointwise_fwd_kernel<T_, N_><<<grid, 256, 0, st>>>((const T_*)x, w, bias, y, rows, HW, K); break;
#define PW_FS(T_) switch (N) { PW_F(T_, 1) PW_F(T_, 2) PW_F(T_, 3) PW_F(T_, 4) PW_F(T_, 5) PW_F(T_, 6) PW_F(T_, 7) PW_F(T_, 8) }
  if (dtype == MTUS_F32) PW_FS(float) else PW_FS(bf16)
#undef PW_FS
#undef PW_F
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_pointwise_conv_bwd(const float* dy, const void* x, const float* w, void* dx, float* dw, float* dbias, int B, int HW, int K,
                                       int N, int dtype, void* stream) {
  MTUS_CHECK_ARG(dy && x && w && dw && dbias && B >= 0 && HW > 0);
  if (!pw_ok(K, N)) return MTUS_ERR_UNSUPPORTED;
  MTUS_CHECK_ARG(dtype == MTUS_F32 || dtype == MTUS_BF16);
  const int64_t rows = (int64_t)B * HW;
  if (rows == 0) return MTUS_OK;
  const int ppb = 256 / (K / 8);
  const size_t sm = sizeof(float) * ((size_t)N * K + N);
  cudaStream_t st = (cudaStream_t)stream;
  // one resident wave: the register count (63 .. 236 with N) decides how many 256-thread CTAs an SM holds
#define PW_B(T_, N_) case N_: { int occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pointwise_bwd_kernel<T_, N_>, 256, sm);           \
    const int g2 = (int)std::min<int64_t>((rows + 2 * ppb - 1) / (2 * ppb), 148 * (occ > 0 ? occ : 1));                                       \
    pointwise_bwd_kernel<T_, N_><<<g2, 256, sm, st>>>(dy, (const T_*)x, w, (T_*)dx, dw, dbias, rows, HW, K); } break;
#define PW_BS(T_) switch (N) { PW_B(T_, 1) PW_B(T_, 2) PW_B(T_, 3) PW_B(T_, 4) PW_B(T_, 5) PW_B(T_, 6) PW_B(T_, 7) PW_B(T_, 8) }
  if (dtype == MTUS_F32) PW_BS(float) else PW_BS(bf16)
#undef PW_BS
#undef PW_B
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}
