// Bandwidth-bound helpers: casts, column sums (bias gradients), per-sample row scaling (drop-path),
// NHWC<->NCHW transposes (the reference's permute(0,3,1,2).contiguous(), code/models/encoders.py:106)
// and the PatchEmbed im2col (timm PatchEmbed.proj as a K=48 GEMM).  All are coalesced, 16-byte
// vectorised, grid-stride kernels; HBM roofline, algorithmic bytes = one read + one write of the tensor.
#include "common.cuh"
#include <stdlib.h>

#include <atomic>

extern "C" int mtus_version(void) { return 101; }

static std::atomic<int64_t> g_launches{0};
extern "C" void mtus_internal_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool mtus_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTUS_PDL"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v == 1;
}
extern "C" int64_t mtus_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char* mtus_status_string(int s) {
  if (s == MTUS_OK) return "ok";
  if (s == MTUS_ERR_BAD_ARG) return "bad argument (shape / alignment / null pointer)";
  if (s == MTUS_ERR_UNSUPPORTED) return "unsupported configuration";
  if (s == MTUS_ERR_DRIVER) return "CUDA driver entry point / tensor-map encoding failed";
  if (s > 0) return cudaGetErrorString((cudaError_t)s);
  return "unknown";
}

static inline int grid_for(int64_t work_items, int threads, int max_blocks = 148 * 16) {
  int64_t b = (work_items + threads - 1) / threads;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (int)b;
}

__global__ void cast_kernel(const float* __restrict__ s, bf16* __restrict__ d, int64_t n8, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float v[8];
    IO<float>::load8(s + i * 8, v);
    IO<bf16>::store8(d + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n - n8 * 8)) {
    const int64_t i = n8 * 8 + threadIdx.x;
    d[i] = __float2bfloat16_rn(s[i]);
  }
}

extern "C" int mtus_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  MTUS_CHECK_ARG(src && dst && n >= 0);
  MTUS_CHECK_ARG(((uintptr_t)src & 31) == 0 && ((uintptr_t)dst & 15) == 0);
  if (n == 0) return MTUS_OK;
  cast_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n / 8, n);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// out[c] += sum_r x[r, c].  One wave of 512-thread CTAs (16 row lanes x 32 column vectors of 8), eight independent 16-byte loads
// in flight per thread, block reduction through shared memory, then ONE vector reduction (red.global.add.v4.f32) per four
// columns.  The first version finished with 4 x 148 CTAs x C scalar atomics onto the same C floats (1 KB = four L2 slices at
// C = 256): the atomic tail was 80 % of the kernel (53.7 us vs 10.7 us without it at [100352, 256] bf16, measured with the
// -DMTUS_DIAG_NOATOM build).
#define COLSUM_TY 16
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
#ifdef MTUS_DIAG_NOATOM
  if (a == 1.2345e-30f)
#endif
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(COLSUM_TY * 32) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t rows, int C,
                                                                int64_t rows_per_chunk, int vec_ok) {
  __shared__ float red[COLSUM_TY][32][9];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = min(rows, r0 + rows_per_chunk);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < C) {
    int64_t r = r0 + ty;
    for (; r + 7 * COLSUM_TY < r1; r += 8 * COLSUM_TY) {
      float v[8][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) IO<T>::load8(x + (r + j * COLSUM_TY) * C + col, v[j]);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        acc[k] += ((v[0][k] + v[1][k]) + (v[2][k] + v[3][k])) + ((v[4][k] + v[5][k]) + (v[6][k] + v[7][k]));
    }
    for (; r < r1; r += COLSUM_TY) {
      float v[8];
      IO<T>::load8(x + r * C + col, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[ty][tx][k] = acc[k];
  __syncthreads();
  // thread (ty < 2, tx) finishes columns col + 4 ty .. + 3
  if (ty < 2 && col < C) {
    float s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s[k] = 0.f;
#pragma unroll
      for (int j = 0; j < COLSUM_TY; ++j) s[k] += red[j][tx][ty * 4 + k];
    }
    float* dst = out + col + ty * 4;
    if (vec_ok) red_add_v4(dst, s[0], s[1], s[2], s[3]);
    else {
#pragma unroll
      for (int k = 0; k < 4; ++k) MTUS_ATOMIC_ADD(dst + k, s[k]);
    }
  }
}

extern "C" int mtus_colsum(const void* x, float* out, int64_t rows, int C, int dtype, void* stream) {
  MTUS_CHECK_ARG(x && out && rows >= 0 && C % 8 == 0);
  if (rows == 0) return MTUS_OK;
  const int gx = ceil_div(C, 256);
  int64_t chunks = (148 + gx - 1) / gx;                 // one wave of CTAs, one per SM
  const int64_t max_chunks = (rows + 8 * COLSUM_TY - 1) / (8 * COLSUM_TY);
  if (chunks > max_chunks) chunks = max_chunks;
  const int64_t rpc = (rows + chunks - 1) / chunks;
  dim3 grid(gx, (unsigned)((rows + rpc - 1) / rpc));
  const int vec_ok = ((uintptr_t)out & 15) == 0;
  if (dtype == MTUS_F32) colsum_kernel<float><<<grid, COLSUM_TY * 32, 0, (cudaStream_t)stream>>>((const float*)x, out, rows, C, rpc, vec_ok);
  else if (dtype == MTUS_BF16) colsum_kernel<bf16><<<grid, COLSUM_TY * 32, 0, (cudaStream_t)stream>>>((const bf16*)x, out, rows, C, rpc, vec_ok);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

template <typename T>
__global__ void scale_rows_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ s, int rps,
                                  int64_t n8, int C8) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const int64_t row = i / C8;
    const float f = __ldg(s + row / rps);
    float v[8];
    IO<T>::load8(x + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= f;
    IO<T>::store8(y + i * 8, v);
  }
}

extern "C" int mtus_scale_rows(const void* x, void* y, const float* rowscale, int rows_per_sample, int64_t rows, int C,
                               int dtype, void* stream) {
  MTUS_CHECK_ARG(x && y && rowscale && rows_per_sample > 0 && C % 8 == 0);
  const int64_t n8 = rows * (C / 8);
  if (n8 == 0) return MTUS_OK;
  if (dtype == MTUS_F32) scale_rows_kernel<float><<<grid_for(n8, 256), 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, rowscale, rows_per_sample, n8, C / 8);
  else if (dtype == MTUS_BF16) scale_rows_kernel<bf16><<<grid_for(n8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, rowscale, rows_per_sample, n8, C / 8);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t n8) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float u[8], v[8];
    IO<T>::load8(a + i * 8, u);
    IO<T>::load8(b + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] += v[k];
    IO<T>::store8(y + i * 8, u);
  }
}

extern "C" int mtus_add(const void* a, const void* b, void* y, int64_t n, int dtype, void* stream) {
  MTUS_CHECK_ARG(a && b && y && n % 8 == 0);
  if (n == 0) return MTUS_OK;
  if (dtype == MTUS_F32) add_kernel<float><<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const float*)a, (const float*)b, (float*)y, n / 8);
  else if (dtype == MTUS_BF16) add_kernel<bf16><<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, (bf16*)y, n / 8);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// ---- [B, HW, C] <-> [B, C, HW] via 32x32 shared-memory tiles (coalesced both sides) ----------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) transpose_kernel(const TI* __restrict__ x, TO* __restrict__ y, int R, int Cc) {
  // per batch: in [R][Cc] -> out [Cc][R]
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const TI* xb = x + (int64_t)b * R * Cc;
  TO* yb = y + (int64_t)b * R * Cc;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + i * 8, c = c0 + tx;
    if (r < R && c < Cc) tile[ty + i * 8][tx] = IO<TI>::ld(xb + (int64_t)r * Cc + c);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, r = r0 + tx;
    if (r < R && c < Cc) IO<TO>::st(yb + (int64_t)c * R + r, tile[tx][ty + i * 8]);
  }
}

template <typename TI, typename TO>
static int transpose_launch(const void* x, void* y, int B, int R, int Cc, cudaStream_t st) {
  if (B == 0) return MTUS_OK;
  dim3 grid(ceil_div(Cc, 32), ceil_div(R, 32), B);
  transpose_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)x, (TO*)y, R, Cc);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

extern "C" int mtus_nhwc_to_nchw(const void* x, void* y, int B, int HW, int C, int dtype, int out_f32, void* stream) {
  MTUS_CHECK_ARG(x && y && B >= 0 && HW > 0 && C > 0 && B <= 65535);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) return transpose_launch<float, float>(x, y, B, HW, C, st);
  if (dtype == MTUS_BF16) return out_f32 ? transpose_launch<bf16, float>(x, y, B, HW, C, st) : transpose_launch<bf16, bf16>(x, y, B, HW, C, st);
  return MTUS_ERR_UNSUPPORTED;
}

extern "C" int mtus_nchw_to_nhwc(const void* x, void* y, int B, int HW, int C, int dtype, int in_f32, void* stream) {
  MTUS_CHECK_ARG(x && y && B >= 0 && HW > 0 && C > 0 && B <= 65535);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTUS_F32) return transpose_launch<float, float>(x, y, B, C, HW, st);
  if (dtype == MTUS_BF16) return in_f32 ? transpose_launch<float, bf16>(x, y, B, C, HW, st) : transpose_launch<bf16, bf16>(x, y, B, C, HW, st);
  return MTUS_ERR_UNSUPPORTED;
}

// ---- PatchEmbed im2col: x [B,3,H,W] -> cols [B*(H/4)*(W/4), 64], k = c*16 + ky*4 + kx, k >= 48 zero ----
template <typename TI, typename TO>
__global__ void patch_im2col_kernel(const TI* __restrict__ x, TO* __restrict__ cols, int B, int H, int W) {
  const int Ho = H / 4, Wo = W / 4;
  const int64_t total = (int64_t)B * Ho * Wo * 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k8 = (int)(i & 7);
    const int64_t m = i >> 3;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (k8 < 6) {
      const int ox = (int)(m % Wo); const int64_t t = m / Wo; const int oy = (int)(t % Ho); const int64_t b = t / Ho;
      const int c = k8 >> 1, ky0 = (k8 & 1) * 2;
      const TI* p = x + ((b * 3 + c) * H + (oy * 4 + ky0)) * (int64_t)W + ox * 4;
      float a[4], q[4];
      IO<TI>::load4(p, a);
      IO<TI>::load4(p + W, q);
      v[0] = a[0]; v[1] = a[1]; v[2] = a[2]; v[3] = a[3]; v[4] = q[0]; v[5] = q[1]; v[6] = q[2]; v[7] = q[3];
    }
    IO<TO>::store8(cols + i * 8, v);
  }
}

extern "C" int mtus_patch_embed_im2col(const void* x, void* cols, int B, int H, int W, int x_is_f32, int dtype,
                                       void* stream) {
  MTUS_CHECK_ARG(x && cols && B >= 0 && H % 4 == 0 && W % 4 == 0);
  if (B == 0) return MTUS_OK;
  const int64_t total = (int64_t)B * (H / 4) * (W / 4) * 8;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(total, 256);
  if (dtype == MTUS_F32) patch_im2col_kernel<float, float><<<g, 256, 0, st>>>((const float*)x, (float*)cols, B, H, W);
  else if (dtype == MTUS_BF16) {
    if (x_is_f32) patch_im2col_kernel<float, bf16><<<g, 256, 0, st>>>((const float*)x, (bf16*)cols, B, H, W);
    else patch_im2col_kernel<bf16, bf16><<<g, 256, 0, st>>>((const bf16*)x, (bf16*)cols, B, H, W);
  } else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// ---- gradient w.r.t. the image: the inverse gather of patch_im2col_kernel ---------------------------------------------------------
// dcols [B*(H/4)*(W/4), ld] holds d(loss)/d(im2col operand) (k = c*16 + ky*4 + kx, 48 valid columns); the 4x4/4 patches tile the
// image, so every pixel of dx [B,3,H,W] (fp32 NCHW, what autograd hands to the module in front of the encoder) is written once.
template <typename TI>
__global__ void patch_col2im_kernel(const TI* __restrict__ dcols, int ld, float* __restrict__ dx, int B, int H, int W) {
  const int Ho = H / 4, Wo = W / 4;
  const int64_t total = (int64_t)B * Ho * Wo * 6;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k8 = (int)(i % 6);
    const int64_t m = i / 6;
    const int ox = (int)(m % Wo); const int64_t t = m / Wo; const int oy = (int)(t % Ho); const int64_t b = t / Ho;
    const int c = k8 >> 1, ky0 = (k8 & 1) * 2;
    float v[8];
    IO<TI>::load8(dcols + m * ld + k8 * 8, v);
    float* p = dx + ((b * 3 + c) * H + (oy * 4 + ky0)) * (int64_t)W + ox * 4;
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + W) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

extern "C" int mtus_patch_embed_col2im(const void* dcols, int ld, float* dx, int B, int H, int W, int dtype, void* stream) {
  MTUS_CHECK_ARG(dcols && dx && B >= 0 && H % 4 == 0 && W % 4 == 0 && ld >= 48 && ld % 8 == 0);
  if (B == 0) return MTUS_OK;
  const int64_t total = (int64_t)B * (H / 4) * (W / 4) * 6;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(total, 256);
  if (dtype == MTUS_F32) patch_col2im_kernel<float><<<g, 256, 0, st>>>((const float*)dcols, ld, dx, B, H, W);
  else if (dtype == MTUS_BF16) patch_col2im_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)dcols, ld, dx, B, H, W);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// ---- input pipeline fused into PatchEmbed's im2col: uint8 HWC image -> normalise -> GEMM operand ----------
// Replaces albumentations Normalize(mean, std, max_pixel_value=255) + ToTensorV2 on the host and the fp32 NCHW batch the
// reference copies to the device (/root/reference/code/train.py:35-44, 305): the device receives the raw uint8 [B,H,W,3]
// batch (8x fewer host->device bytes than fp32 NCHW) and the normalised pixels are written straight into the im2col
// operand -- the normalised image never exists in HBM.  v = (u8 - 255 mean_c) * (1 / (255 std_c)).
// One thread = (patch, row pair): two 12-byte reads (4 pixels x 3 channels each), three 16/32-byte operand stores.
struct NormConst { float sub[3], mul[3]; };

template <typename TO>
__global__ void patch_im2col_u8_kernel(const uint8_t* __restrict__ x, TO* __restrict__ cols, NormConst nc, int B, int H, int W) {
  const int Ho = H / 4, Wo = W / 4;
  const int64_t total = (int64_t)B * Ho * Wo * 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int pr = (int)(i & 1);                      // rows 2 pr, 2 pr + 1 of the 4x4 patch
    const int64_t m = i >> 1;
    const int ox = (int)(m % Wo); const int64_t t = m / Wo; const int oy = (int)(t % Ho); const int64_t b = t / Ho;
    const uint8_t* p0 = x + (((b * H + oy * 4 + 2 * pr) * (int64_t)W) + ox * 4) * 3;
    uint32_t r0[3], r1[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      r0[j] = __ldg(reinterpret_cast<const uint32_t*>(p0) + j);
      r1[j] = __ldg(reinterpret_cast<const uint32_t*>(p0 + (int64_t)W * 3) + j);
    }
    const uint8_t* a0 = reinterpret_cast<const uint8_t*>(r0);
    const uint8_t* a1 = reinterpret_cast<const uint8_t*>(r1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v[8];
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        v[px] = ((float)a0[px * 3 + c] - nc.sub[c]) * nc.mul[c];
        v[4 + px] = ((float)a1[px * 3 + c] - nc.sub[c]) * nc.mul[c];
      }
      IO<TO>::store8(cols + (m * 8 + c * 2 + pr) * 8, v);     // k = c*16 + ky*4 + kx
    }
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    IO<TO>::store8(cols + (m * 8 + 6 + pr) * 8, z);           // k = 48..63: padding of the K = 48 GEMM
  }
}

extern "C" int mtus_patch_embed_im2col_u8(const void* x_u8, const float* mean3, const float* std3, void* cols, int B, int H, int W,
                                          int dtype, void* stream) {
  MTUS_CHECK_ARG(x_u8 && mean3 && std3 && cols && B >= 0 && H % 4 == 0 && W % 4 == 0);
  MTUS_CHECK_ARG((reinterpret_cast<uintptr_t>(x_u8) & 3) == 0);
  if (B == 0) return MTUS_OK;
  NormConst nc;
  for (int c = 0; c < 3; ++c) { MTUS_CHECK_ARG(std3[c] > 0.f); nc.sub[c] = 255.0f * mean3[c]; nc.mul[c] = 1.0f / (255.0f * std3[c]); }
  const int64_t total = (int64_t)B * (H / 4) * (W / 4) * 2;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(total, 256);
  if (dtype == MTUS_F32) patch_im2col_u8_kernel<float><<<g, 256, 0, st>>>((const uint8_t*)x_u8, (float*)cols, nc, B, H, W);
  else if (dtype == MTUS_BF16) patch_im2col_u8_kernel<bf16><<<g, 256, 0, st>>>((const uint8_t*)x_u8, (bf16*)cols, nc, B, H, W);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// ---- fp32 gradient stream -> GEMM operand copy: y = T(rowscale[sample] * g), colsum += column sums of y ----------
template <typename T>
__global__ void __launch_bounds__(256) scale_cast_colsum_kernel(const float* __restrict__ g, const float* __restrict__ rowscale,
                                                                int rows_per_sample, T* __restrict__ y, float* __restrict__ colsum,
                                                                int64_t rows, int C, int64_t rows_per_chunk) {
  __shared__ float red[8][32][9];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = min(rows, r0 + rows_per_chunk);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < C) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float v[8], q[8];
      IO<float>::load8(g + r * C + col, v);
      const float s = rowscale ? __ldg(rowscale + r / rows_per_sample) : 1.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= s;
      IO<T>::store8(y + r * C + col, v);
      IO<T>::load8_reg(v, q);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += q[k];
    }
  }
  if (colsum) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[ty][tx][k] = acc[k];
    __syncthreads();
    if (ty == 0 && col < C) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += red[j][tx][k];
        atomicAdd(colsum + col + k, s);
      }
    }
  }
}

extern "C" int mtus_scale_cast_colsum(const float* g, const float* rowscale, int rows_per_sample, void* y, float* colsum,
                                      int64_t rows, int C, int dtype, void* stream) {
  MTUS_CHECK_ARG(g && y && rows >= 0 && C > 0 && C % 8 == 0);
  if (rows == 0) return MTUS_OK;
  const int cb = ceil_div(C, 256);
  int chunks = ceil_div(148 * 4, cb);
  const int64_t maxc = (rows + 63) / 64;
  if (chunks > maxc) chunks = (int)maxc;
  if (chunks < 1) chunks = 1;
  const int64_t rpc = (rows + chunks - 1) / chunks;
  dim3 grid(cb, (unsigned)((rows + rpc - 1) / rpc));
  const int rps = rows_per_sample > 0 ? rows_per_sample : 1;
  if (dtype == MTUS_F32) scale_cast_colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(g, rowscale, rps, (float*)y, colsum, rows, C, rpc);
  else if (dtype == MTUS_BF16) scale_cast_colsum_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(g, rowscale, rps, (bf16*)y, colsum, rows, C, rpc);
  else return MTUS_ERR_UNSUPPORTED;
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// ---- generic layout / type conversion of a [B][R][Cc] tensor: optional transpose to [B][Cc][R], fp32 <-> dtype ----
template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n8) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float v[8];
    IO<TI>::load8(x + i * 8, v);
    IO<TO>::store8(y + i * 8, v);
  }
}

extern "C" int mtus_convert(const void* x, void* y, int B, int R, int Cc, int transpose, int in_f32, int out_f32, int dtype,
                            void* stream) {
  MTUS_CHECK_ARG(x && y && B >= 0 && R > 0 && Cc > 0 && B <= 65535);
  if (B == 0) return MTUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool fi = in_f32 || dtype == MTUS_F32, fo = out_f32 || dtype == MTUS_F32;
  if (!fi || !fo) MTUS_CHECK_ARG(dtype == MTUS_BF16);
  if (transpose) {
    if (fi && fo) return transpose_launch<float, float>(x, y, B, R, Cc, st);
    if (fi) return transpose_launch<float, bf16>(x, y, B, R, Cc, st);
    if (fo) return transpose_launch<bf16, float>(x, y, B, R, Cc, st);
    return transpose_launch<bf16, bf16>(x, y, B, R, Cc, st);
  }
  const int64_t n = (int64_t)B * R * Cc;
  MTUS_CHECK_ARG(n % 8 == 0);
  if (fi == fo) {
    cudaError_t e = cudaMemcpyAsync(y, x, (size_t)n * (fi ? 4 : 2), cudaMemcpyDeviceToDevice, st);
    return e == cudaSuccess ? MTUS_OK : (int)e;
  }
  if (fi) convert_kernel<float, bf16><<<grid_for(n / 8, 256), 256, 0, st>>>((const float*)x, (bf16*)y, n / 8);
  else convert_kernel<bf16, float><<<grid_for(n / 8, 256), 256, 0, st>>>((const bf16*)x, (float*)y, n / 8);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// ---- flat AdamW (decoupled weight decay) over a contiguous fp32 parameter block, with the global-norm clip
//      coefficient read from device memory (no host sync): torch.optim.AdamW update rule (code/train.py:208,446,455) ----
// Deterministic: every block leaves its partial sum in a per-device scratch slot and the LAST block to finish adds the slots up
// in index order, so the same gradient gives the same norm bit for bit -- on every rank of a data-parallel job (replicas that
// clip with norms differing in the last bit drift apart; an atomicAdd per block would order the additions by arrival).
// The scratch is per device and calls are stream-ordered (the optimizer issues them back to back on one stream).
#define MTUS_SUMSQ_MAX_BLOCKS (148 * 8)
__device__ float g_sumsq_partial[MTUS_SUMSQ_MAX_BLOCKS];
__device__ unsigned int g_sumsq_ticket = 0;

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n4, float* __restrict__ out) {
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  acc = warp_sum(acc);
  __shared__ float red[8];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    g_sumsq_partial[blockIdx.x] = s;
    __threadfence();
    last = (atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    float s = 0.f;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += 32) s += __ldcg(&g_sumsq_partial[i]);   // fixed assignment of slots to lanes
    s = warp_sum(s);                                                                                 // fixed butterfly order
    if (threadIdx.x == 0) { *out += s; g_sumsq_ticket = 0; }
  }
}

extern "C" int mtus_sumsq(const float* g, int64_t n, float* out, void* stream) {
  MTUS_CHECK_ARG(g && out && n >= 0 && n % 4 == 0 && ((uintptr_t)g & 15) == 0);
  if (n == 0) return MTUS_OK;
  sumsq_kernel<<<grid_for(n / 4, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(g, n / 4, out);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

template <bool SHADOW>
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, int64_t n4, float lr, float beta1, float beta2, float eps,
                                                         float wd, float bc1, float bc2_sqrt, const float* __restrict__ grad_scale,
                                                         bf16* __restrict__ shadow) {
  const float gs = grad_scale ? __ldg(grad_scale) : 1.0f;
  const float decay = 1.0f - lr * wd, step = lr / bc1, inv_bc2 = 1.0f / bc2_sqrt;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; const float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = G[k] * gs;
      P[k] *= decay;
      M[k] = M[k] + (1.0f - beta1) * (gk - M[k]);          // lerp
      V[k] = beta2 * V[k] + (1.0f - beta2) * gk * gk;
      const float denom = sqrtf(V[k]) * inv_bc2 + eps;
      P[k] -= step * (M[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (SHADOW) {                         // the bf16 operand copy the next forward reads (same rounding as cast_kernel)
      uint2 r;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
      h[0] = __floats2bfloat162_rn(pp.x, pp.y); h[1] = __floats2bfloat162_rn(pp.z, pp.w);
      reinterpret_cast<uint2*>(shadow)[i] = r;
    }
  }
}

extern "C" int mtus_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                               float weight_decay, int step, const float* grad_scale, void* stream) {
  MTUS_CHECK_ARG(p && g && m && v && n >= 0 && n % 4 == 0 && step >= 1);
  MTUS_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  if (n == 0) return MTUS_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adamw_flat_kernel<false><<<grid_for(n / 4, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n / 4, lr, beta1, beta2, eps, weight_decay,
                                                                                          bc1, sqrtf(bc2), grad_scale, nullptr);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

// Same update; additionally writes the bf16 copy of the updated parameters (the encoder's operand shadow), so the next forward
// does not have to re-read the fp32 block to refresh it (0.35 GB per step for Swin-B).
extern "C" int mtus_adamw_flat_shadow(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr, float beta1,
                                      float beta2, float eps, float weight_decay, int step, const float* grad_scale, void* stream) {
  MTUS_CHECK_ARG(p && g && m && v && shadow_bf16 && n >= 0 && n % 4 == 0 && step >= 1);
  MTUS_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)shadow_bf16 & 7) == 0);
  if (n == 0) return MTUS_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adamw_flat_kernel<true><<<grid_for(n / 4, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n / 4, lr, beta1, beta2, eps, weight_decay,
                                                                                         bc1, sqrtf(bc2), grad_scale, (bf16*)shadow_bf16);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}
