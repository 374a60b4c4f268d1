// fp32-accumulate SIMT GEMM with fused epilogues and implicit 3x3-conv operands.
//
// This is the *fp32 parity mode* engine (north_star: "fp32 mode within rtol 1e-4" -- tcgen05
// kind::tf32 keeps only 10 mantissa bits, so exact-fp32 contractions stay on the FMA pipes) and
// the bring-up engine for bf16 storage.  The bf16 production path is gemm_tc2.cu (tcgen05/TMEM/TMA).
//
// Replaces: the nn.Linear calls of timm WindowAttention.qkv/proj, Mlp.fc1/fc2,
// PatchMerging.reduction, PatchEmbed.proj (as im2col GEMM) and smp FPN 1x1 / 3x3 convs
//  plus their dgrad / wgrad (row a15).
//
// C[M,N] = sum_k A(m,k) * B(n,k).  Each operand is a stored matrix R[r][q] (q contiguous) used
// either K-major ((mn,k) = (r,q)) or MN-major ((mn,k) = (q,r)); optionally R is the implicit
// im2col matrix of an NHWC tensor (r = pixel, q = tap*C + c, zero outside the image).
#include "common.cuh"

struct GOperand {
  const void* p;
  int64_t ld;
  int conv;
  int cH, cW, cC;
};

template <typename T>
__device__ __forceinline__ void gop_load4(const GOperand& o, int64_t r, int q, int64_t R, int Q, float (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if (r >= R || q >= Q) return;
  if (!o.conv) { IO<T>::load4(reinterpret_cast<const T*>(o.p) + r * o.ld + q, v); return; }
  const int tap = q / o.cC, c = q - tap * o.cC;
  const int hw = o.cH * o.cW;
  const int64_t b = r / hw;
  const int rem = (int)(r - b * hw);
  const int y = rem / o.cW, x = rem - y * o.cW;
  const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
  if (yy < 0 || yy >= o.cH || xx < 0 || xx >= o.cW) return;
  IO<T>::load4(reinterpret_cast<const T*>(o.p) + ((b * o.cH + yy) * (int64_t)o.cW + xx) * o.cC + c, v);
}

#define SG_BM 128
#define SG_BN 128
#define SG_BK 8
#define SG_LD 132

template <typename T, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GOperand A, GOperand B, int M, int N, int K, int k_per_split,
                                                        EpiParams ep) {
  __shared__ __align__(16) float As[2][SG_BK][SG_LD];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_LD];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int ntiles = (kend - kbeg + SG_BK - 1) / SG_BK;
  const int ty = t >> 4, tx = t & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto gload = [&](int kt) {
    const int k0 = kbeg + kt * SG_BK;
    if (!A_MN) gop_load4<T>(A, (int64_t)m0 + (t >> 1), k0 + (t & 1) * 4, M, kend, ra);
    else       gop_load4<T>(A, (int64_t)k0 + (t >> 5), m0 + (t & 31) * 4, kend, M, ra);
    if (!B_MN) gop_load4<T>(B, (int64_t)n0 + (t >> 1), k0 + (t & 1) * 4, N, kend, rb);
    else       gop_load4<T>(B, (int64_t)k0 + (t >> 5), n0 + (t & 31) * 4, kend, N, rb);
  };
  auto sstore = [&](int buf) {
    if (!A_MN) {
#pragma unroll
      for (int i = 0; i < 4; ++i) As[buf][(t & 1) * 4 + i][t >> 1] = ra[i];
    } else {
      *reinterpret_cast<float4*>(&As[buf][t >> 5][(t & 31) * 4]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    }
    if (!B_MN) {
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[buf][(t & 1) * 4 + i][t >> 1] = rb[i];
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][t >> 5][(t & 31) * 4]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
  };

  if (ntiles > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  for (int kt = 0; kt < ntiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ntiles) gload(kt + 1);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ntiles) sstore(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = (int64_t)m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      const int nvalid = min(4, N - n);
      if (nvalid <= 0) continue;
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      epilogue4<T>(ep, m, n, nvalid, v);
    }
  }
}

template <typename T>
static int gemm_simt_launch(const mtus_gemm_desc* d, const EpiParams& ep, cudaStream_t st) {
  GOperand A{d->a, d->lda, d->a_conv, d->conv_h, d->conv_w, d->conv_c};
  GOperand B{d->b, d->ldb, d->b_conv, d->conv_h, d->conv_w, d->conv_c};
  int splits = d->split_k > 0 ? d->split_k : 1;
  int kps = ((d->K + splits - 1) / splits + SG_BK - 1) / SG_BK * SG_BK;
  splits = (d->K + kps - 1) / kps;
  dim3 grid(ceil_div(d->N, SG_BN), ceil_div(d->M, SG_BM), splits);
  if (grid.y > 65535) return MTUS_ERR_BAD_ARG;
  if (!d->a_mn_major && !d->b_mn_major) gemm_simt_kernel<T, false, false><<<grid, 256, 0, st>>>(A, B, d->M, d->N, d->K, kps, ep);
  else if (!d->a_mn_major && d->b_mn_major) gemm_simt_kernel<T, false, true><<<grid, 256, 0, st>>>(A, B, d->M, d->N, d->K, kps, ep);
  else if (d->a_mn_major && d->b_mn_major) gemm_simt_kernel<T, true, true><<<grid, 256, 0, st>>>(A, B, d->M, d->N, d->K, kps, ep);
  else gemm_simt_kernel<T, true, false><<<grid, 256, 0, st>>>(A, B, d->M, d->N, d->K, kps, ep);
  MTUS_LAUNCH_STATUS();
  return MTUS_OK;
}

int mtus_gemm_simt(const mtus_gemm_desc* d, const EpiParams& ep, cudaStream_t st) {
  if (d->dtype == MTUS_F32) return gemm_simt_launch<float>(d, ep, st);
  if (d->dtype == MTUS_BF16) return gemm_simt_launch<bf16>(d, ep, st);
  return MTUS_ERR_UNSUPPORTED;
}
