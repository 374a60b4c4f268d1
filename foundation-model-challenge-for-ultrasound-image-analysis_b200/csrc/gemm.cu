// GEMM dispatch (SIMT fp32 engine vs tcgen05 bf16 engine) and the nn.Linear-shaped C-ABI wrappers.
#include "common.cuh"
#include "internal.h"
#include <stdlib.h>
#include <string.h>

static int g_force_backend = -1;  // MTUS_GEMM env override: "simt" | "tc"

static int forced_backend() {
  if (g_force_backend < 0) {
    const char* e = getenv("MTUS_GEMM");
    g_force_backend = 0;
    if (e && !strcmp(e, "simt")) g_force_backend = MTUS_BACKEND_SIMT;
    if (e && !strcmp(e, "tc")) g_force_backend = MTUS_BACKEND_TCGEN05;
  }
  return g_force_backend;
}

extern "C" int mtus_gemm(const mtus_gemm_desc* d, void* stream) {
  MTUS_CHECK_ARG(d && d->a && d->b && d->out);
  MTUS_CHECK_ARG(d->M >= 0 && d->N >= 0 && d->K >= 0);
  if (d->M == 0 || d->N == 0) return MTUS_OK;
  MTUS_CHECK_ARG(d->N % 4 == 0);
  // 4-wide vector loads run along the contiguous index: k for K-major operands, m/n for MN-major ones
  if (!d->a_mn_major || !d->b_mn_major) MTUS_CHECK_ARG(d->K % 4 == 0);
  if (d->a_mn_major) MTUS_CHECK_ARG(d->M % 4 == 0);
  MTUS_CHECK_ARG(!(d->atomic && !d->out_f32));
  EpiParams ep;
  ep.bias = d->bias; ep.act = d->act; ep.aux = d->aux; ep.ld_aux = d->ld_aux;
  ep.res = d->res; ep.ld_res = d->ld_res; ep.res_mode = d->res_mode; ep.H = d->res_h; ep.W = d->res_w;
  ep.rowscale = d->rowscale; ep.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  ep.out = d->out; ep.ld_out = d->ld_out; ep.out_f32 = d->out_f32; ep.atomic = d->atomic;
  ep.res_f32 = d->res_f32;
  cudaStream_t st = (cudaStream_t)stream;
  int backend = d->backend;
  const int f = forced_backend();
  if (f) backend = f;
  if (backend == MTUS_BACKEND_AUTO) backend = (d->dtype == MTUS_BF16) ? MTUS_BACKEND_TCGEN05 : MTUS_BACKEND_SIMT;
  if (backend == MTUS_BACKEND_TCGEN05) {
    if (mtus_gemm_tc2_supported(d)) return mtus_gemm_tc2(d, st);   // persistent TMA-in / TMA-out tcgen05 engine
    if (d->backend == MTUS_BACKEND_TCGEN05 && !f) return MTUS_ERR_UNSUPPORTED;  // explicit request: fail loudly
  }
  // shapes the tensor-core engine does not take (unaligned leading dimensions, conv weight gradients with Cin % 64 != 0,
  // fp32 mode): the SIMT engine
  int rc = mtus_gemm_simt(d, ep, st);
  if (rc) return rc;
  // engines without the fused column sum: one stand-alone pass over the stored output
  if (d->out_colsum) {
    MTUS_CHECK_ARG(!d->out_f32 || d->dtype == MTUS_F32);
    return mtus_colsum(d->out, d->out_colsum, d->M, d->N, d->dtype, stream);
  }
  return MTUS_OK;
}

static int pick_splits(int64_t tiles, int64_t k_blocks) {
  // persistent engine, 1 CTA per SM: MTUS_SPLITK_WAVES (default 4) waves of work items, each split keeping >= 4 k-blocks
  static int waves = 0;
  if (!waves) { const char* e = getenv("MTUS_SPLITK_WAVES"); waves = e ? atoi(e) : 4; if (waves < 1) waves = 1; }
  int64_t want = (148 * waves + tiles - 1) / tiles;
  int64_t maxs = k_blocks / 4;
  if (maxs < 1) maxs = 1;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  return (int)want;
}

extern "C" int mtus_linear_fwd(const void* x, const void* w, const float* bias, void* y, void* gelu_pre,
                               const void* res, const float* rowscale, int rows_per_sample, int64_t M, int N, int K,
                               int dtype, int backend, void* stream) {
  MTUS_CHECK_ARG(x && w && y && M >= 0 && M < (1ll << 31));
  mtus_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = x; d.lda = K; d.b = w; d.ldb = K;
  d.M = (int)M; d.N = N; d.K = K;
  d.bias = bias;
  if (gelu_pre) { d.act = 1; d.aux = gelu_pre; d.ld_aux = N; }
  if (res) { d.res = res; d.ld_res = N; d.res_mode = 1; }
  d.rowscale = rowscale; d.rows_per_sample = rows_per_sample;
  d.out = y; d.ld_out = N;
  d.dtype = dtype; d.backend = backend;
  return mtus_gemm(&d, stream);
}

// y (fp32) = res (fp32, optional) + rowscale * (x w^T + bias): Linear layers that write the fp32 residual stream
// The MLP pair of the bf16 training path.  Forward: y = GELU(x w^T + bias) and dact = GELU'(x w^T + bias) (stored INSTEAD of the
// pre-activation: it shares the sigmoid / erf of the forward, and the backward's epilogue shrinks to one multiply).
extern "C" int mtus_linear_fwd_gelu_dact(const void* x, const void* w, const float* bias, void* y, void* dact, int64_t M, int N, int K,
                                         int dtype, int backend, void* stream) {
  MTUS_CHECK_ARG(x && w && y && dact && M >= 0 && M < (1ll << 31));
  mtus_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = x; d.lda = K; d.b = w; d.ldb = K;
  d.M = (int)M; d.N = N; d.K = K;
  d.bias = bias;
  d.act = 3; d.aux = dact; d.ld_aux = N;
  d.out = y; d.ld_out = N;
  d.dtype = dtype; d.backend = backend;
  return mtus_gemm(&d, stream);
}

// dx[M,K] = (dy[M,N] w[N,K]) * dact[M,K] (+ column sums of dx): the data gradient through fc2 and the stored GELU'.
extern "C" int mtus_linear_dgrad_dact(const void* dy, const void* w, void* dx, const void* dact, float* dx_colsum, int64_t M, int N,
                                      int K, int dtype, int backend, void* stream) {
  MTUS_CHECK_ARG(dy && w && dx && dact && M >= 0 && M < (1ll << 31));
  mtus_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = dy; d.lda = N; d.b = w; d.ldb = K; d.b_mn_major = 1;
  d.M = (int)M; d.N = K; d.K = N;
  d.act = 4; d.aux = const_cast<void*>(dact); d.ld_aux = K;
  d.out = dx; d.ld_out = K;
  d.out_colsum = dx_colsum;
  d.dtype = dtype; d.backend = backend;
  return mtus_gemm(&d, stream);
}

extern "C" int mtus_linear_fwd_stream(const void* x, const void* w, const float* bias, float* y, const float* res,
                                      const float* rowscale, int rows_per_sample, int64_t M, int N, int K, int dtype,
                                      int backend, void* stream) {
  MTUS_CHECK_ARG(x && w && y && M >= 0 && M < (1ll << 31));
  mtus_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = x; d.lda = K; d.b = w; d.ldb = K;
  d.M = (int)M; d.N = N; d.K = K;
  d.bias = bias;
  if (res) { d.res = res; d.ld_res = N; d.res_mode = 1; d.res_f32 = 1; }
  d.rowscale = rowscale; d.rows_per_sample = rows_per_sample;
  d.out = y; d.ld_out = N; d.out_f32 = 1;
  d.dtype = dtype; d.backend = backend;
  return mtus_gemm(&d, stream);
}

extern "C" int mtus_linear_dgrad(const void* dy, const void* w, void* dx, const void* gelu_pre, const float* rowscale,
                                 int rows_per_sample, float* dx_colsum, int64_t M, int N, int K, int dtype, int backend,
                                 void* stream) {
  MTUS_CHECK_ARG(dy && w && dx && M >= 0 && M < (1ll << 31));
  mtus_gemm_desc d;
  memset(&d, 0, sizeof(d));
  // dx[M,K] = sum_n dy[m,n] w[n,k]: reduction over N; B(k_out, n) = w[n][k_out] -> stored [N][K], "MN-major"
  d.a = dy; d.lda = N; d.b = w; d.ldb = K; d.b_mn_major = 1;
  d.M = (int)M; d.N = K; d.K = N;
  if (gelu_pre) { d.act = 2; d.aux = const_cast<void*>(gelu_pre); d.ld_aux = K; }
  d.rowscale = rowscale; d.rows_per_sample = rows_per_sample;
  d.out = dx; d.ld_out = K;
  d.out_colsum = dx_colsum;
  d.dtype = dtype; d.backend = backend;
  return mtus_gemm(&d, stream);
}

extern "C" int mtus_linear_wgrad(const void* dy, const void* x, float* dw, float* db, int64_t M, int N, int K,
                                 int dtype, int backend, void* stream) {
  MTUS_CHECK_ARG(dy && x && dw && M >= 0 && M < (1ll << 31));
  mtus_gemm_desc d;
  memset(&d, 0, sizeof(d));
  // dw[N,K] = sum_m dy[m,n] x[m,k]: both operands stored [M][*] -> MN-major, reduction over M
  d.a = dy; d.lda = N; d.a_mn_major = 1;
  d.b = x; d.ldb = K; d.b_mn_major = 1;
  d.M = N; d.N = K; d.K = (int)M;
  d.out = dw; d.ld_out = K; d.out_f32 = 1; d.atomic = 1;
  d.split_k = pick_splits((int64_t)ceil_div(N, 128) * ceil_div(K, 128), (M + 63) / 64);
  d.dtype = dtype; d.backend = backend;
  int rc = mtus_gemm(&d, stream);
  if (rc) return rc;
  if (db) return mtus_colsum(dy, db, M, N, dtype, stream);
  return MTUS_OK;
}

extern "C" int mtus_conv3x3_fwd(const void* x, const void* w_fwd, void* y, int B, int H, int W, int Cin, int Cout,
                                int dtype, int backend, void* stream) {
  MTUS_CHECK_ARG(x && w_fwd && y);
  mtus_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = x; d.a_conv = 1; d.conv_h = H; d.conv_w = W; d.conv_c = Cin;
  d.b = w_fwd; d.ldb = 9 * Cin;
  d.M = B * H * W; d.N = Cout; d.K = 9 * Cin;
  d.out = y; d.ld_out = Cout;
  d.dtype = dtype; d.backend = backend;
  return mtus_gemm(&d, stream);
}

extern "C" int mtus_conv3x3_dgrad(const void* dy, const void* w_dgrad, void* dx, int B, int H, int W, int Cin,
                                  int Cout, int dtype, int backend, void* stream) {
  MTUS_CHECK_ARG(dy && w_dgrad && dx);
  mtus_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = dy; d.a_conv = 1; d.conv_h = H; d.conv_w = W; d.conv_c = Cout;
  d.b = w_dgrad; d.ldb = 9 * Cout;
  d.M = B * H * W; d.N = Cin; d.K = 9 * Cout;
  d.out = dx; d.ld_out = Cin;
  d.dtype = dtype; d.backend = backend;
  return mtus_gemm(&d, stream);
}

extern "C" int mtus_conv3x3_wgrad(const void* dy, const void* x, float* dw_packed, int B, int H, int W, int Cin,
                                  int Cout, int dtype, int backend, void* stream) {
  MTUS_CHECK_ARG(dy && x && dw_packed);
  mtus_gemm_desc d;
  memset(&d, 0, sizeof(d));
  // dw[co, tap*Cin + c] = sum_pixels dy[p, co] * im2col(x)[p, tap*Cin + c]
  d.a = dy; d.lda = Cout; d.a_mn_major = 1;
  d.b = x; d.b_conv = 1; d.b_mn_major = 1; d.conv_h = H; d.conv_w = W; d.conv_c = Cin;
  d.M = Cout; d.N = 9 * Cin; d.K = B * H * W;
  d.out = dw_packed; d.ld_out = 9 * Cin; d.out_f32 = 1; d.atomic = 1;
  d.split_k = pick_splits((int64_t)ceil_div(Cout, 128) * ceil_div(9 * Cin, 128), ((int64_t)B * H * W + 63) / 64);
  d.dtype = dtype; d.backend = backend;
  return mtus_gemm(&d, stream);
}
