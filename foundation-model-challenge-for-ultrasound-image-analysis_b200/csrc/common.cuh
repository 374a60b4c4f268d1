// Shared device helpers for the MTUS-Net B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/mtus_b200.h"

typedef __nv_bfloat16 bf16;

#define MTUS_CHECK_ARG(cond) do { if (!(cond)) return MTUS_ERR_BAD_ARG; } while (0)
// every kernel launch of this library is counted (bench.py reports it as gpu_launches)
extern "C" void mtus_internal_count_launches(int n);
#define MTUS_LAUNCH_STATUS_N(n) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; mtus_internal_count_launches(n); } while (0)
#define MTUS_LAUNCH_STATUS() MTUS_LAUNCH_STATUS_N(1)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Function attributes (cudaFuncSetAttribute: opt-in shared memory above 48 KB) are PER DEVICE: a process that drives several
// GPUs must set them once on each, not once per process.
struct mtus_per_device_flag {
  bool done[64] = {};
  bool get() const { int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false; return done[dev]; }
  void set() { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true; }
};

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// The kernels of the encoder chain (LayerNorm, GEMM, attention) are launched with programmatic stream serialization:
// kernel N+1 may be scheduled while kernel N is still running, executes its prologue (barrier / TMEM set-up, table
// loads of PARAMETERS, index tables) and then blocks in pdl_wait() until kernel N has completed and flushed.  Every
// such kernel calls pdl_trigger() first thing (so its successor can be scheduled as soon as resources free up) and
// pdl_wait() before its first access to memory an earlier kernel of the step may write.  Kernels that write
// parameters (optimizer, casts) never trigger early and are launched without the attribute.  MTUS_PDL=0 disables it.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool mtus_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t mtus_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = mtus_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- 8-wide vector load/store, fp32 compute -------------------------------------------------
// Diagnostic build (-DMTUS_DIAG_NOATOM, never shipped): the hot reduction kernels skip their global atomics (behind a
// predicate the compiler cannot fold) so the cost of the atomic tail can be read off a timing difference.
#ifdef MTUS_DIAG_NOATOM
#define MTUS_ATOMIC_ADD(p_, v_) do { const float mtus_v_ = (v_); if (mtus_v_ == 1.2345e-30f) atomicAdd((p_), mtus_v_); } while (0)
#else
#define MTUS_ATOMIC_ADD(p_, v_) atomicAdd((p_), (v_))
#endif

template <typename T> struct IO;

template <> struct IO<float> {
  static __device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  static __device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ void load8_reg(const float (&v)[8], float (&r)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = v[i];
  }
};

template <> struct IO<bf16> {
  static __device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
  static __device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
    uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
    uint2 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
    h[0] = __floats2bfloat162_rn(v[0], v[1]);
    h[1] = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = r;
  }
  static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ void load8_reg(const float (&v)[8], float (&r)[8]) {   // value after rounding to bf16
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU (timm Mlp uses nn.GELU()) and its derivative
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// GELU for the tensor-core epilogues (bf16 outputs).  The epilogues of the K = 128 / 256 GEMMs are bound by the
// instruction issue rate of the 8 epilogue warps, so the exact-erf form is replaced by a minimax fit that needs 10
// instructions: gelu(x) ~= x * sigmoid(q(x)), q(x) = x (c0 + c1 x^2 + c2 x^4) on |x| <= 5 (clamped beyond: sigmoid is 0 / 1
// to 3e-7 there).  Fitted against 0.5 x (1 + erf(x / sqrt 2)) on [-9, 9]: |error| <= 2.6e-5 absolute (below half a bf16
// ulp for every |y| >= 0.007), derivative error <= 1.1e-4; the backward uses the derivative of the SAME function.  The
// fp32 parity mode keeps erff (gelu_f / gelu_grad_f above).  Coefficients carry the -log2(e) of exp -> ex2.
#define MTUS_GELU_C0 1.59501577f
#define MTUS_GELU_C1 7.40112920e-02f
#define MTUS_GELU_C2 -7.03033580e-04f
#define MTUS_NLOG2E -1.4426950408889634f
// sigmoid(q(x)) = 0.5 + 0.5 tanh(q(x) / 2): ONE special-function op (tanh.approx, |rel err| <= 2^-11, i.e. <= 2.5e-4
// absolute on the sigmoid -- an eighth of a bf16 ulp of the result) instead of ex2 + rcp; on B200 the epilogues of the
// K = 128 / 256 GEMMs were bound by the XU pipe (MUFU + F2F) and the issue rate (profiles/r2_ncu_gemm_tc2_gelu_epilogue.txt).
__device__ __forceinline__ float gelu_sigmoid_core(float x, float& x2_out, float& xc_out, float& t_out) {
  const float xc = fminf(fmaxf(x, -5.0f), 5.0f);
  const float x2 = xc * xc;
  float p = fmaf(x2, 0.5f * MTUS_GELU_C2, 0.5f * MTUS_GELU_C1);
  p = fmaf(p, x2, 0.5f * MTUS_GELU_C0);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(p * xc));           // tanh(q / 2)
  x2_out = x2; xc_out = xc; t_out = t;
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float gelu_fast_f(float x) {
  float x2, xc, t;
  return x * gelu_sigmoid_core(x, x2, xc, t);
}
// value and derivative from ONE evaluation of the sigmoid core (the forward epilogue that stores GELU' for the backward)
__device__ __forceinline__ float gelu_both_fast_f(float x, float& dgelu) {
  float x2, xc, t;
  const float s = gelu_sigmoid_core(x, x2, xc, t);
  float dq = fmaf(x2, 5.0f * MTUS_GELU_C2, 3.0f * MTUS_GELU_C1);
  dq = fmaf(dq, x2, MTUS_GELU_C0);
  const float w = fmaf(-0.25f * t, t, 0.25f);
  dgelu = fmaf(xc * w, dq, s);
  return x * s;
}
__device__ __forceinline__ float gelu_grad_fast_f(float x) {
  float x2, xc, t;
  const float s = gelu_sigmoid_core(x, x2, xc, t);
  float dq = fmaf(x2, 5.0f * MTUS_GELU_C2, 3.0f * MTUS_GELU_C1);     // q'(x) = c0 + 3 c1 x^2 + 5 c2 x^4
  dq = fmaf(dq, x2, MTUS_GELU_C0);
  const float w = fmaf(-0.25f * t, t, 0.25f);                        // s (1 - s) = (1 - t^2) / 4
  return fmaf(xc * w, dq, s);
}

// ---- fused GEMM epilogue (shared by the SIMT fp32 GEMM and the tcgen05 bf16 GEMM) -----------
struct EpiParams {
  const float* bias;       // [N] fp32 or null
  int act;                 // 0 none | 1 GELU fwd (pre-activation also written to aux) | 2 GELU bwd (acc *= gelu'(aux))
                           // 3 GELU fwd (GELU'(pre-activation) written to aux) | 4 acc *= aux (the derivative stored by 3)
  void* aux;               // [M, ld_aux] in the activation dtype
  int64_t ld_aux;
  const void* res;         // residual in the activation dtype, or null
  int64_t ld_res;
  int res_mode;            // 1 same row | 2 nearest-x2 gather: row (b,y,x) of [B,H,W] reads row (b,y/2,x/2) of [B,H/2,W/2]
  int H, W;
  const float* rowscale;   // per-sample scale (drop-path mask / keep) or null
  int rows_per_sample;
  void* out;               // [M, ld_out]
  int64_t ld_out;
  int out_f32;             // 1: out is fp32 regardless of the activation dtype
  int atomic;              // 1: atomicAdd into the fp32 out (split-K wgrad)
  int res_f32;             // 1: the residual is fp32 regardless of the activation dtype (fp32 residual stream)
};

// Applies the epilogue to 4 consecutive columns (n0..n0+3) of row m.  nvalid = columns in range.
template <typename T>
__device__ __forceinline__ void epilogue4(const EpiParams& ep, int64_t m, int n0, int nvalid, float (&v)[4]) {
  if (ep.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (j < nvalid) v[j] += __ldg(ep.bias + n0 + j);
  }
  if (ep.act == 1 || ep.act == 3) {
    T* a = reinterpret_cast<T*>(ep.aux) + m * ep.ld_aux + n0;
    float sv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sv[j] = ep.act == 1 ? v[j] : gelu_grad_f(v[j]);
    if (nvalid == 4) IO<T>::store4(a, sv); else for (int j = 0; j < nvalid; ++j) IO<T>::st(a + j, sv[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = gelu_f(v[j]);
  } else if (ep.act == 2 || ep.act == 4) {
    const T* a = reinterpret_cast<const T*>(ep.aux) + m * ep.ld_aux + n0;
    float h[4] = {0.f, 0.f, 0.f, 0.f};
    if (nvalid == 4) IO<T>::load4(a, h); else for (int j = 0; j < nvalid; ++j) h[j] = IO<T>::ld(a + j);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= ep.act == 2 ? gelu_grad_f(h[j]) : h[j];
  }
  if (ep.rowscale) {
    const float s = __ldg(ep.rowscale + m / ep.rows_per_sample);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= s;
  }
  if (ep.res) {
    int64_t rm = m;
    if (ep.res_mode == 2) {
      const int64_t hw = (int64_t)ep.H * ep.W;
      const int64_t b = m / hw;
      const int rem = (int)(m - b * hw);
      const int y = rem / ep.W, x = rem - y * ep.W;
      rm = (b * (ep.H >> 1) + (y >> 1)) * (ep.W >> 1) + (x >> 1);
    }
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (ep.res_f32) {
      const float* r = reinterpret_cast<const float*>(ep.res) + rm * ep.ld_res + n0;
      if (nvalid == 4) IO<float>::load4(r, t); else for (int j = 0; j < nvalid; ++j) t[j] = IO<float>::ld(r + j);
    } else {
      const T* r = reinterpret_cast<const T*>(ep.res) + rm * ep.ld_res + n0;
      if (nvalid == 4) IO<T>::load4(r, t); else for (int j = 0; j < nvalid; ++j) t[j] = IO<T>::ld(r + j);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += t[j];
  }
  if (ep.out_f32) {
    float* o = reinterpret_cast<float*>(ep.out) + m * ep.ld_out + n0;
    if (ep.atomic) { for (int j = 0; j < nvalid; ++j) atomicAdd(o + j, v[j]); }
    else if (nvalid == 4) IO<float>::store4(o, v);
    else for (int j = 0; j < nvalid; ++j) o[j] = v[j];
  } else {
    T* o = reinterpret_cast<T*>(ep.out) + m * ep.ld_out + n0;
    if (nvalid == 4) IO<T>::store4(o, v); else for (int j = 0; j < nvalid; ++j) IO<T>::st(o + j, v[j]);
  }
}
