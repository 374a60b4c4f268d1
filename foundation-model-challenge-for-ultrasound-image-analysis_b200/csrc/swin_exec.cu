// Whole-encoder executor: one C-ABI call runs the Swin Transformer forward (or a range of stages of
// the backward) as a fixed schedule of the kernels in this library on the caller's stream.
//
// This is the native runtime that replaces the Python module tree of timm's SwinTransformer /
// FeatureListNet (patch_embed -> layers_0..3 -> 4 NHWC features) and the reference wrapper's
// NHWC->NCHW permute (/root/reference/code/models/encoders.py:103-106).  No Python, no allocation
// and no synchronisation happen between kernels, so the whole step is CUDA-graph capturable.
//
// Memory: parameters live in ONE flat fp32 buffer (timm state-dict order, each tensor 8-element
// aligned) with a bf16 shadow for the GEMM operands; gradients mirror that layout, so every stage's
// gradient is one contiguous slice that can be all-reduced while earlier stages still run backward.
// Activations saved for backward live in one caller-provided workspace laid out by plan().
#include "common.cuh"
#include "graph_cache.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace {

struct PInfo { std::string name; int64_t off; int rank; int64_t shape[4]; int64_t numel; };

struct BlockP { int64_t n1w, n1b, table, qkvw, qkvb, projw, projb, n2w, n2b, fc1w, fc1b, fc2w, fc2b; };
struct StageP { int64_t mg_nw, mg_nb, mg_red; std::vector<BlockP> blk; int64_t begin, end; };
struct BlockA { size_t mean1, rstd1, ln1, qkv, attn, lse, xmid, mean2, rstd2, ln2, h, a, xout; };
struct StageA { size_t mg_ln, mg_mean, mg_rstd, xin, feat; std::vector<BlockA> blk; };

struct Plan {
  int B, S, C0, window, dtype, backend, training;
  float eps;
  int depths[4], heads[4];
  int C[4], res[4], win[4], shift[4];
  int64_t M[4];
  size_t es;  // activation / GEMM-operand element size; the residual stream (x0, xin, xmid, xout) is always fp32
  // params
  int64_t pe_w, pe_b, pe_nw, pe_nb;
  StageP sp[4];
  int64_t n_params;
  std::vector<PInfo> pinfo;
  // activations
  size_t cols, pe_pre, pe_mean, pe_rstd, x0;
  StageA sa[4];
  size_t G, Gb, Gb2, dLN, dQKV, dH, tmpF;
  size_t dp_slot, dF[4];   // drop-path scales copied by forward; incoming feature gradients as fp32 NHWC maps
  size_t ws_bytes;
};

struct Arena {
  size_t off = 0;
  size_t take(size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; }
};

int64_t add_param(Plan& p, const std::string& name, std::initializer_list<int64_t> shape) {
  PInfo pi; pi.name = name; pi.rank = (int)shape.size(); pi.numel = 1;
  int i = 0;
  for (int64_t s : shape) { pi.shape[i++] = s; pi.numel *= s; }
  pi.off = p.n_params;
  p.n_params += (pi.numel + 7) & ~(int64_t)7;
  p.pinfo.push_back(pi);
  return pi.off;
}

bool build_plan(const mtus_swin_config* c, Plan& p) {
  if (!c || c->batch < 0 || c->img_size <= 0 || c->img_size % 4 || c->embed_dim <= 0 || c->embed_dim % 32) return false;
  if (c->dtype != MTUS_F32 && c->dtype != MTUS_BF16) return false;
  if (c->window <= 0 || c->window > 16) return false;
  p.B = c->batch; p.S = c->img_size; p.C0 = c->embed_dim; p.window = c->window; p.dtype = c->dtype;
  p.backend = c->backend; p.training = c->training; p.eps = c->ln_eps > 0 ? c->ln_eps : 1e-5f;
  p.es = (c->dtype == MTUS_BF16) ? 2 : 4;
  int r = c->img_size / 4;
  for (int i = 0; i < 4; ++i) {
    if (c->depths[i] <= 0 || c->heads[i] <= 0) return false;
    p.depths[i] = c->depths[i]; p.heads[i] = c->heads[i];
    p.C[i] = c->embed_dim << i;
    if (p.C[i] != p.heads[i] * 32) return false;       // head_dim 32 (every Swin variant)
    if (i > 0) r = (r + 1) / 2;
    p.res[i] = r;
    p.win[i] = r <= c->window ? r : c->window;          // timm _calc_window_shift
    p.shift[i] = r <= p.win[i] ? 0 : c->window / 2;
    p.M[i] = (int64_t)p.B * r * r;
  }
  // ---- parameters (timm FeatureListNet state-dict order) ----
  p.n_params = 0; p.pinfo.clear();
  p.pe_w = add_param(p, "patch_embed.proj.weight", {p.C0, 3, 4, 4});
  p.pe_b = add_param(p, "patch_embed.proj.bias", {p.C0});
  p.pe_nw = add_param(p, "patch_embed.norm.weight", {p.C0});
  p.pe_nb = add_param(p, "patch_embed.norm.bias", {p.C0});
  for (int i = 0; i < 4; ++i) {
    StageP& s = p.sp[i];
    s.begin = p.n_params;
    const std::string L = "layers_" + std::to_string(i) + ".";
    if (i > 0) {
      s.mg_nw = add_param(p, L + "downsample.norm.weight", {4 * p.C[i - 1]});
      s.mg_nb = add_param(p, L + "downsample.norm.bias", {4 * p.C[i - 1]});
      s.mg_red = add_param(p, L + "downsample.reduction.weight", {p.C[i], 4 * p.C[i - 1]});
    }
    s.blk.resize(p.depths[i]);
    const int64_t Cc = p.C[i], ntab = (int64_t)(2 * p.win[i] - 1) * (2 * p.win[i] - 1);
    for (int j = 0; j < p.depths[i]; ++j) {
      BlockP& b = s.blk[j];
      const std::string Bn = L + "blocks." + std::to_string(j) + ".";
      b.n1w = add_param(p, Bn + "norm1.weight", {Cc});
      b.n1b = add_param(p, Bn + "norm1.bias", {Cc});
      b.table = add_param(p, Bn + "attn.relative_position_bias_table", {ntab, p.heads[i]});
      b.qkvw = add_param(p, Bn + "attn.qkv.weight", {3 * Cc, Cc});
      b.qkvb = add_param(p, Bn + "attn.qkv.bias", {3 * Cc});
      b.projw = add_param(p, Bn + "attn.proj.weight", {Cc, Cc});
      b.projb = add_param(p, Bn + "attn.proj.bias", {Cc});
      b.n2w = add_param(p, Bn + "norm2.weight", {Cc});
      b.n2b = add_param(p, Bn + "norm2.bias", {Cc});
      b.fc1w = add_param(p, Bn + "mlp.fc1.weight", {4 * Cc, Cc});
      b.fc1b = add_param(p, Bn + "mlp.fc1.bias", {4 * Cc});
      b.fc2w = add_param(p, Bn + "mlp.fc2.weight", {Cc, 4 * Cc});
      b.fc2b = add_param(p, Bn + "mlp.fc2.bias", {Cc});
    }
    s.end = p.n_params;
  }
  // ---- activations ----
  Arena a;
  const size_t es = p.es;
  p.cols = a.take((size_t)p.M[0] * 64 * es);
  p.pe_pre = a.take((size_t)p.M[0] * p.C0 * es);
  p.pe_mean = a.take((size_t)p.M[0] * 4);
  p.pe_rstd = a.take((size_t)p.M[0] * 4);
  p.x0 = a.take((size_t)p.M[0] * p.C0 * 4);
  for (int i = 0; i < 4; ++i) {
    StageA& s = p.sa[i];
    const size_t MC = (size_t)p.M[i] * p.C[i] * es;      // one activation map in the operand dtype
    const size_t MX = (size_t)p.M[i] * p.C[i] * 4;       // one residual-stream map (fp32)
    const size_t ML = (size_t)p.M[i] * p.heads[i] * 4;   // log-sum-exp rows
    size_t xin = p.x0;
    if (i > 0) {
      s.mg_ln = a.take(MC * 2);                          // [M_i, 4 C_{i-1}] = [M_i, 2 C_i]
      s.mg_mean = a.take((size_t)p.M[i] * 4);
      s.mg_rstd = a.take((size_t)p.M[i] * 4);
      xin = a.take(MX);
    }
    s.xin = xin;
    s.blk.resize(p.depths[i]);
    // inference: per-stage buffers are recycled (three rotating residual-stream buffers)
    size_t sh_mean1 = 0, sh_rstd1 = 0, sh_ln1 = 0, sh_qkv = 0, sh_attn = 0, sh_lse = 0, sh_mean2 = 0, sh_rstd2 = 0, sh_ln2 = 0, sh_h = 0, sh_a = 0, rot[3] = {0, 0, 0};
    if (!p.training) {
      sh_mean1 = a.take((size_t)p.M[i] * 4); sh_rstd1 = a.take((size_t)p.M[i] * 4); sh_ln1 = a.take(MC); sh_qkv = a.take(3 * MC);
      sh_attn = a.take(MC); sh_lse = a.take(ML); sh_mean2 = a.take((size_t)p.M[i] * 4); sh_rstd2 = a.take((size_t)p.M[i] * 4); sh_ln2 = a.take(MC);
      sh_h = a.take(4 * MC); sh_a = a.take(4 * MC);
      rot[0] = xin; rot[1] = a.take(MX); rot[2] = a.take(MX);
    }
    for (int j = 0; j < p.depths[i]; ++j) {
      BlockA& b = s.blk[j];
      if (p.training) {
        b.mean1 = a.take((size_t)p.M[i] * 4); b.rstd1 = a.take((size_t)p.M[i] * 4); b.ln1 = a.take(MC); b.qkv = a.take(3 * MC);
        b.attn = a.take(MC); b.lse = a.take(ML); b.xmid = a.take(MX); b.mean2 = a.take((size_t)p.M[i] * 4); b.rstd2 = a.take((size_t)p.M[i] * 4);
        b.ln2 = a.take(MC); b.h = a.take(4 * MC); b.a = a.take(4 * MC); b.xout = a.take(MX);
      } else {
        b.mean1 = sh_mean1; b.rstd1 = sh_rstd1; b.ln1 = sh_ln1; b.qkv = sh_qkv; b.attn = sh_attn; b.lse = sh_lse; b.mean2 = sh_mean2; b.rstd2 = sh_rstd2;
        b.ln2 = sh_ln2; b.h = sh_h; b.a = sh_a;
        b.xmid = rot[(2 * j + 1) % 3]; b.xout = rot[(2 * j + 2) % 3];   // x_in of block j is rot[(2j) % 3]
      }
    }
    // stage output as an NHWC map in the operand dtype (what the FPN consumes); fp32 mode: the stream itself
    s.feat = (p.dtype == MTUS_F32) ? s.blk.back().xout : a.take(MC);
  }
  {
    int nblk = 0;
    for (int i = 0; i < 4; ++i) nblk += p.depths[i];
    p.dp_slot = a.take((size_t)2 * nblk * (p.B > 0 ? p.B : 1) * 4);
    for (int i = 0; i < 4; ++i) p.dF[i] = p.training ? a.take((size_t)p.M[i] * p.C[i] * 4) : 0;
  }
  if (p.training) {
    const size_t E0 = (size_t)p.M[0] * p.C0;            // M_i*C_i is largest at stage 0
    p.G = a.take(E0 * 4); p.tmpF = a.take(E0 * 4); p.Gb = a.take(E0 * es); p.Gb2 = a.take(E0 * es); p.dLN = a.take(E0 * es);
    p.dQKV = a.take(3 * E0 * es); p.dH = a.take(4 * E0 * es);
  } else p.G = p.Gb = p.Gb2 = p.dLN = p.tmpF = p.dQKV = p.dH = 0;
  p.ws_bytes = a.off;
  return true;
}

size_t block_xin(const Plan& p, int i, int j) {
  if (j == 0) return p.sa[i].xin;
  return p.sa[i].blk[j - 1].xout;
}

// MTUS_TIME_KERNELS=1: in-situ timing of every executor step (warm caches, real operands): CUDA events around each
// call on its own stream, printed per call site when the executor returns (graphs off; the events serialise PDL).
struct StepTimer {
  struct Rec { const char* what; int line; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  static bool enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MTUS_TIME_KERNELS"); v = (e && atoi(e) != 0) ? 1 : 0; }
    return v == 1;
  }
  void begin(const char* what, int line, cudaStream_t st) {
    Rec r{what, line, nullptr, nullptr};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    recs.push_back(r);
  }
  void end(cudaStream_t st) { cudaEventRecord(recs.back().b, st); }
  void report(const char* title) {
    if (recs.empty()) return;
    cudaDeviceSynchronize();
    struct Agg { const char* what; int line; int n; float ms; };
    std::vector<Agg> agg;
    float total = 0.f;
    for (Rec& r : recs) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.a, r.b);
      cudaEventDestroy(r.a); cudaEventDestroy(r.b);
      total += ms;
      bool found = false;
      for (Agg& g : agg) if (g.line == r.line) { g.n++; g.ms += ms; found = true; break; }
      if (!found) agg.push_back(Agg{r.what, r.line, 1, ms});
    }
    fprintf(stderr, "== %s: %zu timed calls, %.3f ms summed\n", title, recs.size(), total);
    for (Agg& g : agg) fprintf(stderr, "  line %4d  n=%3d  total %8.1f us  avg %7.1f us  %.60s\n", g.line, g.n, g.ms * 1e3f, g.ms * 1e3f / g.n, g.what);
    recs.clear();
  }
};
StepTimer g_timer;

#define RUN_STREAM(expr, strm) do { \
    const bool t__ = StepTimer::enabled(); \
    if (t__) g_timer.begin(#expr, __LINE__, (cudaStream_t)(strm)); \
    int rc__ = (expr); \
    if (t__) g_timer.end((cudaStream_t)(strm)); \
    if (rc__ != MTUS_OK) { fprintf(stderr, "mtus swin_exec: %s -> %d (%s) at %s:%d\n", #expr, rc__, mtus_status_string(rc__), __FILE__, __LINE__); return rc__; } } while (0)
#define RUN(expr) RUN_STREAM(expr, stream)


// Weight-gradient GEMMs have no consumer inside the backward chain, so they run on a side stream and fill the SMs the
// dependent chain (dgrad -> LayerNorm backward -> attention backward -> ...) leaves idle: tails of the persistent GEMMs,
// the bandwidth-bound LayerNorm / attention kernels.  Fork / join is done with events only (graph-capturable):
//   E1 (main -> side): dH ready (after dgrad fc2)           -> wgrad fc2, wgrad fc1      -> D1 (side -> main)
//   E2 (main -> side): dQKV ready (after attention backward) -> wgrad proj, wgrad qkv     -> D2 (side -> main)
// main waits D1 before LayerNorm-1 backward overwrites Gb (read by wgrad fc2) and, one block later, before dgrad fc2
// overwrites dH; it waits D2 of the previous block before LayerNorm-2 backward overwrites Gb2 (read by wgrad proj) and
// attention backward overwrites dQKV.  MTUS_WGRAD_STREAM=0 keeps everything on the caller's stream.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t e1 = nullptr, e2 = nullptr, d1 = nullptr, d2 = nullptr;
  int state = 0;   // 0 unknown, 1 ready, -1 disabled
};
SideStream g_side[16];

SideStream* side_stream() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& ss = g_side[dev];
  if (ss.state == 0) {
    const char* e = getenv("MTUS_WGRAD_STREAM");
    if (e && atoi(e) == 0) { ss.state = -1; return nullptr; }
    bool ok = cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ss.e1, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ss.e2, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ss.d1, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ss.d2, cudaEventDisableTiming) == cudaSuccess;
    ss.state = ok ? 1 : -1;
  }
  return ss.state == 1 ? &ss : nullptr;
}


using namespace mtus_graphs;

#define CU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return (int)e__; } while (0)

}  // namespace

extern "C" int64_t mtus_swin_param_count(const mtus_swin_config* cfg) {
  Plan p;
  if (!build_plan(cfg, p)) return -1;
  return p.n_params;
}

extern "C" int64_t mtus_swin_workspace_bytes(const mtus_swin_config* cfg) {
  Plan p;
  if (!build_plan(cfg, p)) return -1;
  return (int64_t)p.ws_bytes;
}

extern "C" int64_t mtus_swin_feature_offset(const mtus_swin_config* cfg, int stage) {
  Plan p;
  if (!build_plan(cfg, p) || stage < 0 || stage > 3) return -1;
  return (int64_t)p.sa[stage].feat;
}

extern "C" int mtus_swin_param_info(const mtus_swin_config* cfg, int idx, char* name, int64_t* offset, int* rank,
                                    int64_t* shape) {
  Plan p;
  if (!build_plan(cfg, p) || idx < 0 || idx >= (int)p.pinfo.size()) return -1;
  const PInfo& pi = p.pinfo[idx];
  if (name) { strncpy(name, pi.name.c_str(), 127); name[127] = 0; }
  if (offset) *offset = pi.off;
  if (rank) *rank = pi.rank;
  if (shape) for (int i = 0; i < pi.rank; ++i) shape[i] = pi.shape[i];
  return 0;
}

extern "C" int64_t mtus_swin_param_offset(const mtus_swin_config* cfg, const char* name, int64_t* numel) {
  Plan p;
  if (!build_plan(cfg, p) || !name) return -1;
  for (const PInfo& pi : p.pinfo)
    if (pi.name == name) { if (numel) *numel = pi.numel; return pi.off; }
  return -1;
}

static int swin_forward_impl(const mtus_swin_config* cfg, const void* x, int x_is_f32, const float* params,
                             const void* params_lp, const float* droppath, void* workspace, void* const* feats,
                             int feats_layout, int feats_f32, void* stream);

static int swin_forward_impl(const mtus_swin_config* cfg, const void* x, int x_is_f32, const float* params,
                             const void* params_lp, const float* droppath, void* workspace, void* const* feats,
                             int feats_layout, int feats_f32, void* stream);

// Forward = eager prologue + cached body.  The prologue touches the arguments whose addresses change from step to
// step (the image batch, the drop-path draws): patch im2col into the workspace and a copy of the drop-path scales
// into the workspace.  Everything after that reads only the parameter blocks and the workspace, so the body's CUDA
// graph is keyed on addresses that stay put.
static int swin_forward_any(const mtus_swin_config* cfg, const void* x, int x_kind, const float* norm_mean, const float* norm_std,
                            const float* params, const void* params_lp, const float* droppath, void* workspace,
                            void* const* feats, int feats_layout, int feats_f32, void* stream);

extern "C" int mtus_swin_forward(const mtus_swin_config* cfg, const void* x, int x_is_f32, const float* params,
                                 const void* params_lp, const float* droppath, void* workspace, void* const* feats,
                                 int feats_layout, int feats_f32, void* stream) {
  return swin_forward_any(cfg, x, x_is_f32 ? 1 : 0, nullptr, nullptr, params, params_lp, droppath, workspace, feats, feats_layout, feats_f32, stream);
}

// Same forward fed by the raw uint8 [B,S,S,3] (HWC) batch: Normalize(mean, std, max 255) + ToTensor + the NCHW fp32 batch of the
// reference's input pipeline (code/train.py:35-44, 305) are fused into the patch-embed im2col.
extern "C" int mtus_swin_forward_u8(const mtus_swin_config* cfg, const void* x_u8, const float* mean3, const float* std3,
                                    const float* params, const void* params_lp, const float* droppath, void* workspace,
                                    void* const* feats, int feats_layout, int feats_f32, void* stream) {
  MTUS_CHECK_ARG(mean3 && std3);
  return swin_forward_any(cfg, x_u8, 2, mean3, std3, params, params_lp, droppath, workspace, feats, feats_layout, feats_f32, stream);
}

static int swin_forward_any(const mtus_swin_config* cfg, const void* x, int x_kind, const float* norm_mean, const float* norm_std,
                            const float* params, const void* params_lp, const float* droppath, void* workspace,
                            void* const* feats, int feats_layout, int feats_f32, void* stream) {
  Plan p;
  if (!build_plan(cfg, p)) return MTUS_ERR_BAD_ARG;
  MTUS_CHECK_ARG(x && params && workspace && feats);
  MTUS_CHECK_ARG(p.dtype == MTUS_F32 || params_lp);
  if (p.B == 0) return MTUS_OK;
  char* ws = reinterpret_cast<char*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  const int x_is_f32 = x_kind == 1;
  if (x_kind == 2) RUN(mtus_patch_embed_im2col_u8(x, norm_mean, norm_std, ws + p.cols, p.B, p.S, p.S, p.dtype, stream));
  else RUN(mtus_patch_embed_im2col(x, ws + p.cols, p.B, p.S, p.S, x_is_f32 || p.dtype == MTUS_F32, p.dtype, stream));
  const float* dp = nullptr;
  if (droppath) {
    int nblk = 0;
    for (int i = 0; i < 4; ++i) nblk += p.depths[i];
    CU(cudaMemcpyAsync(ws + p.dp_slot, droppath, (size_t)2 * nblk * p.B * 4, cudaMemcpyDeviceToDevice, st));
    dp = reinterpret_cast<const float*>(ws + p.dp_slot);
  }
  int dev = 0; cudaGetDevice(&dev);
  KeyBuilder kb;
  int fmask = 0;
  for (int i = 0; i < 4; ++i) if (feats[i]) fmask |= 1 << i;
  kb.add((int)1).add(dev).add(*cfg).add((int)(dp != nullptr)).add(fmask).end_shape().add(params).add(params_lp).add(workspace);
  const int rc = run_cached(kb, st, [&](void* s_) {
    return swin_forward_impl(cfg, x, x_is_f32, params, params_lp, dp, workspace, feats, feats_layout, feats_f32, s_);
  });
  if (rc != MTUS_OK) return rc;
  // eager epilogue: features materialised in the caller's own (fresh) tensors
  for (int i = 0; i < 4; ++i) {
    if (!feats[i]) continue;
    MTUS_CHECK_ARG(!(feats_f32 && feats_layout != 0 && p.dtype != MTUS_F32));
    RUN(mtus_convert(ws + p.sa[i].blk.back().xout, feats[i], p.B, p.res[i] * p.res[i], p.C[i], feats_layout == 0 ? 1 : 0, 1, feats_f32, p.dtype, stream));
  }
  if (StepTimer::enabled()) g_timer.report("mtus_swin_forward");
  return MTUS_OK;
}

static int swin_forward_impl(const mtus_swin_config* cfg, const void* x, int x_is_f32, const float* params,
                             const void* params_lp, const float* droppath, void* workspace, void* const* feats,
                             int feats_layout, int feats_f32, void* stream) {
  Plan p;
  if (!build_plan(cfg, p)) return MTUS_ERR_BAD_ARG;
  MTUS_CHECK_ARG(x && params && workspace && feats);
  MTUS_CHECK_ARG(p.dtype == MTUS_F32 || params_lp);
  if (p.B == 0) return MTUS_OK;
  char* ws = reinterpret_cast<char*>(workspace);
  const int dt = p.dtype, be = p.backend;
  // GEMM weights: bf16 shadow in bf16 mode, the fp32 master otherwise
  auto W = [&](int64_t off) -> const void* {
    return dt == MTUS_BF16 ? (const void*)(reinterpret_cast<const bf16*>(params_lp) + off) : (const void*)(params + off);
  };
  auto F = [&](int64_t off) { return params + off; };
  auto A = [&](size_t off) -> void* { return ws + off; };
  auto FA = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };

  // ---- patch embed: (im2col done by the caller) GEMM(K=48, +bias) -> LayerNorm (output: the fp32 residual stream) ----
  (void)x_is_f32;
  {
    mtus_gemm_desc d; memset(&d, 0, sizeof(d));
    d.a = A(p.cols); d.lda = 64; d.b = W(p.pe_w); d.ldb = 48;
    d.M = (int)p.M[0]; d.N = p.C0; d.K = 48; d.bias = F(p.pe_b);
    d.out = A(p.pe_pre); d.ld_out = p.C0; d.dtype = dt; d.backend = be;
    RUN(mtus_gemm(&d, stream));
  }
  RUN(mtus_layernorm_fwd_mixed(A(p.pe_pre), 0, F(p.pe_nw), F(p.pe_nb), A(p.x0), 1, FA(p.pe_mean), FA(p.pe_rstd), p.M[0], p.C0, p.eps, dt, stream));

  int gblk = 0;
  for (int i = 0; i < 4; ++i) {
    const int Cc = p.C[i], res = p.res[i];
    const int64_t M = p.M[i];
    const int rps = res * res;
    if (i > 0) {
      const StageA& s = p.sa[i];
      const size_t prev_out = p.sa[i - 1].blk.back().xout;
      RUN(mtus_patch_merge_ln_fwd_mixed(A(prev_out), F(p.sp[i].mg_nw), F(p.sp[i].mg_nb), A(s.mg_ln), FA(s.mg_mean), FA(s.mg_rstd), p.B,
                                        p.res[i - 1], p.res[i - 1], p.C[i - 1], p.eps, dt, stream));
      RUN(mtus_linear_fwd_stream(A(s.mg_ln), W(p.sp[i].mg_red), nullptr, FA(p.sa[i].xin), nullptr, nullptr, 1, M, Cc, 4 * p.C[i - 1], dt, be, stream));
    }
    for (int j = 0; j < p.depths[i]; ++j, ++gblk) {
      const BlockP& bp = p.sp[i].blk[j];
      const BlockA& ba = p.sa[i].blk[j];
      const size_t xin = block_xin(p, i, j);
      const int shift = (j % 2) ? p.shift[i] : 0;
      const float* dp1 = droppath ? droppath + (size_t)(2 * gblk) * p.B : nullptr;
      const float* dp2 = droppath ? droppath + (size_t)(2 * gblk + 1) * p.B : nullptr;
      RUN(mtus_layernorm_fwd_mixed(A(xin), 1, F(bp.n1w), F(bp.n1b), A(ba.ln1), 0, FA(ba.mean1), FA(ba.rstd1), M, Cc, p.eps, dt, stream));
      RUN(mtus_linear_fwd(A(ba.ln1), W(bp.qkvw), F(bp.qkvb), A(ba.qkv), nullptr, nullptr, nullptr, 1, M, 3 * Cc, Cc, dt, be, stream));
      RUN(mtus_window_attn_fwd(A(ba.qkv), F(bp.table), F(bp.qkvb), A(ba.attn), FA(ba.lse), p.B, res, res, Cc, p.heads[i], p.win[i], p.win[i],
                               shift, shift, dt, stream));
      RUN(mtus_linear_fwd_stream(A(ba.attn), W(bp.projw), F(bp.projb), FA(ba.xmid), FA(xin), dp1, rps, M, Cc, Cc, dt, be, stream));
      RUN(mtus_layernorm_fwd_mixed(A(ba.xmid), 1, F(bp.n2w), F(bp.n2b), A(ba.ln2), 0, FA(ba.mean2), FA(ba.rstd2), M, Cc, p.eps, dt, stream));
      // fc1 + GELU: activation -> a, GELU'(pre-activation) -> h (saved for the backward's mtus_linear_dgrad_dact)
      RUN(mtus_linear_fwd_gelu_dact(A(ba.ln2), W(bp.fc1w), F(bp.fc1b), A(ba.a), A(ba.h), M, 4 * Cc, Cc, dt, be, stream));
      RUN(mtus_linear_fwd_stream(A(ba.a), W(bp.fc2w), F(bp.fc2b), FA(ba.xout), FA(ba.xmid), dp2, rps, M, Cc, 4 * Cc, dt, be, stream));
    }
    // stage output (fp32 stream) -> in-workspace NHWC copy in the operand dtype for zero-copy consumers; features the
    // caller wants materialised in its own tensors are converted by the (eager) epilogue of mtus_swin_forward
    const size_t xo = p.sa[i].blk.back().xout;
    if (!feats[i] && dt != MTUS_F32) RUN(mtus_convert(A(xo), A(p.sa[i].feat), p.B, res * res, Cc, 0, 1, 0, dt, stream));
  }
  (void)feats_layout; (void)feats_f32;
  return MTUS_OK;
}

extern "C" int mtus_swin_backward(const mtus_swin_config* cfg, const float* params, const void* params_lp,
                                  const float* droppath, void* workspace, const void* const* dfeats, int dfeats_layout,
                                  int dfeats_f32, float* grads, int stage_hi, int stage_lo, void* stream) {
  MTUS_CHECK_ARG(cfg && stage_hi <= 4 && stage_lo >= 0 && stage_lo < stage_hi);
  int hi = 0, lo = 0;
  for (int i = 0; i < stage_hi; ++i) hi += cfg->depths[i];
  for (int i = 0; i < stage_lo; ++i) lo += cfg->depths[i];
  return mtus_swin_backward_blocks(cfg, params, params_lp, droppath, workspace, dfeats, dfeats_layout, dfeats_f32, grads, hi, lo, stream);
}

static int swin_backward_impl(const mtus_swin_config* cfg, const float* params, const void* params_lp,
                              const float* droppath, void* workspace, const void* const* dfeats, int dfeats_layout,
                              int dfeats_f32, float* grads, int block_hi, int block_lo, void* stream);

// Backward = eager prologue + cached body, like forward: the incoming feature gradients (fresh autograd tensors every
// step) are converted to fp32 NHWC maps in fixed workspace slots by the call that starts at the top block; the body
// reads only the slots, the parameter blocks, the workspace (drop-path scales included: forward left them there) and
// the gradient block.
extern "C" int mtus_swin_backward_blocks(const mtus_swin_config* cfg, const float* params, const void* params_lp,
                                         const float* droppath, void* workspace, const void* const* dfeats,
                                         int dfeats_layout, int dfeats_f32, float* grads, int block_hi, int block_lo,
                                         void* stream) {
  Plan p;
  if (!build_plan(cfg, p)) return MTUS_ERR_BAD_ARG;
  MTUS_CHECK_ARG(params && workspace && dfeats && grads && p.training);
  MTUS_CHECK_ARG(p.dtype == MTUS_F32 || params_lp);
  const int n_blocks = p.depths[0] + p.depths[1] + p.depths[2] + p.depths[3];
  MTUS_CHECK_ARG(block_hi <= n_blocks && block_lo >= 0 && block_lo < block_hi);
  if (p.B == 0) return MTUS_OK;
  char* ws = reinterpret_cast<char*>(workspace);
  const int dt = p.dtype;
  int mask = 0;
  for (int i = 0; i < 4; ++i) if (dfeats[i]) mask |= 1 << i;
  if (block_hi == n_blocks) {
    for (int i = 0; i < 4; ++i) {
      if (!dfeats[i]) continue;
      float* dst = reinterpret_cast<float*>(ws + p.dF[i]);
      if (dfeats_layout == 0) RUN(mtus_convert(dfeats[i], dst, p.B, p.C[i], p.res[i] * p.res[i], 1, dfeats_f32, 1, dt, stream));
      else {
        MTUS_CHECK_ARG(!(dfeats_f32 && dt != MTUS_F32));
        RUN(mtus_convert(dfeats[i], dst, p.B, p.res[i] * p.res[i], p.C[i], 0, dt == MTUS_F32, 1, dt, stream));
      }
    }
  }
  const float* dp = droppath ? reinterpret_cast<const float*>(ws + p.dp_slot) : nullptr;
  int dev = 0; cudaGetDevice(&dev);
  KeyBuilder kb;
  kb.add((int)2).add(dev).add(*cfg).add((int)(dp != nullptr)).add(mask).add(block_hi).add(block_lo).end_shape()
    .add(params).add(params_lp).add(workspace).add(grads);
  const int rc = run_cached(kb, (cudaStream_t)stream, [&](void* s_) {
    return swin_backward_impl(cfg, params, params_lp, dp, workspace, dfeats, dfeats_layout, dfeats_f32, grads, block_hi, block_lo, s_);
  });
  if (StepTimer::enabled()) g_timer.report("mtus_swin_backward");
  return rc;
}

extern "C" void mtus_graph_cache_stats(int64_t* hits, int64_t* misses, int64_t* entries) {
  if (hits) *hits = g_graph_hits;
  if (misses) *misses = g_graph_misses;
  if (entries) *entries = (int64_t)g_graphs.size();
}

// Gradient w.r.t. the input image, for a trainable module in FRONT of the encoder (the reference's input-level TaskPrompt2D,
// code/models/task_prompt.py:132-143, applied at code/models/multitask_model.py:198-199).  Valid right after a backward that
// ran down to block 0 on the same workspace: the patch-embed LayerNorm backward left d(loss)/d(conv output) in the dLN slot;
// dcols = dLN x W_pe ([M0, C0] x [C0, 48]) goes into the (now free) im2col slot and is scattered back to [B,3,S,S] fp32.
extern "C" int mtus_swin_input_grad(const mtus_swin_config* cfg, const float* params, const void* params_lp, void* workspace,
                                    float* dx, void* stream) {
  Plan p;
  if (!build_plan(cfg, p)) return MTUS_ERR_BAD_ARG;
  MTUS_CHECK_ARG(params && workspace && dx && p.training);
  MTUS_CHECK_ARG(p.dtype == MTUS_F32 || params_lp);
  if (p.B == 0) return MTUS_OK;
  char* ws = reinterpret_cast<char*>(workspace);
  const void* w = p.dtype == MTUS_BF16 ? (const void*)(reinterpret_cast<const bf16*>(params_lp) + p.pe_w) : (const void*)(params + p.pe_w);
  RUN(mtus_linear_dgrad(ws + p.dLN, w, ws + p.cols, nullptr, nullptr, 0, nullptr, p.M[0], p.C0, 48, p.dtype, p.backend, stream));
  RUN(mtus_patch_embed_col2im(ws + p.cols, 48, dx, p.B, p.S, p.S, p.dtype, stream));
  return MTUS_OK;
}

static int swin_backward_impl(const mtus_swin_config* cfg, const float* params, const void* params_lp,
                              const float* droppath, void* workspace, const void* const* dfeats, int dfeats_layout,
                              int dfeats_f32, float* grads, int block_hi, int block_lo, void* stream) {
  Plan p;
  if (!build_plan(cfg, p)) return MTUS_ERR_BAD_ARG;
  MTUS_CHECK_ARG(params && workspace && dfeats && grads && p.training);
  MTUS_CHECK_ARG(p.dtype == MTUS_F32 || params_lp);
  const int n_blocks = p.depths[0] + p.depths[1] + p.depths[2] + p.depths[3];
  MTUS_CHECK_ARG(block_hi <= n_blocks && block_lo >= 0 && block_lo < block_hi);
  if (p.B == 0) return MTUS_OK;
  char* ws = reinterpret_cast<char*>(workspace);
  const int dt = p.dtype, be = p.backend;
  cudaStream_t st = (cudaStream_t)stream;
  auto W = [&](int64_t off) -> const void* {
    return dt == MTUS_BF16 ? (const void*)(reinterpret_cast<const bf16*>(params_lp) + off) : (const void*)(params + off);
  };
  auto F = [&](int64_t off) { return params + off; };
  auto GR = [&](int64_t off) { return grads + off; };
  auto A = [&](size_t off) -> void* { return ws + off; };
  auto FA = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  float* G = FA(p.G); float* tmpF = FA(p.tmpF);
  void* Gb = A(p.Gb); void* Gb2 = A(p.Gb2); void* dLN = A(p.dLN); void* dQKV = A(p.dQKV); void* dH = A(p.dH);
  // The gradient w.r.t. the residual stream lives in fp32 (G).  Gb (MLP branch) / Gb2 (attention branch) are its copies
  // in the GEMM operand dtype, already multiplied by the drop-path scale of the branch that consumes them; whoever
  // writes them also adds their column sums to the bias gradient of that branch's output Linear, so no separate
  // cast / scale / column-sum passes exist.
  SideStream* ss = side_stream();
  void* wst = ss ? (void*)ss->s : stream;     // stream of the weight-gradient GEMMs
  bool d2_pending = false;

  // incoming feature gradients: fp32 NHWC maps in the workspace slots dF[i] (written by the caller's prologue)
  (void)dfeats_layout; (void)dfeats_f32; (void)tmpF;

  // global block index g counts blocks in forward order; this call runs blocks block_hi-1 down to block_lo and, after
  // block 0 of a stage, that stage's PatchMerging (or patch-embed) backward
  int gblk_end = n_blocks;
  for (int i = 3; i >= 0; --i) {
    const int stage_first = gblk_end - p.depths[i];           // global index of this stage's block 0
    const int j_hi = (block_hi < gblk_end ? block_hi : gblk_end) - stage_first;   // local blocks [j_lo, j_hi)
    const int j_lo = (block_lo > stage_first ? block_lo : stage_first) - stage_first;
    if (j_hi <= j_lo) { gblk_end = stage_first; continue; }
    const int Cc = p.C[i], res = p.res[i];
    const int64_t M = p.M[i];
    const int rps = res * res;
    int gblk = stage_first + j_hi - 1;
    if (i == 3 && j_hi == p.depths[3]) {  // top of the chain: G = NHWC(dfeat3) or zero, Gb = dp2 * G for the last block's fc2
      if (dfeats[3]) { cudaError_t e = cudaMemcpyAsync(G, FA(p.dF[3]), (size_t)M * Cc * 4, cudaMemcpyDeviceToDevice, st); if (e != cudaSuccess) return (int)e; }
      else { cudaError_t e = cudaMemsetAsync(G, 0, (size_t)M * Cc * 4, st); if (e != cudaSuccess) return (int)e; }
      const BlockP& bl = p.sp[3].blk.back();
      const float* dp2 = droppath ? droppath + (size_t)(2 * gblk + 1) * p.B : nullptr;
      RUN(mtus_scale_cast_colsum(G, dp2, rps, Gb, GR(bl.fc2b), M, Cc, dt, stream));
    }
    for (int j = j_hi - 1; j >= j_lo; --j, --gblk) {
      const BlockP& bp = p.sp[i].blk[j];
      const BlockA& ba = p.sa[i].blk[j];
      const size_t xin = block_xin(p, i, j);
      const int shift = (j % 2) ? p.shift[i] : 0;
      const float* dp1 = droppath ? droppath + (size_t)(2 * gblk) * p.B : nullptr;
      // ---- MLP branch (Gb = dp2 * G; fc2.bias gradient already accumulated by the producer of Gb) ----
      // fc1.bias gradient = column sums of dH: inside the GEMM epilogue (warp-shuffle reductions, 2.6 us of a 22.7 us launch at
      // stage 3 and 9.5 of 60 us at stage 1), or -- MTUS_FC1B_SIDE=1 -- as a colsum pass over dH on the weight-gradient stream
      static int fc1b_side = -1;
      if (fc1b_side < 0) { const char* e = getenv("MTUS_FC1B_SIDE"); fc1b_side = e ? atoi(e) : 0; }
      const bool side_cs = ss && fc1b_side && dt == MTUS_BF16;
      RUN(mtus_linear_dgrad_dact(Gb, W(bp.fc2w), dH, A(ba.h), side_cs ? nullptr : GR(bp.fc1b), M, Cc, 4 * Cc, dt, be, stream));
      if (ss) { CU(cudaEventRecord(ss->e1, st)); CU(cudaStreamWaitEvent(ss->s, ss->e1, 0)); }
      if (side_cs) RUN_STREAM(mtus_colsum(dH, GR(bp.fc1b), M, 4 * Cc, dt, wst), wst);
      RUN_STREAM(mtus_linear_wgrad(Gb, A(ba.a), GR(bp.fc2w), nullptr, M, Cc, 4 * Cc, dt, be, wst), wst);
      RUN_STREAM(mtus_linear_wgrad(dH, A(ba.ln2), GR(bp.fc1w), nullptr, M, 4 * Cc, Cc, dt, be, wst), wst);
      if (ss) CU(cudaEventRecord(ss->d1, ss->s));
      RUN(mtus_linear_dgrad(dH, W(bp.fc1w), dLN, nullptr, nullptr, 1, nullptr, M, 4 * Cc, Cc, dt, be, stream));
      if (ss && d2_pending) { CU(cudaStreamWaitEvent(st, ss->d2, 0)); d2_pending = false; }   // Gb2 / dQKV free again
      RUN(mtus_layernorm_bwd_mixed(dLN, 0, A(ba.xmid), 1, F(bp.n2w), FA(ba.mean2), FA(ba.rstd2), G, G, Gb2, dp1, rps, GR(bp.projb),
                                   GR(bp.n2w), GR(bp.n2b), M, Cc, dt, stream));
      // ---- attention branch (Gb2 = dp1 * G) ----
      RUN(mtus_linear_dgrad(Gb2, W(bp.projw), dLN, nullptr, nullptr, 1, nullptr, M, Cc, Cc, dt, be, stream));
      RUN(mtus_window_attn_bwd(dLN, A(ba.qkv), A(ba.attn), FA(ba.lse), F(bp.table), F(bp.qkvb), dQKV, GR(bp.table), GR(bp.qkvb), GR(bp.qkvb),
                               p.B, res, res, Cc, p.heads[i], p.win[i], p.win[i], shift, shift, dt, stream));
      if (ss) { CU(cudaEventRecord(ss->e2, st)); CU(cudaStreamWaitEvent(ss->s, ss->e2, 0)); }
      RUN_STREAM(mtus_linear_wgrad(Gb2, A(ba.attn), GR(bp.projw), nullptr, M, Cc, Cc, dt, be, wst), wst);
      RUN_STREAM(mtus_linear_wgrad(dQKV, A(ba.ln1), GR(bp.qkvw), nullptr, M, 3 * Cc, Cc, dt, be, wst), wst);
      if (ss) { CU(cudaEventRecord(ss->d2, ss->s)); d2_pending = true; }
      RUN(mtus_linear_dgrad(dQKV, W(bp.qkvw), dLN, nullptr, nullptr, 1, nullptr, M, 3 * Cc, Cc, dt, be, stream));
      // LN1 backward closes the block: the new Gb feeds the previous block's fc2 (its drop-path scale, its bias
      // gradient), or -- unscaled -- the PatchMerging reduction of this stage; stage 0 / block 0 needs no copy
      const float* dp2_prev = (j > 0 && droppath) ? droppath + (size_t)(2 * (gblk - 1) + 1) * p.B : nullptr;
      float* cs_prev = (j > 0) ? GR(p.sp[i].blk[j - 1].fc2b) : nullptr;
      void* lp = (j > 0 || i > 0) ? Gb : nullptr;
      if (ss) CU(cudaStreamWaitEvent(st, ss->d1, 0));   // wgrad fc2 / fc1 of this block have read Gb / dH
      RUN(mtus_layernorm_bwd_mixed(dLN, 0, A(xin), 1, F(bp.n1w), FA(ba.mean1), FA(ba.rstd1), G, G, lp, dp2_prev, rps, cs_prev,
                                   GR(bp.n1w), GR(bp.n1b), M, Cc, dt, stream));
    }
    // join: the stage's weight gradients are complete on the caller's stream (its all-reduce may start right after)
    if (ss && d2_pending) { CU(cudaStreamWaitEvent(st, ss->d2, 0)); d2_pending = false; }
    gblk_end = stage_first;
    if (j_lo > 0) continue;                                   // the stage's first block is not part of this call
    if (i > 0) {
      // ---- patch merging backward: reduction wgrad/dgrad, then LN backward scattered to [B,H,W,C_{i-1}] ----
      const StageA& s = p.sa[i];
      const int Cp = p.C[i - 1], rp = p.res[i - 1];
      RUN(mtus_linear_wgrad(Gb, A(s.mg_ln), GR(p.sp[i].mg_red), nullptr, M, Cc, 4 * Cp, dt, be, stream));
      RUN(mtus_linear_dgrad(Gb, W(p.sp[i].mg_red), dH, nullptr, nullptr, 1, nullptr, M, Cc, 4 * Cp, dt, be, stream));
      const float* dres = nullptr;
      if (dfeats[i - 1]) dres = FA(p.dF[i - 1]);
      const BlockP& bl = p.sp[i - 1].blk.back();                 // consumer of the new Gb: last block of stage i-1
      const float* dp2 = droppath ? droppath + (size_t)(2 * (gblk_end - 1) + 1) * p.B : nullptr;
      RUN(mtus_patch_merge_ln_bwd_mixed(dH, A(p.sa[i - 1].blk.back().xout), F(p.sp[i].mg_nw), FA(s.mg_mean), FA(s.mg_rstd), dres, G, Gb, dp2,
                                        rp * rp, GR(bl.fc2b), GR(p.sp[i].mg_nw), GR(p.sp[i].mg_nb), p.B, rp, rp, Cp, dt, stream));
    } else {
      // ---- patch embed backward: LN (dy = the fp32 stream), then conv weight / bias gradients (none w.r.t. the image) ----
      RUN(mtus_layernorm_bwd_mixed(G, 1, A(p.pe_pre), 0, F(p.pe_nw), FA(p.pe_mean), FA(p.pe_rstd), nullptr, nullptr, dLN, nullptr, 1, GR(p.pe_b),
                                   GR(p.pe_nw), GR(p.pe_nb), p.M[0], p.C0, dt, stream));
      mtus_gemm_desc d; memset(&d, 0, sizeof(d));
      d.a = dLN; d.lda = p.C0; d.a_mn_major = 1;
      d.b = A(p.cols); d.ldb = 64; d.b_mn_major = 1;
      d.M = p.C0; d.N = 48; d.K = (int)p.M[0];
      d.out = GR(p.pe_w); d.ld_out = 48; d.out_f32 = 1; d.atomic = 1;
      d.split_k = 148;
      d.dtype = dt; d.backend = be;
      RUN(mtus_gemm(&d, stream));
    }
  }
  return MTUS_OK;
}
