// Whole-decoder executor for smp's FPNDecoder: one C-ABI call per
// direction runs lateral 1x1 convs (GEMM with the nearest-x2 top-down add fused in the epilogue),
// the Conv3x3-GroupNorm(32)-ReLU(-bilinear x2) towers (implicit-GEMM convolutions on NHWC), and the
// cat/add merge fused with Dropout2d scaling and the NHWC->NCHW transpose the heads expect.
// Replaces segmentation_models_pytorch.decoders.fpn.decoder.FPNDecoder.forward as called from
// /root/reference/code/models/multitask_model.py:211,221 (constructed at decoders.py:42-49).
#include "common.cuh"
#include "graph_cache.cuh"
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

namespace {

struct PInfo { std::string name; int64_t off; int rank; int64_t shape[4]; int64_t numel; };

struct ConvL {            // one Conv3x3GNReLU layer
  int cin, size, up;      // input channels, input spatial size, followed by bilinear x2?
  int64_t w, gnw, gnb;    // parameter offsets
  size_t wf, wd;          // packed weights (activation dtype) in the workspace
  size_t t, mean, rstd, u, v;  // conv out, GN stats, post-ReLU, upsampled (v == u when !up)
};

struct Plan {
  int B, P, S, cat, dtype, backend, training;
  int cin[4], size[4];    // c2..c5
  size_t es;
  int64_t lat_w[4], lat_b[4];   // index 0..3 = p2..p5 lateral (p5 = "p5", others "p{k}.skip_conv")
  std::vector<ConvL> tower[4];  // tower[0] = seg_blocks.0 (on p5) ... tower[3] = seg_blocks.3 (on p2)
  int64_t n_params;
  std::vector<PInfo> pinfo;
  size_t lp;              // bf16 shadow of the flat params (bf16 mode)
  size_t cin_nhwc[4];     // NHWC copies of the inputs when the caller passes NCHW
  size_t pl[4];           // p2..p5
  size_t gM[4], s1[4], s2[4], gP[4], tmpL, dwp[4], gnws[4];   // backward scratch (s1 / s2 / dwp / gnws per tower: the towers run concurrently)
  size_t cs_slot;         // Dropout2d channel scales copied by forward (the caller's tensor is new every step)
  size_t ws_bytes;
  int out_channels;
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

bool build_plan(const mtus_fpn_config* c, Plan& p) {
  if (!c || c->batch < 0 || (c->dtype != MTUS_F32 && c->dtype != MTUS_BF16)) return false;
  if (c->pyramid_channels <= 0 || c->pyramid_channels % 32 || c->seg_channels <= 0 || c->seg_channels % 32) return false;
  p.B = c->batch; p.P = c->pyramid_channels; p.S = c->seg_channels; p.cat = c->merge_cat; p.dtype = c->dtype;
  p.backend = c->backend; p.training = c->training; p.es = c->dtype == MTUS_BF16 ? 2 : 4;
  for (int k = 0; k < 4; ++k) {
    p.cin[k] = c->in_channels[k]; p.size[k] = c->sizes[k];
    if (p.cin[k] <= 0 || p.cin[k] % 8 || p.size[k] <= 0) return false;
    if (k > 0 && p.size[k - 1] != 2 * p.size[k]) return false;   // smp adds nearest-x2 maps: exact doubling required
  }
  p.out_channels = p.cat ? 4 * p.S : p.S;
  // ---- parameters in smp state-dict order ----
  p.n_params = 0; p.pinfo.clear();
  p.lat_w[3] = add_param(p, "p5.weight", {p.P, p.cin[3], 1, 1});
  p.lat_b[3] = add_param(p, "p5.bias", {p.P});
  for (int k = 2; k >= 0; --k) {
    const std::string n = "p" + std::to_string(k + 2) + ".skip_conv.";
    p.lat_w[k] = add_param(p, n + "weight", {p.P, p.cin[k], 1, 1});
    p.lat_b[k] = add_param(p, n + "bias", {p.P});
  }
  for (int i = 0; i < 4; ++i) {            // seg_blocks.i works on level k = 3 - i with n_upsamples = 3 - i
    const int k = 3 - i, nup = 3 - i, nl = nup > 0 ? nup : 1;
    p.tower[i].resize(nl);
    int size = p.size[k];
    for (int l = 0; l < nl; ++l) {
      ConvL& L = p.tower[i][l];
      L.cin = l == 0 ? p.P : p.S; L.size = size; L.up = nup > 0;
      const std::string n = "seg_blocks." + std::to_string(i) + ".block." + std::to_string(l) + ".block.";
      L.w = add_param(p, n + "0.weight", {p.S, L.cin, 3, 3});
      L.gnw = add_param(p, n + "1.weight", {p.S});
      L.gnb = add_param(p, n + "1.bias", {p.S});
      if (L.up) size *= 2;
    }
    if (size != p.size[0]) return false;
  }
  // ---- workspace ----
  Arena a;
  const size_t es = p.es;
  p.lp = p.dtype == MTUS_BF16 ? a.take((size_t)p.n_params * 2) : 0;
  p.cs_slot = a.take((size_t)(p.B > 0 ? p.B : 1) * (p.cat ? 4 * p.S : p.S) * 4);
  for (int k = 0; k < 4; ++k) p.cin_nhwc[k] = a.take((size_t)p.B * p.size[k] * p.size[k] * p.cin[k] * es);
  for (int k = 0; k < 4; ++k) p.pl[k] = a.take((size_t)p.B * p.size[k] * p.size[k] * p.P * es);
  for (int i = 0; i < 4; ++i)
    for (ConvL& L : p.tower[i]) {
      const size_t px = (size_t)p.B * L.size * L.size;
      L.wf = a.take((size_t)p.S * 9 * L.cin * es);
      L.wd = a.take((size_t)p.S * 9 * L.cin * es);
      L.t = a.take(px * p.S * es);
      L.mean = a.take((size_t)p.B * 32 * 4);
      L.rstd = a.take((size_t)p.B * 32 * 4);
      L.u = a.take(px * p.S * es);
      L.v = L.up ? a.take(px * 4 * p.S * es) : L.u;
    }
  if (p.training) {
    const size_t slice = (size_t)p.B * p.size[0] * p.size[0] * p.S * es;
    for (int i = 0; i < 4; ++i) p.gM[i] = a.take(slice);
    for (int i = 0; i < 4; ++i) {
      // tower i works on maps of at most size[0]^2 pixels; tower 3 (on p2) never upsamples and needs s1 only
      p.s1[i] = a.take(slice); p.s2[i] = i < 3 ? a.take(slice) : p.s1[i];
      p.dwp[i] = a.take((size_t)p.S * 9 * (p.P > p.S ? p.P : p.S) * 4);
      p.gnws[i] = a.take((size_t)2 * p.B * 32 * 4);
    }
    for (int k = 0; k < 4; ++k) p.gP[k] = a.take((size_t)p.B * p.size[k] * p.size[k] * p.P * es);
    size_t mx = 0;
    for (int k = 0; k < 4; ++k) { const size_t v = (size_t)p.B * p.size[k] * p.size[k] * p.cin[k] * es; if (v > mx) mx = v; }
    p.tmpL = a.take(mx);
  } else { p.tmpL = 0; for (int k = 0; k < 4; ++k) p.gP[k] = p.gM[k] = p.s1[k] = p.s2[k] = p.dwp[k] = p.gnws[k] = 0; }
  p.ws_bytes = a.off;
  return true;
}

#define RUN(expr) do { int rc__ = (expr); if (rc__ != MTUS_OK) { fprintf(stderr, "mtus fpn_exec: %s -> %d (%s) at %s:%d\n", #expr, rc__, mtus_status_string(rc__), __FILE__, __LINE__); return rc__; } } while (0)

}  // namespace

extern "C" int64_t mtus_fpn_param_count(const mtus_fpn_config* cfg) {
  Plan p;
  return build_plan(cfg, p) ? p.n_params : -1;
}
extern "C" int64_t mtus_fpn_workspace_bytes(const mtus_fpn_config* cfg) {
  Plan p;
  return build_plan(cfg, p) ? (int64_t)p.ws_bytes : -1;
}
extern "C" int mtus_fpn_param_info(const mtus_fpn_config* cfg, int idx, char* name, int64_t* offset, int* rank, int64_t* shape) {
  Plan p;
  if (!build_plan(cfg, p) || idx < 0 || idx >= (int)p.pinfo.size()) return -1;
  const PInfo& pi = p.pinfo[idx];
  if (name) { strncpy(name, pi.name.c_str(), 127); name[127] = 0; }
  if (offset) *offset = pi.off;
  if (rank) *rank = pi.rank;
  if (shape) for (int i = 0; i < pi.rank; ++i) shape[i] = pi.shape[i];
  return 0;
}

// Like the encoder executor, each direction is an eager part that touches the tensors whose addresses change every
// step (Dropout2d scales, the output, the incoming gradient) plus a body that only touches the features, the parameter
// block and the workspace; the body is captured into a CUDA graph keyed on those addresses and replayed (graph_cache.cuh).
static int fpn_forward_body(const mtus_fpn_config* cfg, const void* const* feats, int feats_layout, int feats_f32,
                            const float* params, void* workspace, const void** merged, void* stream);

extern "C" int mtus_fpn_forward(const mtus_fpn_config* cfg, const void* const* feats, int feats_layout, int feats_f32,
                                const float* params, const float* chanscale, void* workspace, void* out, int out_f32,
                                void* stream) {
  return mtus_fpn_forward_film(cfg, feats, feats_layout, feats_f32, params, chanscale, nullptr, workspace, out, out_f32, stream);
}

extern "C" int64_t mtus_fpn_tower_output_offset(const mtus_fpn_config* cfg, int level) {
  Plan p;
  if (!build_plan(cfg, p) || level < 0 || level > 3) return -1;
  return (int64_t)p.tower[level].back().v;
}

// Forward with the FiLM epilogue of the reference's MultiTaskModel (multitask_model.py:214-216, film_layer.py:94-99) fused into
// the merge kernel: out = chanscale[b,c] * merged + chanshift[c], where the caller folds gamma[c] into chanscale (times the
// Dropout2d scale when active) and passes beta as chanshift.  Backward is mtus_fpn_backward unchanged (it re-reads the folded
// scale from the workspace); mtus_film_grad yields dgamma / dbeta.
extern "C" int mtus_fpn_forward_film(const mtus_fpn_config* cfg, const void* const* feats, int feats_layout, int feats_f32,
                                     const float* params, const float* chanscale, const float* chanshift, void* workspace,
                                     void* out, int out_f32, void* stream) {
  Plan p;
  if (!build_plan(cfg, p)) return MTUS_ERR_BAD_ARG;
  MTUS_CHECK_ARG(feats && params && workspace && out);
  if (p.B == 0) return MTUS_OK;
  for (int k = 0; k < 4; ++k) MTUS_CHECK_ARG(feats[k]);
  char* wsb = reinterpret_cast<char*>(workspace);
  const float* cs = nullptr;
  if (chanscale) {
    cudaError_t e = cudaMemcpyAsync(wsb + p.cs_slot, chanscale, (size_t)p.B * p.out_channels * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    cs = reinterpret_cast<const float*>(wsb + p.cs_slot);
  }
  const void* merged[4] = {nullptr, nullptr, nullptr, nullptr};
  int dev = 0; cudaGetDevice(&dev);
  mtus_graphs::KeyBuilder kb;
  kb.add((int)3).add(dev).add(*cfg).add(feats_layout).add(feats_f32).end_shape()
    .add(feats[0]).add(feats[1]).add(feats[2]).add(feats[3]).add(params).add(workspace);
  int rc = mtus_graphs::run_cached(kb, (cudaStream_t)stream, [&](void* s_) {
    return fpn_forward_body(cfg, feats, feats_layout, feats_f32, params, workspace, merged, s_);
  });
  if (rc != MTUS_OK) return rc;
  if (!merged[0]) {   // graph replay: the body did not run on the host; the tower outputs are fixed workspace slots
    for (int i = 0; i < 4; ++i) merged[i] = wsb + p.tower[i].back().v;
  }
  RUN(mtus_fpn_merge_film_fwd(merged, 4, p.cat, cs, chanshift, out, p.B, p.size[0] * p.size[0], p.S, p.dtype, out_f32 & 1, (out_f32 >> 1) & 1, stream));
  return MTUS_OK;
}

// The four Conv3x3-GN-ReLU(-bilinear) towers are independent of each other and most of their kernels under-fill the machine
// (a 7x7 map is 13 GEMM tiles, a 14x14 map 49): towers 0-2 run on three side streams, tower 3 (the 56x56 map) on the caller's
// stream.  Fork / join with events only, so the executor's CUDA-graph capture records the parallel branches.
// MTUS_FPN_STREAMS=0 keeps everything on the caller's stream.
struct TowerStreams {
  cudaStream_t s[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, done[3] = {nullptr, nullptr, nullptr};
  int state = 0;   // 0 unknown, 1 ready, -1 disabled
};
static TowerStreams g_tower_streams[16];

static TowerStreams* tower_streams() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  TowerStreams& ts = g_tower_streams[dev];
  if (ts.state == 0) {
    const char* e = getenv("MTUS_FPN_STREAMS");
    if (e && atoi(e) == 0) { ts.state = -1; return nullptr; }
    bool ok = cudaEventCreateWithFlags(&ts.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3; ++i) {
      ok = ok && cudaStreamCreateWithFlags(&ts.s[i], cudaStreamNonBlocking) == cudaSuccess;
      ok = ok && cudaEventCreateWithFlags(&ts.done[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ts.state = ok ? 1 : -1;
  }
  return ts.state == 1 ? &ts : nullptr;
}
#define CUF(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return (int)e__; } while (0)

static int fpn_forward_body(const mtus_fpn_config* cfg, const void* const* feats, int feats_layout, int feats_f32,
                            const float* params, void* workspace, const void** merged_out, void* stream) {
  Plan p;
  if (!build_plan(cfg, p)) return MTUS_ERR_BAD_ARG;
  char* ws = reinterpret_cast<char*>(workspace);
  const int dt = p.dtype, be = p.backend;
  auto A = [&](size_t off) -> void* { return ws + off; };
  auto FA = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto F = [&](int64_t off) { return params + off; };
  if (dt == MTUS_BF16) RUN(mtus_cast_f32_to_bf16(params, A(p.lp), p.n_params, stream));
  auto W = [&](int64_t off) -> const void* {
    return dt == MTUS_BF16 ? (const void*)(reinterpret_cast<const bf16*>(ws + p.lp) + off) : (const void*)(params + off);
  };
  const void* cin[4];
  for (int k = 0; k < 4; ++k) {
    MTUS_CHECK_ARG(feats[k]);
    if (feats_layout == 0) {
      RUN(mtus_nchw_to_nhwc(feats[k], A(p.cin_nhwc[k]), p.B, p.size[k] * p.size[k], p.cin[k], dt, feats_f32, stream));
      cin[k] = A(p.cin_nhwc[k]);
    } else { MTUS_CHECK_ARG(!(feats_f32 && dt != MTUS_F32)); cin[k] = feats[k]; }
  }
  // lateral 1x1 convs, top-down: p5 = conv(c5); p_k = conv(c_k) + nearest_x2(p_{k+1})
  for (int k = 3; k >= 0; --k) {
    mtus_gemm_desc d; memset(&d, 0, sizeof(d));
    d.a = cin[k]; d.lda = p.cin[k]; d.b = W(p.lat_w[k]); d.ldb = p.cin[k];
    d.M = p.B * p.size[k] * p.size[k]; d.N = p.P; d.K = p.cin[k]; d.bias = F(p.lat_b[k]);
    // bf16: the persistent TMA-in / TMA-out GEMM engine has no gathered-residual epilogue, so the nearest-x2 top-down
    // add is one extra in-place elementwise pass (still ~4x faster than the one-tile-per-CTA engine that fuses it);
    // fp32 (parity mode): fused in the SIMT epilogue
    const bool split_add = (k < 3) && dt == MTUS_BF16 && be != MTUS_BACKEND_SIMT;
    if (k < 3 && !split_add) { d.res = A(p.pl[k + 1]); d.ld_res = p.P; d.res_mode = 2; d.res_h = p.size[k]; d.res_w = p.size[k]; }
    d.out = A(p.pl[k]); d.ld_out = p.P; d.dtype = dt; d.backend = be;
    RUN(mtus_gemm(&d, stream));
    if (split_add) RUN(mtus_upsample_add_fwd(A(p.pl[k]), A(p.pl[k + 1]), A(p.pl[k]), p.B, p.size[k], p.size[k], p.P, dt, stream));
  }
  // towers (concurrent: see TowerStreams)
  TowerStreams* ts = tower_streams();
  cudaStream_t st = (cudaStream_t)stream;
  if (ts) CUF(cudaEventRecord(ts->fork, st));
  const void* merged[4];
  for (int i = 0; i < 4; ++i) {
    void* tstream = (ts && i < 3) ? (void*)ts->s[i] : stream;
    if (ts && i < 3) CUF(cudaStreamWaitEvent(ts->s[i], ts->fork, 0));
    const void* x = A(p.pl[3 - i]);
    for (const ConvL& L : p.tower[i]) {
      RUN(mtus_conv3x3_repack(F(L.w), A(L.wf), p.training ? A(L.wd) : nullptr, p.S, L.cin, dt, tstream));
      RUN(mtus_conv3x3_fwd(x, A(L.wf), A(L.t), p.B, L.size, L.size, L.cin, p.S, dt, be, tstream));
      // statistics + normalise + ReLU in one cluster kernel (one read, one write of the map) where a sample fits in shared memory
      RUN(mtus_groupnorm_act_fused_fwd(A(L.t), F(L.gnw), F(L.gnb), A(L.u), FA(L.mean), FA(L.rstd), p.B, L.size * L.size, p.S, 32, 1e-5f, 0, dt, tstream));
      if (L.up) RUN(mtus_bilinear2x_fwd(A(L.u), A(L.v), p.B, L.size, L.size, p.S, dt, tstream));
      x = A(L.v);
    }
    merged[i] = x;
    if (ts && i < 3) CUF(cudaEventRecord(ts->done[i], ts->s[i]));
  }
  if (ts) for (int i = 0; i < 3; ++i) CUF(cudaStreamWaitEvent(st, ts->done[i], 0));
  for (int i = 0; i < 4; ++i) merged_out[i] = merged[i];
  return MTUS_OK;
}

static int fpn_backward_body(const mtus_fpn_config* cfg, const void* const* feats, int feats_layout, int feats_f32,
                             const float* params, void* workspace, void* const* dfeats, int dfeats_layout, int dfeats_f32,
                             float* grads, void* stream);

extern "C" int mtus_fpn_backward(const mtus_fpn_config* cfg, const void* const* feats, int feats_layout, int feats_f32,
                                 const float* params, const float* chanscale, void* workspace, const void* dout,
                                 int dout_f32, void* const* dfeats, int dfeats_layout, int dfeats_f32, float* grads,
                                 void* stream) {
  Plan p;
  if (!build_plan(cfg, p)) return MTUS_ERR_BAD_ARG;
  MTUS_CHECK_ARG(params && workspace && dout && dfeats && grads && p.training && feats);
  if (p.B == 0) return MTUS_OK;
  for (int k = 0; k < 4; ++k) MTUS_CHECK_ARG(dfeats[k] && (feats_layout == 0 || feats[k]));
  char* wsb = reinterpret_cast<char*>(workspace);
  // eager: merge backward reads the incoming gradient (a fresh autograd tensor); the Dropout2d scales are the copy
  // forward left in the workspace (chanscale only says whether dropout was active)
  {
    void* dm[4];
    for (int i = 0; i < 4; ++i) dm[i] = wsb + p.gM[i];
    const float* cs = chanscale ? reinterpret_cast<const float*>(wsb + p.cs_slot) : nullptr;
    RUN(mtus_fpn_merge_bwd(dout, 4, p.cat, cs, dm, p.B, p.size[0] * p.size[0], p.S, p.dtype, dout_f32 & 1, (dout_f32 >> 1) & 1, stream));
  }
  int dev = 0; cudaGetDevice(&dev);
  mtus_graphs::KeyBuilder kb;
  kb.add((int)4).add(dev).add(*cfg).add(feats_layout).add(feats_f32).add(dfeats_layout).add(dfeats_f32).end_shape()
    .add(feats[0]).add(feats[1]).add(feats[2]).add(feats[3]).add(params).add(workspace)
    .add(dfeats[0]).add(dfeats[1]).add(dfeats[2]).add(dfeats[3]).add(grads);
  return mtus_graphs::run_cached(kb, (cudaStream_t)stream, [&](void* s_) {
    return fpn_backward_body(cfg, feats, feats_layout, feats_f32, params, workspace, dfeats, dfeats_layout, dfeats_f32, grads, s_);
  });
}

static int fpn_backward_body(const mtus_fpn_config* cfg, const void* const* feats, int feats_layout, int feats_f32,
                             const float* params, void* workspace, void* const* dfeats, int dfeats_layout, int dfeats_f32,
                             float* grads, void* stream) {
  Plan p;
  if (!build_plan(cfg, p)) return MTUS_ERR_BAD_ARG;
  (void)feats_f32;
  char* ws = reinterpret_cast<char*>(workspace);
  const int dt = p.dtype, be = p.backend;
  cudaStream_t st = (cudaStream_t)stream;
  auto A = [&](size_t off) -> void* { return ws + off; };
  auto FA = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto F = [&](int64_t off) { return params + off; };
  auto GR = [&](int64_t off) { return grads + off; };
  auto W = [&](int64_t off) -> const void* {
    return dt == MTUS_BF16 ? (const void*)(reinterpret_cast<const bf16*>(ws + p.lp) + off) : (const void*)(params + off);
  };
  // (merge backward ran eagerly in the caller: the four NHWC tower-output gradients are in gM[i])
  void* dm[4];
  for (int i = 0; i < 4; ++i) dm[i] = A(p.gM[i]);

  // towers backward: leaves the gradient w.r.t. p_k in gP[k].  Buffer discipline per layer:
  //   g --bilinear'--> s1 (if upsampling) --GN/ReLU'--> dt (s2 | s1) --dgrad--> old g buffer (dead by then)
  // (the towers are independent: towers 0-2 on the side streams, each with its own scratch; see TowerStreams)
  TowerStreams* ts = tower_streams();
  if (ts) CUF(cudaEventRecord(ts->fork, st));
  for (int i = 0; i < 4; ++i) {
    const int k = 3 - i;
    void* tstream = (ts && i < 3) ? (void*)ts->s[i] : stream;
    cudaStream_t tst = (cudaStream_t)tstream;
    if (ts && i < 3) CUF(cudaStreamWaitEvent(ts->s[i], ts->fork, 0));
    void* g = dm[i];
    for (int l = (int)p.tower[i].size() - 1; l >= 0; --l) {
      const ConvL& L = p.tower[i][l];
      const int px = L.size * L.size;
      const void* xin = l == 0 ? A(p.pl[k]) : A(p.tower[i][l - 1].v);
      const void* du = g;
      void* dtp = A(p.s1[i]);
      if (L.up) {
        RUN(mtus_bilinear2x_bwd(g, A(p.s1[i]), p.B, L.size, L.size, p.S, dt, tstream));
        du = A(p.s1[i]); dtp = A(p.s2[i]);
      }
      // ReLU gate recomputed from x (beta given, y = nullptr): one read of the map less than gating on the saved output
      RUN(mtus_groupnorm_act_bwd(du, A(L.t), nullptr, FA(L.mean), FA(L.rstd), F(L.gnw), F(L.gnb), dtp, GR(L.gnw), GR(L.gnb), FA(p.gnws[i]), p.B, px,
                                 p.S, 32, 0, dt, tstream));
      cudaError_t e = cudaMemsetAsync(A(p.dwp[i]), 0, (size_t)p.S * 9 * L.cin * 4, tst);
      if (e != cudaSuccess) return (int)e;
      RUN(mtus_conv3x3_wgrad(dtp, xin, FA(p.dwp[i]), p.B, L.size, L.size, L.cin, p.S, dt, be, tstream));
      RUN(mtus_conv3x3_unpack_grad(FA(p.dwp[i]), GR(L.w), p.S, L.cin, tstream));
      void* dxin = l == 0 ? A(p.gP[k]) : g;
      RUN(mtus_conv3x3_dgrad(dtp, A(L.wd), dxin, p.B, L.size, L.size, L.cin, p.S, dt, be, tstream));
      g = dxin;
    }
    if (ts && i < 3) CUF(cudaEventRecord(ts->done[i], ts->s[i]));
  }
  if (ts) for (int i = 0; i < 3; ++i) CUF(cudaStreamWaitEvent(st, ts->done[i], 0));
  // top-down chain: Dp_{k+1} += 2x2-sum(Dp_k), k = 0..2 (p2 -> p5)
  for (int k = 0; k < 3; ++k)
    RUN(mtus_upsample_add_bwd(A(p.gP[k]), A(p.gP[k + 1]), 1, p.B, p.size[k], p.size[k], p.P, dt, stream));
  // lateral convs backward
  for (int k = 0; k < 4; ++k) {
    const int64_t M = (int64_t)p.B * p.size[k] * p.size[k];
    const void* ck = feats_layout == 0 ? (const void*)A(p.cin_nhwc[k]) : feats[k];
    MTUS_CHECK_ARG(ck);
    RUN(mtus_linear_wgrad(A(p.gP[k]), ck, GR(p.lat_w[k]), GR(p.lat_b[k]), M, p.P, p.cin[k], dt, be, stream));
    MTUS_CHECK_ARG(dfeats[k]);
    if (dfeats_layout == 1) {
      MTUS_CHECK_ARG(!(dfeats_f32 && dt != MTUS_F32));
      RUN(mtus_linear_dgrad(A(p.gP[k]), W(p.lat_w[k]), dfeats[k], nullptr, nullptr, 1, nullptr, M, p.P, p.cin[k], dt, be, stream));
    } else {
      void* tmp = A(p.tmpL);
      RUN(mtus_linear_dgrad(A(p.gP[k]), W(p.lat_w[k]), tmp, nullptr, nullptr, 1, nullptr, M, p.P, p.cin[k], dt, be, stream));
      RUN(mtus_nhwc_to_nchw(tmp, dfeats[k], p.B, p.size[k] * p.size[k], p.cin[k], dt, dfeats_f32, stream));
    }
  }
  return MTUS_OK;
}
