/*
 * mtus_b200.h -- C ABI of the B200-native MTUS-Net hot path (Swin encoder + FPN decoder).
 *
 * The reference (HJJ-D/Foundation-Model-Challenge-for-Ultrasound-Image-Analysis) has no FFI /
 * plugin interface: its boundary for this path is the Python module API
 *   models.encoders.build_encoder          (code/models/encoders.py:665)
 *   models.decoders.build_decoders         (code/models/decoders.py:63)
 *   MultiTaskModel.forward(x, task_id)     (code/models/multitask_model.py:176)
 * and the arithmetic lives in timm / segmentation_models_pytorch.  This
 * header is the C-ABI a maintainer would bind from those call sites (ctypes stub in
 * INTEGRATION.md).  Conventions: raw device pointers, explicit sizes, an int dtype
 * (MTUS_F32 / MTUS_BF16 = storage type of activations; accumulation is always fp32;
 * parameters gamma/beta/bias/rel-pos tables are always fp32), a cudaStream_t passed as void*,
 * int status return (0 ok, <0 argument errors below, >0 a cudaError_t).  No function allocates,
 * frees or synchronises; the caller owns every buffer.
 */
#ifndef MTUS_B200_H_
#define MTUS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTUS_OK 0
#define MTUS_ERR_BAD_ARG (-1)
#define MTUS_ERR_UNSUPPORTED (-2)
#define MTUS_ERR_DRIVER (-3)

#define MTUS_F32 0
#define MTUS_BF16 1

#define MTUS_BACKEND_AUTO 0
#define MTUS_BACKEND_SIMT 1    /* fp32-FMA engine (fp32 parity mode, bring-up) */
#define MTUS_BACKEND_TCGEN05 2 /* tcgen05 + TMEM + TMA engine (bf16 storage only) */

int mtus_version(void);
const char* mtus_status_string(int status);
/* number of kernels this library has launched since it was loaded (bench.py: gpu_launches) */
int64_t mtus_launch_count(void);

/* ---- LayerNorm (timm norm1/norm2/PatchEmbed.norm; SURVEY 8a a3,a4) ------------------------- */
int mtus_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                       int64_t rows, int C, float eps, int dtype, void* stream);
/* dx = dres + LN'(dy) (dres optional); dgamma/dbeta are ACCUMULATED (caller zeroes them). */
int mtus_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                       const void* dres, void* dx, float* dgamma, float* dbeta, int64_t rows, int C, int dtype,
                       void* stream);

/* Mixed-precision variants (fp32 residual stream): x / y element types are fp32 when the *_f32 flag is set,
 * `dtype` otherwise.  Backward: dx (fp32, may be NULL) = dres (fp32, may be NULL) + LN'(dy); dx_lp (dtype, may be
 * NULL) = lp_rowscale[sample] * dx rounded to the GEMM operand type and lp_colsum [C] (may be NULL) += its column
 * sums; dgamma / dbeta ACCUMULATED. */
int mtus_layernorm_fwd_mixed(const void* x, int x_f32, const float* gamma, const float* beta, void* y, int y_f32,
                             float* mean, float* rstd, int64_t rows, int C, float eps, int dtype, void* stream);
int mtus_layernorm_bwd_mixed(const void* dy, int dy_f32, const void* x, int x_f32, const float* gamma, const float* mean,
                             const float* rstd, const float* dres, float* dx, void* dx_lp, const float* lp_rowscale,
                             int rows_per_sample, float* lp_colsum, float* dgamma, float* dbeta, int64_t rows, int C,
                             int dtype, void* stream);
/* x / dres / dx: fp32 [B,H,W,C]; y / dy: dtype [B,ceil(H/2),ceil(W/2),4C]; dx_lp: dtype [B,H,W,C]; lp_colsum [C] */
int mtus_patch_merge_ln_fwd_mixed(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                  float* rstd, int B, int H, int W, int C, float eps, int dtype, void* stream);
int mtus_patch_merge_ln_bwd_mixed(const void* dy, const void* x, const float* gamma, const float* mean,
                                  const float* rstd, const float* dres, float* dx, void* dx_lp, const float* lp_rowscale,
                                  int rows_per_sample, float* lp_colsum, float* dgamma, float* dbeta, int B, int H, int W,
                                  int C, int dtype, void* stream);

/* ---- PatchMerging gather + LayerNorm(4C) (timm PatchMerging) ----------------- */
int mtus_patch_merge_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                            float* rstd, int B, int H, int W, int C, float eps, int dtype, void* stream);
int mtus_patch_merge_ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                            const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, int B,
                            int H, int W, int C, int dtype, void* stream);

/* ---- generic fused GEMM: C[M,N] = sum_k A(m,k) B(n,k), fp32 accumulate --------------------- */
typedef struct mtus_gemm_desc {
  const void* a;  int64_t lda; int a_mn_major; int a_conv;  /* stored [rows][cols], cols contiguous */
  const void* b;  int64_t ldb; int b_mn_major; int b_conv;  /* *_mn_major=0: (mn,k)=(row,col); 1: (mn,k)=(col,row) */
  int conv_h, conv_w, conv_c;  /* NHWC geometry when a_conv/b_conv: implicit 3x3 im2col, col = tap*C + c */
  int M, N, K;
  const float* bias;           /* [N] or NULL */
  int act;                     /* 0 none | 1 GELU (pre-activation also stored to aux) | 2 multiply by GELU'(aux) */
  void* aux; int64_t ld_aux;
  const void* res; int64_t ld_res; int res_mode; int res_h, res_w; /* 1 same row | 2 nearest-x2 gather */
  const float* rowscale; int rows_per_sample; /* drop-path scale per sample */
  void* out; int64_t ld_out; int out_f32; int atomic;
  int split_k;
  int dtype;
  int backend;
  int res_f32;                 /* residual (and out, with out_f32) in fp32: the fp32 residual stream of the bf16 mode */
  float* out_colsum;           /* [N] or NULL: += column sums of the stored output (the bias gradient of the Linear
                                  that produced this GEMM's A operand in forward) */
} mtus_gemm_desc;

int mtus_gemm(const mtus_gemm_desc* desc, void* stream);

/* nn.Linear forward: y[M,N] = x[M,K] w[N,K]^T + bias, optional GELU (h = pre-activation out),
 * optional residual y = res + rowscale[m / rows_per_sample] * (...)  (timm qkv/proj/fc1/fc2/reduction; 8a a6,a7,a8). */
int mtus_linear_fwd(const void* x, const void* w, const float* bias, void* y, void* gelu_pre, const void* res,
                    const float* rowscale, int rows_per_sample, int64_t M, int N, int K, int dtype, int backend,
                    void* stream);
/* y (fp32) = res (fp32, optional) + rowscale * (x w^T + bias): the Linear layers that write the fp32 residual
 * stream (proj, fc2, PatchMerging reduction) */
int mtus_linear_fwd_stream(const void* x, const void* w, const float* bias, float* y, const float* res,
                           const float* rowscale, int rows_per_sample, int64_t M, int N, int K, int dtype, int backend,
                           void* stream);
/* dx[M,K] = dy[M,N] w[N,K]  (optionally * GELU'(gelu_pre[M,K]), optionally scaled per sample);
 * dx_colsum [K] (may be NULL) += column sums of dx = the bias gradient of the Linear that produced gelu_pre */
int mtus_linear_dgrad(const void* dy, const void* w, void* dx, const void* gelu_pre, const float* rowscale,
                      int rows_per_sample, float* dx_colsum, int64_t M, int N, int K, int dtype, int backend, void* stream);
/* The MLP pair of the training path (timm Mlp: fc1 -> GELU -> fc2, autograd of the same).  Forward: y = GELU(x w^T + bias) and
 * dact = GELU'(x w^T + bias), saved INSTEAD of the pre-activation (it shares the forward's sigmoid / erf evaluation).  Backward:
 * dx = (dy w) * dact, dx_colsum [K] (may be NULL) += column sums of dx = fc1's bias gradient. */
int mtus_linear_fwd_gelu_dact(const void* x, const void* w, const float* bias, void* y, void* dact, int64_t M, int N, int K,
                              int dtype, int backend, void* stream);
int mtus_linear_dgrad_dact(const void* dy, const void* w, void* dx, const void* dact, float* dx_colsum, int64_t M, int N, int K,
                           int dtype, int backend, void* stream);
/* dw[N,K] += dy[M,N]^T x[M,K] (fp32, accumulated); db[N] += colsum(dy) when db != NULL */
int mtus_linear_wgrad(const void* dy, const void* x, float* dw, float* db, int64_t M, int N, int K, int dtype,
                      int backend, void* stream);

/* ---- elementwise helpers -------------------------------------------------------------------- */
int mtus_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
int mtus_colsum(const void* x, float* out, int64_t rows, int C, int dtype, void* stream); /* out += */
int mtus_scale_rows(const void* x, void* y, const float* rowscale, int rows_per_sample, int64_t rows, int C,
                    int dtype, void* stream);
int mtus_add(const void* a, const void* b, void* y, int64_t n, int dtype, void* stream);
/* fp32 gradient stream -> GEMM operand: y (dtype) = rowscale[row / rows_per_sample] * g, colsum [C] (may be NULL) += column sums of y */
int mtus_scale_cast_colsum(const float* g, const float* rowscale, int rows_per_sample, void* y, float* colsum,
                           int64_t rows, int C, int dtype, void* stream);
/* [B][R][Cc] -> [B][R][Cc] (transpose = 0) or [B][Cc][R] (transpose = 1), element types fp32 (flag set) or dtype */
int mtus_convert(const void* x, void* y, int B, int R, int Cc, int transpose, int in_f32, int out_f32, int dtype,
                 void* stream);
int mtus_nhwc_to_nchw(const void* x, void* y, int B, int HW, int C, int dtype, int out_f32, void* stream);
int mtus_nchw_to_nhwc(const void* x, void* y, int B, int HW, int C, int dtype, int in_f32, void* stream);

/* ---- pointwise (1x1) convolution with N <= 8 output channels: the Conv2d(128 -> num_classes, 1) tails of the segmentation
 * and detection heads (code/models/heads.py:16-42, 404-428).  x: [B, HW, K] channels-last rows (dtype), w: [N, K] fp32,
 * bias: [N] fp32 or NULL, y / dy: [B, N, HW] fp32 (NCHW planes).  K / 8 must be a power of two <= 32.
 * bwd: dx (dtype, may be NULL) is written, dw [N, K] and dbias [N] are ACCUMULATED (+=). */
int mtus_pointwise_conv_fwd(const void* x, const float* w, const float* bias, float* y, int B, int HW, int K, int N, int dtype,
                            void* stream);
int mtus_pointwise_conv_bwd(const float* dy, const void* x, const float* w, void* dx, float* dw, float* dbias, int B, int HW, int K,
                            int N, int dtype, void* stream);

/* ---- optimizer step over the flat parameter blocks (torch.optim.AdamW rule, code/train.py:208,455) ----------- */
/* out += sum(g^2)  (global gradient norm for clip_grad_norm_, code/train.py:446) */
int mtus_sumsq(const float* g, int64_t n, float* out, void* stream);
/* p, m, v updated in place from g * (*grad_scale) (grad_scale: device scalar = clip coefficient, or NULL) */
int mtus_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, const float* grad_scale, void* stream);
/* same update, and shadow_bf16[i] = bf16(p[i]) of the UPDATED parameters (the operand copy the bf16 forward reads; saves
 * the separate mtus_cast_f32_to_bf16 pass over the block before the next forward) */
int mtus_adamw_flat_shadow(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr, float beta1,
                           float beta2, float eps, float weight_decay, int step, const float* grad_scale, void* stream);

/* ---- PatchEmbed (timm PatchEmbed; 8a a3): 4x4/4 conv as im2col (K padded 48->64) + GEMM + LN -- */
int mtus_patch_embed_im2col(const void* x_nchw, void* cols, int B, int H, int W, int x_is_f32, int dtype,
                            void* stream);

/* ---- fused shifted-window attention (timm _attn + WindowAttention; 8a a5,a6) ----------------
 * qkv: [B,H,W,3C] NHWC (output of the qkv Linear, bias included), out: [B,H,W,C].
 * The cyclic shift, zero padding to a multiple of the window, window partition/reverse, q scaling,
 * relative-position bias, shift mask (-100), softmax and P@V all happen inside the kernel; nothing
 * window-shaped is materialised in HBM.  qkv_bias ([3C], fp32) supplies q/k/v of the padded tokens
 * (timm pads AFTER norm1, so pad tokens carry only the bias); may be NULL when H,W divide by window.
 * lse (may be NULL): [B*H*W, heads] fp32 log-sum-exp of every score row, saved for the backward pass. */
int mtus_window_attn_fwd(const void* qkv, const float* rel_table, const float* qkv_bias, void* out, float* lse, int B,
                         int H, int W, int C, int heads, int win_h, int win_w, int shift_h, int shift_w, int dtype,
                         void* stream);
/* launches of the tcgen05 / TMEM / TMA forward engine (attention_tc.cu; opt-in with MTUS_ATTN_TC=1: bf16, windows <= 64 tokens,
 * unpadded maps, even head count) since process start -- lets tests assert which engine ran. */
int64_t mtus_window_attn_tc_launch_count(void);
/* out = the forward output (delta_i = dout_i . out_i); lse = what forward wrote (required by the tensor-core engine:
 * bf16, windows of <= 64 tokens; [B*H*W, heads] fp32, log2 domain; may be NULL for the general engine);
 * dqkv fully written; drel_table [(2wh-1)(2ww-1), heads] and dqkv_bias [3C] (gradient reaching the bias through
 * padded tokens; may be NULL when H,W divide by the window) are ACCUMULATED; dqkv_colsum [3C] (may be NULL)
 * += column sums of dqkv over real tokens, i.e. the qkv Linear's bias gradient. */
int mtus_window_attn_bwd(const void* dout, const void* qkv, const void* out, const float* lse, const float* rel_table,
                         const float* qkv_bias, void* dqkv, float* drel_table, float* dqkv_bias, float* dqkv_colsum,
                         int B, int H, int W, int C, int heads, int win_h, int win_w, int shift_h, int shift_w, int dtype,
                         void* stream);

/* ---- FPN pieces (smp FPNDecoder; 8a a11-a13), all NHWC --------------------------------------- */
/* nearest-x2 top-down: y[b,h,w,:] = skip[b,h,w,:] + top[b,h/2,w/2,:] */
int mtus_upsample_add_fwd(const void* skip, const void* top, void* y, int B, int H, int W, int C, int dtype,
                          void* stream);
/* dtop[b,i,j,:] = (acc ? dtop : 0) + sum of the 2x2 block of dy */
int mtus_upsample_add_bwd(const void* dy, void* dtop, int accumulate, int B, int H, int W, int C, int dtype,
                          void* stream);
/* GroupNorm(G) statistics over (H*W, C/G) per (sample, group), NHWC */
int mtus_groupnorm_stats(const void* x, float* mean, float* rstd, int B, int HW, int C, int G, float eps, int dtype,
                         void* stream);
/* y = relu(gn(x)) */
int mtus_groupnorm_relu_fwd(const void* x, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, void* y, int B, int HW, int C, int G, int dtype, void* stream);
/* dx; dgamma/dbeta ACCUMULATED.  y is the saved post-ReLU output (mask = y > 0). ws: 2*B*G floats. */
int mtus_groupnorm_relu_bwd(const void* dy, const void* x, const void* y, const float* mean, const float* rstd,
                            const float* gamma, void* dx, float* dgamma, float* dbeta, float* ws, int B, int HW,
                            int C, int G, int dtype, void* stream);
/* Same kernels with a selectable activation: act 0 = ReLU, 1 = SiLU (the reference's segmentation head stacks
 * Conv3x3 -> GroupNorm -> SiLU, code/models/heads.py:16-42).  SiLU backward recomputes the
 * pre-activation from x, so y may be NULL and beta is required; ReLU backward needs y. */
int mtus_groupnorm_act_fwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                           void* y, int B, int HW, int C, int G, int act, int dtype, void* stream);
int mtus_groupnorm_act_bwd(const void* dy, const void* x, const void* y, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, void* dx, float* dgamma, float* dbeta, float* ws,
                           int B, int HW, int C, int G, int act, int dtype, void* stream);
/* Statistics + normalise + activation in ONE call (outputs y and the saved mean / rstd [B, G]): a thread-block-cluster
 * kernel that reads each sample once into shared memory, exchanges the group sums through distributed shared memory and
 * writes y once, when a sample's slice fits; otherwise mtus_groupnorm_stats + mtus_groupnorm_act_fwd.  Replaces
 * native_group_norm + relu / silu of smp Conv3x3GNReLU and of the reference's segmentation head (code/models/heads.py:16-42).
 * mtus_groupnorm_act_bwd with act = 0: y may be NULL when beta is given (the ReLU gate is recomputed from x). */
int mtus_groupnorm_act_fused_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                 int B, int HW, int C, int G, float eps, int act, int dtype, void* stream);
/* BatchNorm2d + activation over NHWC rows [M = B*H*W, C] (act 0 = ReLU: the reference's baseline detection head,
 * code/models/heads.py:404-428).  stats: per-channel batch mean and 1/sqrt(biased var + eps).  bwd:
 * training = 1 differentiates through the batch statistics, training = 0 treats mean / rstd as constants (running
 * statistics); dgamma / dbeta ACCUMULATED; ws: 2*C floats. */
int mtus_batchnorm_stats(const void* x, float* mean, float* rstd, int64_t M, int C, float eps, int dtype, void* stream);
int mtus_batchnorm_act_fwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                           void* y, int64_t M, int C, int act, int dtype, void* stream);
int mtus_batchnorm_act_bwd(const void* dy, const void* x, const void* y, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, void* dx, float* dgamma, float* dbeta, float* ws,
                           int64_t M, int C, int act, int training, int dtype, void* stream);
/* bilinear x2, align_corners=True, NHWC */
int mtus_bilinear2x_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream);
int mtus_bilinear2x_bwd(const void* dy, void* dx, int B, int H, int W, int C, int dtype, void* stream);
/* merge: nsrc NHWC maps [B,HW,C] -> NCHW out [B, nsrc*C (cat) | C (add), HW], times chanscale[b, c_out]
 * (Dropout2d mask / (1-p); NULL = identity).  out dtype fp32 when out_f32; out_nhwc: channels-last output. */
int mtus_fpn_merge_fwd(const void* const* srcs, int nsrc, int policy_cat, const float* chanscale, void* out, int B,
                       int HW, int C, int dtype, int out_f32, int out_nhwc, void* stream);
/* merge with the FiLM epilogue (film_layer.py:94-99 applied to the decoder output, multitask_model.py:214-216; SURVEY 8f
 * N3): out = chanscale[b,c] * merged + chanshift[c] (chanshift [C_out] fp32 or NULL; gamma is folded into chanscale). */
int mtus_fpn_merge_film_fwd(const void* const* srcs, int nsrc, int policy_cat, const float* chanscale,
                            const float* chanshift, void* out, int B, int HW, int C, int dtype, int out_f32, int out_nhwc,
                            void* stream);
/* dgamma[c] += sum dout * dropscale[b,c] * merged, dbeta[c] += sum dout; dout and srcs NHWC; dropscale [B, C_out] or NULL */
int mtus_film_grad(const void* dout, int dout_f32, const void* const* srcs, int nsrc, int policy_cat,
                   const float* dropscale, float* dgamma, float* dbeta, int B, int HW, int C, int dtype, void* stream);
int mtus_fpn_merge_bwd(const void* dout, int nsrc, int policy_cat, const float* chanscale, void* const* dsrcs, int B,
                       int HW, int C, int dtype, int in_f32, int in_nhwc, void* stream);
/* repack Conv2d weight [Cout,Cin,3,3] fp32 -> fwd [Cout, 9*Cin] and dgrad [Cin, 9*Cout] (taps flipped) */
int mtus_conv3x3_repack(const float* w, void* w_fwd, void* w_dgrad, int Cout, int Cin, int dtype, void* stream);
/* dw[Cout,Cin,3,3] (fp32, ACCUMULATED) from the packed [Cout, 9*Cin] gradient */
int mtus_conv3x3_unpack_grad(const float* dw_packed, float* dw, int Cout, int Cin, void* stream);
int mtus_conv3x3_fwd(const void* x, const void* w_fwd, void* y, int B, int H, int W, int Cin, int Cout, int dtype,
                     int backend, void* stream);
int mtus_conv3x3_dgrad(const void* dy, const void* w_dgrad, void* dx, int B, int H, int W, int Cin, int Cout,
                       int dtype, int backend, void* stream);
int mtus_conv3x3_wgrad(const void* dy, const void* x, float* dw_packed, int B, int H, int W, int Cin, int Cout,
                       int dtype, int backend, void* stream);

/* ---- whole-encoder / whole-decoder executors (C++ runtime; one call per direction) ------------ */
typedef struct mtus_swin_config {
  int batch, img_size, embed_dim;
  int depths[4], heads[4];
  int window;
  int dtype;          /* activation storage dtype */
  int backend;        /* GEMM backend */
  int training;       /* 1: save activations for backward */
  float ln_eps;
} mtus_swin_config;

/* number of fp32 elements in the flat parameter / gradient buffer, and the byte size of the
 * activation workspace the caller must provide to forward (and keep until backward). */
int64_t mtus_swin_param_count(const mtus_swin_config* cfg);
int64_t mtus_swin_workspace_bytes(const mtus_swin_config* cfg);
/* offset (elements) of a named tensor in the flat buffer; names follow timm's state-dict keys
 * (patch_embed.proj.weight, layers_1.downsample.norm.weight, layers_2.blocks.3.attn.qkv.bias, ...).
 * Returns -1 when unknown.  numel written to *numel. */
int64_t mtus_swin_param_offset(const mtus_swin_config* cfg, const char* name, int64_t* numel);
/* enumerate parameters: fills name (<=127 chars), offset, rank and shape of parameter idx; returns 0/-1 */
int mtus_swin_param_info(const mtus_swin_config* cfg, int idx, char* name, int64_t* offset, int* rank,
                         int64_t* shape);
/* byte offset, inside the workspace, of the NHWC output [B,H_i,W_i,C_i] of stage i (valid after
 * forward): lets the caller expose the features as zero-copy channels-last views. */
int64_t mtus_swin_feature_offset(const mtus_swin_config* cfg, int stage);
/* x: [B,3,S,S] NCHW (fp32 when x_is_f32 else dtype); params: flat fp32; params_lp: flat bf16 copy
 * (required for dtype=BF16, produced by mtus_cast_f32_to_bf16); droppath: [n_blocks*2, B] fp32 scales
 * or NULL; feats[4]: outputs (entries may be NULL = not materialised), feats_layout 0 = NCHW
 * (the reference's permute(0,3,1,2).contiguous()), 1 = NHWC; fp32 when feats_f32 (NCHW only). */
int mtus_swin_forward(const mtus_swin_config* cfg, const void* x, int x_is_f32, const float* params,
                      const void* params_lp, const float* droppath, void* workspace, void* const* feats,
                      int feats_layout, int feats_f32, void* stream);
/* Input pipeline fused into the patch-embed operand (replaces albumentations Normalize + ToTensorV2 and the
 * fp32 NCHW host->device copy of code/train.py:35-44, 305): x_u8 is the raw uint8 [B,H,W,3] (HWC) batch, mean3 / std3 are
 * HOST arrays (Normalize(mean, std, max_pixel_value=255)); cols = the [B*(H/4)*(W/4), 64] im2col operand. */
int mtus_patch_embed_im2col_u8(const void* x_u8, const float* mean3, const float* std3, void* cols, int B, int H, int W,
                               int dtype, void* stream);
/* Inverse of mtus_patch_embed_im2col for gradients: dcols [B*(H/4)*(W/4), ld] (dtype; 48 valid columns, ld % 8 == 0) is
 * scattered to dx [B,3,H,W] fp32 NCHW; every pixel is written exactly once (no zero-fill needed). */
int mtus_patch_embed_col2im(const void* dcols, int ld, float* dx, int B, int H, int W, int dtype, void* stream);
/* mtus_swin_forward fed by the uint8 HWC batch (everything else identical). */
int mtus_swin_forward_u8(const mtus_swin_config* cfg, const void* x_u8, const float* mean3, const float* std3,
                         const float* params, const void* params_lp, const float* droppath, void* workspace,
                         void* const* feats, int feats_layout, int feats_f32, void* stream);
/* dfeats[4]: NCHW grads (NULL = zero); grads: flat fp32, ACCUMULATED into (caller zeroes).
 * droppath: non-NULL iff forward ran with drop-path (the scales themselves are re-read from the workspace, where
 * forward left a copy).  The call that starts at the top (stage_hi = 4 / block_hi = all blocks) converts dfeats
 * into workspace slots; later partial calls of the same backward reuse them (pass the same dfeats).
 * stage_hi/stage_lo: run backward for stages stage_hi-1 down to stage_lo (4,0 = everything), so the
 * caller can interleave gradient all-reduce of finished stages with the remaining backward. */
int mtus_swin_backward(const mtus_swin_config* cfg, const float* params, const void* params_lp,
                       const float* droppath, void* workspace, const void* const* dfeats, int dfeats_layout,
                       int dfeats_f32, float* grads, int stage_hi, int stage_lo, void* stream);
/* Gradient w.r.t. the input image: dx [B,3,S,S] fp32 NCHW.  For a trainable module in front of the encoder (the reference's
 * input-level TaskPrompt2D, code/models/task_prompt.py:132-143, applied at code/models/multitask_model.py:198-199; timm's
 * trunk gives it through autograd).  Call right after a backward that ran down to block 0 on the same workspace (it reads the
 * patch-embed gradient that backward left there and overwrites the im2col slot). */
int mtus_swin_input_grad(const mtus_swin_config* cfg, const float* params, const void* params_lp, void* workspace,
                         float* dx, void* stream);
/* Same, at block granularity: global block indices count the blocks of all stages in forward order
 * (Swin-B: 0..23); runs blocks block_hi-1 down to block_lo, plus the PatchMerging / patch-embed backward of every
 * stage whose first block is in the range.  Lets the data-parallel wrapper all-reduce finished slices of the flat
 * gradient while the long stage 3 (18 blocks, 2/3 of the parameters) is still running backward. */
int mtus_swin_backward_blocks(const mtus_swin_config* cfg, const float* params, const void* params_lp,
                              const float* droppath, void* workspace, const void* const* dfeats, int dfeats_layout,
                              int dfeats_f32, float* grads, int block_hi, int block_lo, void* stream);

/* The schedule behind mtus_swin_forward / mtus_swin_backward* (everything after the input-dependent prologue) is
 * captured into a CUDA graph per distinct (config, parameter / workspace / gradient addresses, range) and replayed
 * on later calls; keep those buffers alive and at fixed addresses across steps to benefit (MTUS_GRAPHS=0 disables).
 * Counters of the cache since process start. */
void mtus_graph_cache_stats(int64_t* hits, int64_t* misses, int64_t* entries);

typedef struct mtus_fpn_config {
  int batch;
  int in_channels[4]; /* c2..c5 channels (stride 4..32) */
  int sizes[4];       /* spatial size (square) of c2..c5 */
  int pyramid_channels, seg_channels;
  int merge_cat;
  int dtype, backend, training;
} mtus_fpn_config;

int64_t mtus_fpn_param_count(const mtus_fpn_config* cfg);
int64_t mtus_fpn_workspace_bytes(const mtus_fpn_config* cfg);
int mtus_fpn_param_info(const mtus_fpn_config* cfg, int idx, char* name, int64_t* offset, int* rank, int64_t* shape);
/* feats[4]: encoder features c2..c5 (feats_layout 0 NCHW | 1 NHWC); chanscale: [B, out_channels]
 * Dropout2d scales or NULL; out: [B, out_channels, S2, S2].  out_f32 / dout_f32 are bit flags: bit 0 = fp32
 * elements, bit 1 = channels-last (NHWC) memory instead of NCHW. */
int mtus_fpn_forward(const mtus_fpn_config* cfg, const void* const* feats, int feats_layout, int feats_f32,
                     const float* params, const float* chanscale, void* workspace, void* out, int out_f32,
                     void* stream);
/* same with the fused FiLM epilogue (see mtus_fpn_merge_film_fwd); chanscale must then hold gamma[c] (times the Dropout2d
 * scale when active), chanshift = beta [out_channels] or NULL. */
int mtus_fpn_forward_film(const mtus_fpn_config* cfg, const void* const* feats, int feats_layout, int feats_f32,
                          const float* params, const float* chanscale, const float* chanshift, void* workspace, void* out,
                          int out_f32, void* stream);
/* byte offset, inside the workspace, of the NHWC [B,S2,S2,seg_channels] output of tower `level` (0..3 = seg_blocks.0..3, i.e. the p5..p2 towers, in merge order):
 * the merge sources, needed by mtus_film_grad. */
int64_t mtus_fpn_tower_output_offset(const mtus_fpn_config* cfg, int level);
/* dfeats[4]: gradients w.r.t. c2..c5, fully written, in dfeats_layout; grads: flat fp32, ACCUMULATED. */
int mtus_fpn_backward(const mtus_fpn_config* cfg, const void* const* feats, int feats_layout, int feats_f32,
                      const float* params, const float* chanscale, void* workspace, const void* dout, int dout_f32,
                      void* const* dfeats, int dfeats_layout, int dfeats_f32, float* grads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTUS_B200_H_ */
