"""Thin per-kernel Python wrappers over the C-ABI (torch tensors in, torch tensors out).

Used by the parity tests and for experimentation; the models call the whole-encoder / whole-decoder
executors instead.  Every function raises on a non-zero status -- no fallback.
"""

import ctypes as C

import torch

from . import _lib
from ._lib import ptr, stream_ptr, check, lib, F32, BF16


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def layernorm_fwd(x, gamma, beta, eps=1e-5):
    rows, Cc = x.numel() // x.shape[-1], x.shape[-1]
    y = torch.empty_like(x)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    check(lib().mtus_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), rows, Cc, eps, _dt(x), stream_ptr()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None):
    rows, Cc = x.numel() // x.shape[-1], x.shape[-1]
    dx = torch.empty_like(x)
    dg = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dg)
    check(lib().mtus_layernorm_bwd(ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(dg), ptr(db), rows, Cc, _dt(x), stream_ptr()), "layernorm_bwd")
    return dx, dg, db


def patch_merge_ln_fwd(x, gamma, beta, eps=1e-5):
    B, H, W, Cc = x.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    y = torch.empty(B, Ho, Wo, 4 * Cc, dtype=x.dtype, device=x.device)
    mean = torch.empty(B * Ho * Wo, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    check(lib().mtus_patch_merge_ln_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), B, H, W, Cc, eps, _dt(x), stream_ptr()), "patch_merge_ln_fwd")
    return y, mean, rstd


def patch_merge_ln_bwd(dy, x, gamma, mean, rstd, dres=None):
    B, H, W, Cc = x.shape
    dx = torch.empty_like(x)
    dg = torch.zeros(4 * Cc, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dg)
    check(lib().mtus_patch_merge_ln_bwd(ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(dg), ptr(db), B, H, W, Cc, _dt(x), stream_ptr()), "patch_merge_ln_bwd")
    return dx, dg, db


def linear_fwd(x, w, bias=None, gelu=False, res=None, rowscale=None, rows_per_sample=1, backend=0):
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=x.dtype, device=x.device)
    pre = torch.empty_like(y) if gelu else None
    check(lib().mtus_linear_fwd(ptr(x), ptr(w), ptr(bias), ptr(y), ptr(pre), ptr(res), ptr(rowscale), rows_per_sample, M, N, K, _dt(x), backend, stream_ptr()), "linear_fwd")
    return (y, pre) if gelu else y


def linear_dgrad(dy, w, gelu_pre=None, rowscale=None, rows_per_sample=1, backend=0, with_colsum=False):
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty(M, K, dtype=dy.dtype, device=dy.device)
    colsum = torch.zeros(K, dtype=torch.float32, device=dy.device) if with_colsum else None
    check(lib().mtus_linear_dgrad(ptr(dy), ptr(w), ptr(dx), ptr(gelu_pre), ptr(rowscale), rows_per_sample, ptr(colsum), M, N, K, _dt(dy), backend, stream_ptr()), "linear_dgrad")
    return (dx, colsum) if with_colsum else dx


def linear_fwd_gelu_dact(x, w, bias=None, backend=0):
    """(gelu(x w^T + bias), gelu'(x w^T + bias)): the forward of the MLP pair that saves the derivative for the backward."""
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=x.dtype, device=x.device)
    dact = torch.empty_like(y)
    check(lib().mtus_linear_fwd_gelu_dact(ptr(x), ptr(w), ptr(bias), ptr(y), ptr(dact), M, N, K, _dt(x), backend, stream_ptr()), "linear_fwd_gelu_dact")
    return y, dact


def linear_dgrad_dact(dy, w, dact, backend=0, with_colsum=False):
    """dx = (dy w) * dact (+ column sums of dx)."""
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty(M, K, dtype=dy.dtype, device=dy.device)
    colsum = torch.zeros(K, dtype=torch.float32, device=dy.device) if with_colsum else None
    check(lib().mtus_linear_dgrad_dact(ptr(dy), ptr(w), ptr(dx), ptr(dact), ptr(colsum), M, N, K, _dt(dy), backend, stream_ptr()), "linear_dgrad_dact")
    return (dx, colsum) if with_colsum else dx


def linear_fwd_stream(x, w, bias=None, res=None, rowscale=None, rows_per_sample=1, backend=0):
    """y (fp32) = res (fp32) + rowscale * (x w^T + bias): the Linear layers that write the fp32 residual stream."""
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    check(lib().mtus_linear_fwd_stream(ptr(x), ptr(w), ptr(bias), ptr(y), ptr(res), ptr(rowscale), rows_per_sample, M, N, K, _dt(x), backend, stream_ptr()), "linear_fwd_stream")
    return y


def layernorm_fwd_mixed(x, gamma, beta, out_dtype, eps=1e-5):
    rows, Cc = x.numel() // x.shape[-1], x.shape[-1]
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    dt = BF16 if torch.bfloat16 in (x.dtype, out_dtype) else F32
    check(lib().mtus_layernorm_fwd_mixed(ptr(x), int(x.dtype == torch.float32), ptr(gamma), ptr(beta), ptr(y), int(out_dtype == torch.float32),
                                         ptr(mean), ptr(rstd), rows, Cc, eps, dt, stream_ptr()), "layernorm_fwd_mixed")
    return y, mean, rstd


def layernorm_bwd_mixed(dy, x, gamma, mean, rstd, dres=None, lp_dtype=None, rowscale=None, rows_per_sample=1, want_dx=True):
    """Returns (dx fp32 | None, dx_lp | None, lp_colsum | None, dgamma, dbeta)."""
    rows, Cc = x.numel() // x.shape[-1], x.shape[-1]
    dx = torch.empty(x.shape, dtype=torch.float32, device=x.device) if want_dx else None
    lp = torch.empty(x.shape, dtype=lp_dtype, device=x.device) if lp_dtype is not None else None
    cs = torch.zeros(Cc, dtype=torch.float32, device=x.device) if lp is not None else None
    dg = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dg)
    dt = BF16 if torch.bfloat16 in (dy.dtype, x.dtype, lp_dtype) else F32
    check(lib().mtus_layernorm_bwd_mixed(ptr(dy), int(dy.dtype == torch.float32), ptr(x), int(x.dtype == torch.float32), ptr(gamma), ptr(mean),
                                         ptr(rstd), ptr(dres), ptr(dx), ptr(lp), ptr(rowscale), rows_per_sample, ptr(cs), ptr(dg), ptr(db),
                                         rows, Cc, dt, stream_ptr()), "layernorm_bwd_mixed")
    return dx, lp, cs, dg, db


def patch_merge_ln_fwd_mixed(x, gamma, beta, out_dtype, eps=1e-5):
    B, H, W, Cc = x.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    y = torch.empty(B, Ho, Wo, 4 * Cc, dtype=out_dtype, device=x.device)
    mean = torch.empty(B * Ho * Wo, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    check(lib().mtus_patch_merge_ln_fwd_mixed(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), B, H, W, Cc, eps,
                                              BF16 if out_dtype == torch.bfloat16 else F32, stream_ptr()), "patch_merge_ln_fwd_mixed")
    return y, mean, rstd


def patch_merge_ln_bwd_mixed(dy, x, gamma, mean, rstd, dres=None, rowscale=None):
    """dy: [B,Ho,Wo,4C] (operand dtype), x / dres: fp32 [B,H,W,C].  Returns (dx fp32, dx_lp, lp_colsum [C], dgamma, dbeta)."""
    B, H, W, Cc = x.shape
    dx = torch.empty_like(x)
    lp = torch.empty(x.shape, dtype=dy.dtype, device=x.device)
    cs = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    dg = torch.zeros(4 * Cc, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dg)
    if H % 2 or W % 2:     # positions outside the gather never get written
        dx.zero_(); lp.zero_()
    check(lib().mtus_patch_merge_ln_bwd_mixed(ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(lp), ptr(rowscale), H * W,
                                              ptr(cs), ptr(dg), ptr(db), B, H, W, Cc, _dt(dy), stream_ptr()), "patch_merge_ln_bwd_mixed")
    return dx, lp, cs, dg, db


def scale_cast_colsum(g, out_dtype, rowscale=None, rows_per_sample=1):
    rows, Cc = g.shape
    y = torch.empty(rows, Cc, dtype=out_dtype, device=g.device)
    cs = torch.zeros(Cc, dtype=torch.float32, device=g.device)
    check(lib().mtus_scale_cast_colsum(ptr(g), ptr(rowscale), rows_per_sample, ptr(y), ptr(cs), rows, Cc,
                                       BF16 if out_dtype == torch.bfloat16 else F32, stream_ptr()), "scale_cast_colsum")
    return y, cs


def colsum(x, out=None):
    """out[c] += sum_r x[r, c] for x [rows, C] (fp32 or bf16); out fp32 [C] (zeros when not given)."""
    rows, Cc = x.shape
    if out is None:
        out = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    check(lib().mtus_colsum(ptr(x), ptr(out), rows, Cc, _dt(x), stream_ptr()), "colsum")
    return out


def convert(x, out_dtype, transpose=False):
    """[B,R,C] -> [B,R,C] or [B,C,R] with a dtype change (fp32 <-> bf16)."""
    B, R, Cc = x.shape
    y = torch.empty((B, Cc, R) if transpose else (B, R, Cc), dtype=out_dtype, device=x.device)
    dt = BF16 if torch.bfloat16 in (x.dtype, out_dtype) else F32
    check(lib().mtus_convert(ptr(x), ptr(y), B, R, Cc, int(transpose), int(x.dtype == torch.float32), int(out_dtype == torch.float32), dt, stream_ptr()), "convert")
    return y


def linear_wgrad(dy, x, with_bias=True, backend=0):
    M, N = dy.shape
    K = x.shape[1]
    dw = torch.zeros(N, K, dtype=torch.float32, device=dy.device)
    db = torch.zeros(N, dtype=torch.float32, device=dy.device) if with_bias else None
    check(lib().mtus_linear_wgrad(ptr(dy), ptr(x), ptr(dw), ptr(db), M, N, K, _dt(dy), backend, stream_ptr()), "linear_wgrad")
    return dw, db


def window_attn_fwd(qkv, table, qkv_bias, heads, window, shift, return_lse=False):
    B, H, W, C3 = qkv.shape
    Cc = C3 // 3
    out = torch.empty(B, H, W, Cc, dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty(B * H * W, heads, dtype=torch.float32, device=qkv.device)
    check(lib().mtus_window_attn_fwd(ptr(qkv), ptr(table), ptr(qkv_bias), ptr(out), ptr(lse), B, H, W, Cc, heads, window, window, shift, shift, _dt(qkv), stream_ptr()), "window_attn_fwd")
    return (out, lse) if return_lse else out


def window_attn_bwd(dout, qkv, out, table, qkv_bias, heads, window, shift, lse=None, with_colsum=False):
    """Returns (dqkv, dtable, dbias[, dcolsum]); dbias = gradient reaching the qkv bias through padded tokens."""
    B, H, W, C3 = qkv.shape
    Cc = C3 // 3
    if lse is None:   # recompute what forward would have saved
        _, lse = window_attn_fwd(qkv, table, qkv_bias, heads, window, shift, return_lse=True)
    dqkv = torch.empty_like(qkv)
    dtable = torch.zeros_like(table)
    dbias = torch.zeros(C3, dtype=torch.float32, device=qkv.device)
    dcol = torch.zeros(C3, dtype=torch.float32, device=qkv.device) if with_colsum else None
    check(lib().mtus_window_attn_bwd(ptr(dout), ptr(qkv), ptr(out), ptr(lse), ptr(table), ptr(qkv_bias), ptr(dqkv), ptr(dtable), ptr(dbias), ptr(dcol), B, H, W, Cc, heads, window, window, shift, shift, _dt(qkv), stream_ptr()), "window_attn_bwd")
    return (dqkv, dtable, dbias, dcol) if with_colsum else (dqkv, dtable, dbias)


def conv3x3_repack(w, dtype):
    Cout, Cin = w.shape[0], w.shape[1]
    wf = torch.empty(Cout, 9 * Cin, dtype=dtype, device=w.device)
    wd = torch.empty(Cin, 9 * Cout, dtype=dtype, device=w.device)
    check(lib().mtus_conv3x3_repack(ptr(w), ptr(wf), ptr(wd), Cout, Cin, F32 if dtype == torch.float32 else BF16, stream_ptr()), "conv3x3_repack")
    return wf, wd


def conv3x3_fwd(x, wf, backend=0):
    B, H, W, Cin = x.shape
    Cout = wf.shape[0]
    y = torch.empty(B, H, W, Cout, dtype=x.dtype, device=x.device)
    check(lib().mtus_conv3x3_fwd(ptr(x), ptr(wf), ptr(y), B, H, W, Cin, Cout, _dt(x), backend, stream_ptr()), "conv3x3_fwd")
    return y


def conv3x3_dgrad(dy, wd, backend=0):
    B, H, W, Cout = dy.shape
    Cin = wd.shape[0]
    dx = torch.empty(B, H, W, Cin, dtype=dy.dtype, device=dy.device)
    check(lib().mtus_conv3x3_dgrad(ptr(dy), ptr(wd), ptr(dx), B, H, W, Cin, Cout, _dt(dy), backend, stream_ptr()), "conv3x3_dgrad")
    return dx


def conv3x3_wgrad(dy, x, backend=0):
    B, H, W, Cout = dy.shape
    Cin = x.shape[3]
    dwp = torch.zeros(Cout, 9 * Cin, dtype=torch.float32, device=dy.device)
    check(lib().mtus_conv3x3_wgrad(ptr(dy), ptr(x), ptr(dwp), B, H, W, Cin, Cout, _dt(dy), backend, stream_ptr()), "conv3x3_wgrad")
    dw = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float32, device=dy.device)
    check(lib().mtus_conv3x3_unpack_grad(ptr(dwp), ptr(dw), Cout, Cin, stream_ptr()), "conv3x3_unpack_grad")
    return dw


def groupnorm_relu_fwd(x, gamma, beta, groups=32, eps=1e-5):
    B, H, W, Cc = x.shape
    mean = torch.empty(B * groups, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    y = torch.empty_like(x)
    check(lib().mtus_groupnorm_act_fused_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), B, H * W, Cc, groups, eps, 0, _dt(x), stream_ptr()), "groupnorm_act_fused_fwd")
    return y, mean, rstd


def groupnorm_relu_fwd_two_pass(x, gamma, beta, groups=32, eps=1e-5):
    """The separate statistics + normalise kernels (what the fused call falls back to for maps that do not fit in shared memory)."""
    B, H, W, Cc = x.shape
    mean = torch.empty(B * groups, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    check(lib().mtus_groupnorm_stats(ptr(x), ptr(mean), ptr(rstd), B, H * W, Cc, groups, eps, _dt(x), stream_ptr()), "groupnorm_stats")
    y = torch.empty_like(x)
    check(lib().mtus_groupnorm_relu_fwd(ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(y), B, H * W, Cc, groups, _dt(x), stream_ptr()), "groupnorm_relu_fwd")
    return y, mean, rstd


def groupnorm_relu_bwd(dy, x, y, mean, rstd, gamma, groups=32, beta=None):
    """beta given: the ReLU gate is recomputed from x (y is not read); beta None: the gate comes from the saved output y."""
    B, H, W, Cc = x.shape
    dx = torch.empty_like(x)
    dg = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dg)
    ws = torch.empty(2 * B * groups, dtype=torch.float32, device=x.device)
    if beta is not None:
        check(lib().mtus_groupnorm_act_bwd(ptr(dy), ptr(x), None, ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(dx), ptr(dg), ptr(db), ptr(ws), B, H * W, Cc, groups, 0, _dt(x), stream_ptr()), "groupnorm_act_bwd")
    else:
        check(lib().mtus_groupnorm_relu_bwd(ptr(dy), ptr(x), ptr(y), ptr(mean), ptr(rstd), ptr(gamma), ptr(dx), ptr(dg), ptr(db), ptr(ws), B, H * W, Cc, groups, _dt(x), stream_ptr()), "groupnorm_relu_bwd")
    return dx, dg, db


def groupnorm_silu_fwd(x, gamma, beta, groups=32, eps=1e-5):
    """y = silu(group_norm(x)) on an NHWC tensor [B,H,W,C]; returns (y, mean, rstd)."""
    B, H, W, Cc = x.shape
    mean = torch.empty(B * groups, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    y = torch.empty_like(x)
    check(lib().mtus_groupnorm_act_fused_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), B, H * W, Cc, groups, eps, 1, _dt(x), stream_ptr()), "groupnorm_act_fused_fwd")
    return y, mean, rstd


def groupnorm_silu_bwd(dy, x, mean, rstd, gamma, beta, groups=32):
    B, H, W, Cc = x.shape
    dx = torch.empty_like(x)
    dg = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dg)
    ws = torch.empty(2 * B * groups, dtype=torch.float32, device=x.device)
    check(lib().mtus_groupnorm_act_bwd(ptr(dy), ptr(x), None, ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(dx), ptr(dg), ptr(db), ptr(ws), B, H * W, Cc, groups, 1, _dt(x), stream_ptr()), "groupnorm_act_bwd")
    return dx, dg, db


def bilinear2x_fwd(x):
    B, H, W, Cc = x.shape
    y = torch.empty(B, 2 * H, 2 * W, Cc, dtype=x.dtype, device=x.device)
    check(lib().mtus_bilinear2x_fwd(ptr(x), ptr(y), B, H, W, Cc, _dt(x), stream_ptr()), "bilinear2x_fwd")
    return y


def bilinear2x_bwd(dy):
    B, H2, W2, Cc = dy.shape
    dx = torch.empty(B, H2 // 2, W2 // 2, Cc, dtype=dy.dtype, device=dy.device)
    check(lib().mtus_bilinear2x_bwd(ptr(dy), ptr(dx), B, H2 // 2, W2 // 2, Cc, _dt(dy), stream_ptr()), "bilinear2x_bwd")
    return dx


def upsample_add_fwd(skip, top):
    B, H, W, Cc = skip.shape
    y = torch.empty_like(skip)
    check(lib().mtus_upsample_add_fwd(ptr(skip), ptr(top), ptr(y), B, H, W, Cc, _dt(skip), stream_ptr()), "upsample_add_fwd")
    return y


def upsample_add_bwd(dy):
    B, H, W, Cc = dy.shape
    dtop = torch.empty(B, H // 2, W // 2, Cc, dtype=dy.dtype, device=dy.device)
    check(lib().mtus_upsample_add_bwd(ptr(dy), ptr(dtop), 0, B, H, W, Cc, _dt(dy), stream_ptr()), "upsample_add_bwd")
    return dtop


def nhwc_to_nchw(x):
    B, H, W, Cc = x.shape
    y = torch.empty(B, Cc, H, W, dtype=x.dtype, device=x.device)
    check(lib().mtus_nhwc_to_nchw(ptr(x), ptr(y), B, H * W, Cc, _dt(x), 0, stream_ptr()), "nhwc_to_nchw")
    return y


def nchw_to_nhwc(x):
    B, Cc, H, W = x.shape
    y = torch.empty(B, H, W, Cc, dtype=x.dtype, device=x.device)
    check(lib().mtus_nchw_to_nhwc(ptr(x), ptr(y), B, H * W, Cc, _dt(x), 0, stream_ptr()), "nchw_to_nhwc")
    return y


def batchnorm_stats(x, eps=1e-5):
    """Per-channel batch mean and 1/sqrt(biased var + eps) of NHWC rows x [..., C]."""
    Cc = x.shape[-1]
    M = x.numel() // Cc
    mean = torch.empty(Cc, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    check(lib().mtus_batchnorm_stats(ptr(x), ptr(mean), ptr(rstd), M, Cc, eps, _dt(x), stream_ptr()), "batchnorm_stats")
    return mean, rstd


def batchnorm_relu_fwd(x, mean, rstd, gamma, beta):
    Cc = x.shape[-1]
    y = torch.empty_like(x)
    check(lib().mtus_batchnorm_act_fwd(ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(y), x.numel() // Cc, Cc, 0, _dt(x), stream_ptr()), "batchnorm_act_fwd")
    return y


def batchnorm_relu_bwd(dy, x, y, mean, rstd, gamma, training=True):
    Cc = x.shape[-1]
    dx = torch.empty_like(x)
    dg = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dg)
    ws = torch.empty(2 * Cc, dtype=torch.float32, device=x.device)
    check(lib().mtus_batchnorm_act_bwd(ptr(dy), ptr(x), ptr(y), ptr(mean), ptr(rstd), ptr(gamma), None, ptr(dx), ptr(dg), ptr(db), ptr(ws),
                                       x.numel() // Cc, Cc, 0, int(training), _dt(x), stream_ptr()), "batchnorm_act_bwd")
    return dx, dg, db


def pointwise_conv_fwd(x, w, bias=None):
    """x [B,H,W,K] channels-last rows, w [N,K] fp32, bias [N] -> y [B,N,H,W] fp32 (N <= 8; the head tails)."""
    B, H, W, K = x.shape
    N = w.shape[0]
    y = torch.empty(B, N, H, W, dtype=torch.float32, device=x.device)
    check(lib().mtus_pointwise_conv_fwd(ptr(x), ptr(w), ptr(bias), ptr(y), B, H * W, K, N, _dt(x), stream_ptr()), "pointwise_conv_fwd")
    return y


def pointwise_conv_bwd(dy, x, w, need_dx=True):
    """dy [B,N,H,W] fp32 contiguous -> (dx [B,H,W,K] | None, dw [N,K] fp32, dbias [N] fp32)."""
    B, H, W, K = x.shape
    N = w.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.zeros(N, K, dtype=torch.float32, device=x.device)
    db = torch.zeros(N, dtype=torch.float32, device=x.device)
    check(lib().mtus_pointwise_conv_bwd(ptr(dy), ptr(x), ptr(w), ptr(dx), ptr(dw), ptr(db), B, H * W, K, N, _dt(x), stream_ptr()), "pointwise_conv_bwd")
    return dx, dw, db
