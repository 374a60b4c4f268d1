"""ctypes binding of ``libmtus_b200.so`` (the C-ABI declared in include/mtus_b200.h).

The product path has NO fallback: if the shared library is missing or a kernel returns a
non-zero status, a RuntimeError is raised (SURVEY §8b "Errors").  ``build()`` compiles the
library in-tree with nvcc for sm_100a (works without a GPU).
"""

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# MTUS_B200_SO: developer knob -- load another build of the SAME library (e.g. the -DMTUS_DIAG_NOATOM diagnostic build)
_SO = os.environ.get("MTUS_B200_SO") or os.path.join(_HERE, "libmtus_b200.so")
_lock = threading.Lock()
_lib = None

F32, BF16 = 0, 1
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TCGEN05 = 0, 1, 2

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", vp), ("lda", i64), ("a_mn_major", i32), ("a_conv", i32),
        ("b", vp), ("ldb", i64), ("b_mn_major", i32), ("b_conv", i32),
        ("conv_h", i32), ("conv_w", i32), ("conv_c", i32),
        ("M", i32), ("N", i32), ("K", i32),
        ("bias", vp), ("act", i32), ("aux", vp), ("ld_aux", i64),
        ("res", vp), ("ld_res", i64), ("res_mode", i32), ("res_h", i32), ("res_w", i32),
        ("rowscale", vp), ("rows_per_sample", i32),
        ("out", vp), ("ld_out", i64), ("out_f32", i32), ("atomic", i32),
        ("split_k", i32), ("dtype", i32), ("backend", i32), ("res_f32", i32), ("out_colsum", vp),
    ]


class SwinConfig(C.Structure):
    _fields_ = [
        ("batch", i32), ("img_size", i32), ("embed_dim", i32),
        ("depths", i32 * 4), ("heads", i32 * 4), ("window", i32),
        ("dtype", i32), ("backend", i32), ("training", i32), ("ln_eps", f32),
    ]


class FpnConfig(C.Structure):
    _fields_ = [
        ("batch", i32), ("in_channels", i32 * 4), ("sizes", i32 * 4),
        ("pyramid_channels", i32), ("seg_channels", i32), ("merge_cat", i32),
        ("dtype", i32), ("backend", i32), ("training", i32),
    ]


# name -> (restype, argtypes); every symbol include/mtus_b200.h declares
_P = C.POINTER
SIGNATURES = {
    "mtus_version": (i32, []),
    "mtus_status_string": (C.c_char_p, [i32]),
    "mtus_launch_count": (i64, []),
    "mtus_layernorm_fwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, f32, i32, vp]),
    "mtus_layernorm_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, vp]),
    "mtus_patch_merge_ln_fwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, vp]),
    "mtus_patch_merge_ln_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_gemm": (i32, [_P(GemmDesc), vp]),
    "mtus_linear_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, vp]),
    "mtus_linear_fwd_stream": (i32, [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, i32, vp]),
    "mtus_linear_dgrad": (i32, [vp, vp, vp, vp, vp, i32, vp, i64, i32, i32, i32, i32, vp]),
    "mtus_linear_fwd_gelu_dact": (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
    "mtus_linear_dgrad_dact": (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
    "mtus_layernorm_fwd_mixed": (i32, [vp, i32, vp, vp, vp, i32, vp, vp, i64, i32, f32, i32, vp]),
    "mtus_layernorm_bwd_mixed": (i32, [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, i64, i32, i32, vp]),
    "mtus_patch_merge_ln_fwd_mixed": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, vp]),
    "mtus_patch_merge_ln_bwd_mixed": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_linear_wgrad": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
    "mtus_cast_f32_to_bf16": (i32, [vp, vp, i64, vp]),
    "mtus_colsum": (i32, [vp, vp, i64, i32, i32, vp]),
    "mtus_scale_rows": (i32, [vp, vp, vp, i32, i64, i32, i32, vp]),
    "mtus_add": (i32, [vp, vp, vp, i64, i32, vp]),
    "mtus_sumsq": (i32, [vp, i64, vp, vp]),
    "mtus_adamw_flat": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, vp, vp]),
    "mtus_pointwise_conv_fwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_pointwise_conv_bwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_adamw_flat_shadow": (i32, [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, vp, vp]),
    "mtus_scale_cast_colsum": (i32, [vp, vp, i32, vp, vp, i64, i32, i32, vp]),
    "mtus_convert": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_nhwc_to_nchw": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_nchw_to_nhwc": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_patch_embed_im2col": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_patch_embed_im2col_u8": (i32, [vp, _P(f32), _P(f32), vp, i32, i32, i32, i32, vp]),
    "mtus_patch_embed_col2im": (i32, [vp, i32, vp, i32, i32, i32, i32, vp]),
    "mtus_window_attn_fwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_window_attn_tc_launch_count": (i64, []),
    "mtus_window_attn_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_upsample_add_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_upsample_add_bwd": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_groupnorm_stats": (i32, [vp, vp, vp, i32, i32, i32, i32, f32, i32, vp]),
    "mtus_groupnorm_relu_fwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_groupnorm_relu_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_groupnorm_act_fwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_groupnorm_act_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_groupnorm_act_fused_fwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, i32, vp]),
    "mtus_batchnorm_stats": (i32, [vp, vp, vp, i64, i32, f32, i32, vp]),
    "mtus_batchnorm_act_fwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp]),
    "mtus_batchnorm_act_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
    "mtus_bilinear2x_fwd": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_bilinear2x_bwd": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "mtus_fpn_merge_fwd": (i32, [_P(vp), i32, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_fpn_merge_film_fwd": (i32, [_P(vp), i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_film_grad": (i32, [vp, i32, _P(vp), i32, i32, vp, vp, vp, i32, i32, i32, i32, vp]),
    "mtus_fpn_merge_bwd": (i32, [vp, i32, i32, vp, _P(vp), i32, i32, i32, i32, i32, i32, vp]),
    "mtus_conv3x3_repack": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "mtus_conv3x3_unpack_grad": (i32, [vp, vp, i32, i32, vp]),
    "mtus_conv3x3_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_conv3x3_dgrad": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_conv3x3_wgrad": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mtus_swin_param_count": (i64, [_P(SwinConfig)]),
    "mtus_swin_workspace_bytes": (i64, [_P(SwinConfig)]),
    "mtus_swin_param_offset": (i64, [_P(SwinConfig), C.c_char_p, _P(i64)]),
    "mtus_swin_param_info": (i32, [_P(SwinConfig), i32, C.c_char_p, _P(i64), _P(i32), _P(i64)]),
    "mtus_swin_feature_offset": (i64, [_P(SwinConfig), i32]),
    "mtus_swin_forward": (i32, [_P(SwinConfig), vp, i32, vp, vp, vp, vp, _P(vp), i32, i32, vp]),
    "mtus_swin_forward_u8": (i32, [_P(SwinConfig), vp, _P(f32), _P(f32), vp, vp, vp, vp, _P(vp), i32, i32, vp]),
    "mtus_swin_backward": (i32, [_P(SwinConfig), vp, vp, vp, vp, _P(vp), i32, i32, vp, i32, i32, vp]),
    "mtus_swin_input_grad": (i32, [_P(SwinConfig), vp, vp, vp, vp, vp]),
    "mtus_graph_cache_stats": (None, [vp, vp, vp]),
    "mtus_swin_backward_blocks": (i32, [_P(SwinConfig), vp, vp, vp, vp, _P(vp), i32, i32, vp, i32, i32, vp]),
    "mtus_fpn_param_count": (i64, [_P(FpnConfig)]),
    "mtus_fpn_workspace_bytes": (i64, [_P(FpnConfig)]),
    "mtus_fpn_param_info": (i32, [_P(FpnConfig), i32, C.c_char_p, _P(i64), _P(i32), _P(i64)]),
    "mtus_fpn_forward": (i32, [_P(FpnConfig), _P(vp), i32, i32, vp, vp, vp, vp, i32, vp]),
    "mtus_fpn_forward_film": (i32, [_P(FpnConfig), _P(vp), i32, i32, vp, vp, vp, vp, vp, i32, vp]),
    "mtus_fpn_tower_output_offset": (i64, [_P(FpnConfig), i32]),
    "mtus_fpn_backward": (i32, [_P(FpnConfig), _P(vp), i32, i32, vp, vp, vp, vp, i32, _P(vp), i32, i32, vp, vp]),
}


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu into libmtus_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    jobs = str(max(1, min(16, os.cpu_count() or 4)))
    proc = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j", jobs], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout[-4000:])
        print(proc.stderr[-4000:])
    if proc.returncode != 0:
        raise RuntimeError("mtus_b200: building libmtus_b200.so failed (see output above)")
    return _SO


def lib():
    """Load (once) and return the ctypes handle; raises if the library is absent -- no CPU fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_SO):
                raise RuntimeError(
                    f"mtus_b200: {_SO} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or `make -C <package>/csrc`). There is no fallback path.")
            h = C.CDLL(_SO)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(h, name)  # AttributeError here = header / library mismatch
                fn.restype, fn.argtypes = res, args
            _lib = h
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().mtus_status_string(status)
        raise RuntimeError(f"mtus_b200 {what} failed: status {status} ({msg.decode() if msg else '?'})")


def ptr(t):
    """Device pointer of a torch tensor (or None) as a ctypes void pointer."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """Current torch stream of ``device`` (default: the current device).  The executors key their side streams and graph
    caches on cudaGetDevice(), so callers that take a tensor's device also make it current (``device_guard``)."""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def device_guard(t):
    """Context manager making ``t.device`` the current CUDA device for the enclosed C-ABI calls."""
    import torch
    return torch.cuda.device(t.device)


def ptr_array(tensors):
    arr = (vp * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
