"""Host-free kernel timing for the roofline numbers of ``bench.py`` and ``tools/roofline_table.py``.

Every kernel is timed as a CUDA GRAPH of back-to-back launches over a ring of pre-allocated operand sets whose total
size exceeds the 126 MB L2 several times: the timed region is one ``cudaGraphLaunch`` between two events on the launching
stream -- no Python, no allocation, no fill kernels inside it (round-1 timed Python wrappers that allocated their outputs,
which made the number host-enqueue-bound and irreproducible between boxes).  Outputs that a kernel accumulates into
(split-K weight gradients, column sums) are simply accumulated further: the values are irrelevant, the traffic is the same.

``achieved`` = ALGORITHMIC flops / bytes per launch (SURVEY §8d definitions, restated next to each entry) divided by the
average launch duration.
"""

import ctypes as C
import os

import torch

from . import _lib
from ._lib import ptr, F32, BF16

L2_BYTES = 126e6


def graph_time(sets, run, launches=None, min_ring_bytes=4 * L2_BYTES, bytes_per_set=None, reps=3):
    """Average seconds per launch of ``run(set)`` replayed from a CUDA graph over the ring ``sets``."""
    n = len(sets)
    launches = launches or max(2 * n, 24)
    launches = ((launches + n - 1) // n) * n
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for i in range(min(n, 4)):
            run(sets[i])                       # eager warm-up: tensor maps, attributes, first-touch
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(launches):
                run(sets[i % n])
        g.replay()
        st.synchronize()
        best = None
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            g.replay()
            b.record(st)
            st.synchronize()
            t = a.elapsed_time(b) * 1e-3 / launches
            best = t if best is None else min(best, t)
    torch.cuda.current_stream().wait_stream(st)
    del g
    return best


def _ring(make_set, bytes_per_set, max_sets=24):
    n = int(max(3, min(max_sets, 4 * L2_BYTES // max(bytes_per_set, 1) + 1)))
    return [make_set() for _ in range(n)]


def _sp():
    return _lib.stream_ptr()


# ---------------------------------------------------------------------------------------------------------------------
# tensor-core GEMMs (gemm_tc2_kernel): the Linear layers of one Swin block, forward / dgrad / wgrad
# ---------------------------------------------------------------------------------------------------------------------
def linear_cases(M, Cc):
    """(name, N, K, kind) of the four Linears of a block of width Cc over M tokens."""
    return [("qkv", 3 * Cc, Cc, "bias"), ("proj", Cc, Cc, "stream"), ("fc1", 4 * Cc, Cc, "gelu"), ("fc2", Cc, 4 * Cc, "stream")]


def time_linear(M, N, K, kind, direction, device):
    """Returns (seconds per launch, algorithmic flops, algorithmic bytes).

    forward  bias:   y[M,N] bf16 = x w^T + b                      bytes (M K + N K + M N) 2
             gelu:   a, h[M,N] bf16 = gelu, gelu'(x w^T + b)      bytes (M K + N K + 2 M N) 2   (the executor's MLP pair: the
                     derivative is saved instead of the pre-activation, mtus_linear_fwd_gelu_dact / mtus_linear_dgrad_dact)
             stream: y[M,N] fp32 = res fp32 + x w^T + b           bytes (M K + N K) 2 + 2 M N 4
    dgrad    dx[M,K] bf16 = dy w (x h for fc1's producer)         bytes (M N + N K + M K) 2 (+ M K 2 for h)
    wgrad    dw[N,K] fp32 += dy^T x (split-K, TMA reduce-add)     bytes (M N + M K) 2 + N K 4
    """
    L = _lib.lib()
    be = _lib.BACKEND_TCGEN05
    bf = torch.bfloat16
    flops = 2.0 * M * N * K
    if direction == "fwd":
        def make():
            x = (torch.randn(M, K, device=device) * 0.5).to(bf)
            w = (torch.randn(N, K, device=device) * 0.02).to(bf)
            b = torch.zeros(N, device=device)
            if kind == "stream":
                return x, w, b, torch.randn(M, N, device=device), torch.empty(M, N, device=device)
            if kind == "gelu":
                return x, w, b, torch.empty(M, N, device=device, dtype=bf), torch.empty(M, N, device=device, dtype=bf)
            return x, w, b, None, torch.empty(M, N, device=device, dtype=bf)
        by = (M * K + N * K) * 2 + M * N * (8 if kind == "stream" else (4 if kind == "gelu" else 2))

        def run(s):
            x, w, b, aux, y = s
            if kind == "stream":
                _lib.check(L.mtus_linear_fwd_stream(ptr(x), ptr(w), ptr(b), ptr(y), ptr(aux), None, 1, M, N, K, BF16, be, _sp()), "fwd_stream")
            elif kind == "gelu":
                _lib.check(L.mtus_linear_fwd_gelu_dact(ptr(x), ptr(w), ptr(b), ptr(y), ptr(aux), M, N, K, BF16, be, _sp()), "fwd_gelu")
            else:
                _lib.check(L.mtus_linear_fwd(ptr(x), ptr(w), ptr(b), ptr(y), None, None, None, 1, M, N, K, BF16, be, _sp()), "fwd")
    elif direction == "dgrad":
        # dgrad of fc2 multiplies by gelu'(h) (its output feeds fc1's pre-activation): kind "stream" with N = C, K = 4C
        with_gelu = (kind == "stream" and K == 4 * N)

        def make():
            dy = (torch.randn(M, N, device=device) * 0.5).to(bf)
            w = (torch.randn(N, K, device=device) * 0.02).to(bf)
            h = torch.randn(M, K, device=device).to(bf) if with_gelu else None
            cs = torch.zeros(K, device=device) if (with_gelu and not os.environ.get("MTUS_KBENCH_NO_COLSUM")) else None
            return dy, w, h, cs, torch.empty(M, K, device=device, dtype=bf)
        by = (M * N + N * K + M * K) * 2 + (M * K * 2 if with_gelu else 0)

        def run(s):
            dy, w, h, cs, dx = s
            if with_gelu:
                _lib.check(L.mtus_linear_dgrad_dact(ptr(dy), ptr(w), ptr(dx), ptr(h), ptr(cs), M, N, K, BF16, be, _sp()), "dgrad_dact")
            else:
                _lib.check(L.mtus_linear_dgrad(ptr(dy), ptr(w), ptr(dx), None, None, 1, None, M, N, K, BF16, be, _sp()), "dgrad")
    else:
        def make():
            dy = (torch.randn(M, N, device=device) * 0.5).to(bf)
            x = (torch.randn(M, K, device=device) * 0.5).to(bf)
            return dy, x, torch.zeros(N, K, device=device)
        by = (M * N + M * K) * 2 + N * K * 4

        def run(s):
            dy, x, dw = s
            _lib.check(L.mtus_linear_wgrad(ptr(dy), ptr(x), ptr(dw), None, M, N, K, BF16, be, _sp()), "wgrad")
    sets = _ring(make, by)
    t = graph_time(sets, run)
    del sets
    return t, flops, float(by)


def swin_b_stage_shapes(batch=32):
    """(stage, M tokens, C) of Swin-B at 224x224."""
    return [(1, batch * 3136, 128), (2, batch * 784, 256), (3, batch * 196, 512), (4, batch * 49, 1024)]


# ---------------------------------------------------------------------------------------------------------------------
# bandwidth-bound kernels (SURVEY §8d algorithmic bytes; s = 2 bytes for bf16 operands, the residual stream is fp32)
# ---------------------------------------------------------------------------------------------------------------------
def time_layernorm_fwd(rows, Cc, device):
    """lnv2_fwd: fp32 stream in, bf16 operand out (+ mean / rstd): bytes rows C (4 + 2)."""
    L = _lib.lib()
    g, b = torch.ones(Cc, device=device), torch.zeros(Cc, device=device)

    def make():
        return (torch.randn(rows, Cc, device=device), torch.empty(rows, Cc, device=device, dtype=torch.bfloat16),
                torch.empty(rows, device=device), torch.empty(rows, device=device))

    def run(s):
        x, y, mean, rstd = s
        _lib.check(L.mtus_layernorm_fwd_mixed(ptr(x), 1, ptr(g), ptr(b), ptr(y), 0, ptr(mean), ptr(rstd), rows, Cc, 1e-5, BF16, _sp()), "ln_fwd")
    by = rows * Cc * 6.0
    return graph_time(_ring(make, by), run), by


def time_layernorm_bwd(rows, Cc, device):
    """lnv2_bwd: bf16 dy, fp32 x, fp32 stream gradient in and out, bf16 operand copy out: bytes rows C (2 + 4 + 4 + 4 + 2)."""
    L = _lib.lib()
    g = torch.ones(Cc, device=device)

    def make():
        x = torch.randn(rows, Cc, device=device)
        return (torch.randn(rows, Cc, device=device).bfloat16(), x, x.mean(1).contiguous(), (x.var(1, unbiased=False) + 1e-5).rsqrt().contiguous(),
                torch.randn(rows, Cc, device=device), torch.empty(rows, Cc, device=device), torch.empty(rows, Cc, device=device, dtype=torch.bfloat16),
                torch.zeros(Cc, device=device), torch.zeros(Cc, device=device), torch.zeros(Cc, device=device))

    def run(s):
        dy, x, mean, rstd, dres, dx, lp, cs, dg, db = s
        _lib.check(L.mtus_layernorm_bwd_mixed(ptr(dy), 0, ptr(x), 1, ptr(g), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(lp), None, 1,
                                              ptr(cs), ptr(dg), ptr(db), rows, Cc, BF16, _sp()), "ln_bwd")
    by = rows * Cc * 16.0
    return graph_time(_ring(make, by), run), by


def time_patch_merge_ln(B, H, Cc, device, backward=False):
    """PatchMerging gather + LayerNorm(4C).  fwd: fp32 [B,H,H,C] in, bf16 [B,H/2,H/2,4C] out: bytes B H^2 C (4 + 2);
    bwd: bf16 dy in, fp32 x in, fp32 dres in, fp32 dx out, bf16 copy out: bytes B H^2 C (2 + 4 + 4 + 4 + 2)."""
    L = _lib.lib()
    g, b = torch.ones(4 * Cc, device=device), torch.zeros(4 * Cc, device=device)
    Ho = H // 2
    bf = torch.bfloat16
    if not backward:
        def make():
            return (torch.randn(B, H, H, Cc, device=device), torch.empty(B, Ho, Ho, 4 * Cc, device=device, dtype=bf),
                    torch.empty(B * Ho * Ho, device=device), torch.empty(B * Ho * Ho, device=device))

        def run(s):
            x, y, mean, rstd = s
            _lib.check(L.mtus_patch_merge_ln_fwd_mixed(ptr(x), ptr(g), ptr(b), ptr(y), ptr(mean), ptr(rstd), B, H, H, Cc, 1e-5, BF16, _sp()), "merge_fwd")
        by = B * H * H * Cc * 6.0
    else:
        def make():
            x = torch.randn(B, H, H, Cc, device=device)
            xg = x.reshape(B, Ho, 2, Ho, 2, Cc).permute(0, 1, 3, 4, 2, 5).flatten(3)
            mean = xg.mean(-1).flatten().contiguous()
            rstd = (xg.var(-1, unbiased=False) + 1e-5).rsqrt().flatten().contiguous()
            return (torch.randn(B, Ho, Ho, 4 * Cc, device=device).to(bf), x, mean, rstd, torch.randn(B, H, H, Cc, device=device),
                    torch.empty(B, H, H, Cc, device=device), torch.empty(B, H, H, Cc, device=device, dtype=bf),
                    torch.zeros(Cc, device=device), torch.zeros(4 * Cc, device=device), torch.zeros(4 * Cc, device=device))

        def run(s):
            dy, x, mean, rstd, dres, dx, lp, cs, dg, db = s
            _lib.check(L.mtus_patch_merge_ln_bwd_mixed(ptr(dy), ptr(x), ptr(g), ptr(mean), ptr(rstd), ptr(dres), ptr(dx), ptr(lp), None, H * H,
                                                       ptr(cs), ptr(dg), ptr(db), B, H, H, Cc, BF16, _sp()), "merge_bwd")
        by = B * H * H * Cc * 16.0
    return graph_time(_ring(make, by), run), by


def time_window_attn(B, H, Cc, heads, win, shift, device, backward=False):
    """Stand-alone window attention: fwd reads q, k, v and writes o: 4 T C s bytes; bwd reads q, k, v, o, dO and writes
    dq, dk, dv: 8 T C s bytes.  Attention-only flops 4 N^2 d per (window, head) forward, x2.5 backward (5 contractions)."""
    L = _lib.lib()
    bf = torch.bfloat16
    tab = torch.randn((2 * win - 1) ** 2, heads, device=device) * 0.1
    bias = torch.zeros(3 * Cc, device=device)
    T = B * H * H
    nwin = B * ((H + win - 1) // win) ** 2
    flops = 4.0 * (win * win) ** 2 * 32 * nwin * heads

    def fwd(qkv, out, lse):
        _lib.check(L.mtus_window_attn_fwd(ptr(qkv), ptr(tab), ptr(bias), ptr(out), ptr(lse), B, H, H, Cc, heads, win, win, shift, shift, BF16, _sp()), "attn_fwd")
    if not backward:
        def make():
            return ((torch.randn(B, H, H, 3 * Cc, device=device)).to(bf), torch.empty(B, H, H, Cc, device=device, dtype=bf),
                    torch.empty(T, heads, device=device))

        def run(s):
            fwd(*s)
        by = 4.0 * T * Cc * 2
    else:
        def make():
            qkv = (torch.randn(B, H, H, 3 * Cc, device=device)).to(bf)
            out, lse = torch.empty(B, H, H, Cc, device=device, dtype=bf), torch.empty(T, heads, device=device)
            fwd(qkv, out, lse)
            return (qkv, out, lse, torch.randn(B, H, H, Cc, device=device).to(bf), torch.empty_like(qkv), torch.zeros_like(tab),
                    torch.zeros(3 * Cc, device=device), torch.zeros(3 * Cc, device=device))

        def run(s):
            qkv, out, lse, dout, dqkv, dtab, dbias, dcol = s
            _lib.check(L.mtus_window_attn_bwd(ptr(dout), ptr(qkv), ptr(out), ptr(lse), ptr(tab), ptr(bias), ptr(dqkv), ptr(dtab), ptr(dbias), ptr(dcol),
                                              B, H, H, Cc, heads, win, win, shift, shift, BF16, _sp()), "attn_bwd")
        by = 8.0 * T * Cc * 2
        flops *= 2.5
    return graph_time(_ring(make, by), run), by, flops


def time_fpn_upadd(B, H, Cc, device):
    """FPN top-down: y = skip + nearest_up(top): bytes (1/4 + 1 + 1) B H^2 C s."""
    L = _lib.lib()
    bf = torch.bfloat16

    def make():
        return (torch.randn(B, H, H, Cc, device=device).to(bf), torch.randn(B, H // 2, H // 2, Cc, device=device).to(bf),
                torch.empty(B, H, H, Cc, device=device, dtype=bf))

    def run(s):
        skip, top, y = s
        _lib.check(L.mtus_upsample_add_fwd(ptr(skip), ptr(top), ptr(y), B, H, H, Cc, BF16, _sp()), "upadd")
    by = 2.25 * B * H * H * Cc * 2
    return graph_time(_ring(make, by), run), by


def time_fpn_merge(B, H, Cc, nsrc, device):
    """MergeBlock cat + Dropout2d scale, NHWC out: bytes 2 B H^2 (nsrc C) s."""
    L = _lib.lib()
    bf = torch.bfloat16

    def make():
        srcs = [torch.randn(B, H * H, Cc, device=device).to(bf) for _ in range(nsrc)]
        return srcs, _lib.ptr_array(srcs), torch.ones(B, nsrc * Cc, device=device), torch.empty(B, H * H, nsrc * Cc, device=device, dtype=bf)

    def run(s):
        srcs, arr, scale, out = s
        _lib.check(L.mtus_fpn_merge_fwd(arr, nsrc, 1, ptr(scale), ptr(out), B, H * H, Cc, BF16, 0, 1, _sp()), "merge")
    by = 2.0 * B * H * H * nsrc * Cc * 2
    return graph_time(_ring(make, by), run), by


def time_groupnorm_relu(B, H, Cc, device, fused=True):
    """GroupNorm(32) statistics + normalise + ReLU.  fused: ONE cluster kernel (a sample is read once into shared memory, group
    sums cross CTAs through distributed shared memory) = the 2 B H^2 C s of SURVEY 8d (one read + one write).  fused=False: the
    two-pass statistics kernels + the normalise kernel (3 reads + 1 write), the fallback for maps that do not fit."""
    L = _lib.lib()
    bf = torch.bfloat16
    g, b = torch.ones(Cc, device=device), torch.zeros(Cc, device=device)

    def make():
        return (torch.randn(B, H, H, Cc, device=device).to(bf), torch.empty(B * 32, device=device), torch.empty(B * 32, device=device),
                torch.empty(B, H, H, Cc, device=device, dtype=bf))

    def run(s):
        x, mean, rstd, y = s
        if fused:
            _lib.check(L.mtus_groupnorm_act_fused_fwd(ptr(x), ptr(g), ptr(b), ptr(y), ptr(mean), ptr(rstd), B, H * H, Cc, 32, 1e-5, 0, BF16, _sp()), "gn_fused")
            return
        _lib.check(L.mtus_groupnorm_stats(ptr(x), ptr(mean), ptr(rstd), B, H * H, Cc, 32, 1e-5, BF16, _sp()), "gn_stats")
        _lib.check(L.mtus_groupnorm_relu_fwd(ptr(x), ptr(mean), ptr(rstd), ptr(g), ptr(b), ptr(y), B, H * H, Cc, 32, BF16, _sp()), "gn_relu")
    by = 2.0 * B * H * H * Cc * 2
    sets = _ring(make, 1.5 * by)
    return graph_time(sets, run), by


def time_bilinear(B, H, Cc, device):
    """bilinear x2 (align_corners=True): bytes (1/4 + 1) B (2H)^2 C s."""
    L = _lib.lib()
    bf = torch.bfloat16

    def make():
        return torch.randn(B, H, H, Cc, device=device).to(bf), torch.empty(B, 2 * H, 2 * H, Cc, device=device, dtype=bf)

    def run(s):
        x, y = s
        _lib.check(L.mtus_bilinear2x_fwd(ptr(x), ptr(y), B, H, H, Cc, BF16, _sp()), "bilinear")
    by = 1.25 * B * 4 * H * H * Cc * 2
    return graph_time(_ring(make, by), run), by


def time_colsum(rows, Cc, device):
    """out[c] += sum_r x[r, c] (bias gradient of the FPN lateral convolutions): bytes rows C s."""
    L = _lib.lib()

    def make():
        return torch.randn(rows, Cc, device=device).to(torch.bfloat16), torch.zeros(Cc, device=device)

    def run(s):
        x, out = s
        _lib.check(L.mtus_colsum(ptr(x), ptr(out), rows, Cc, BF16, _sp()), "colsum")
    by = rows * Cc * 2.0
    return graph_time(_ring(make, by), run), by


def time_groupnorm_bwd(B, H, Cc, device, act=0):
    """GroupNorm(32) + activation backward (gate recomputed from x): reads dy and x, writes dx: bytes 3 B H^2 C s."""
    L = _lib.lib()
    bf = torch.bfloat16
    g, b = torch.ones(Cc, device=device), torch.zeros(Cc, device=device)

    def make():
        x = torch.randn(B, H, H, Cc, device=device)
        xg = x.reshape(B, H * H, 32, Cc // 32)
        mean = xg.mean((1, 3)).contiguous()
        rstd = (xg.var((1, 3), unbiased=False) + 1e-5).rsqrt().contiguous()
        return (torch.randn(B, H, H, Cc, device=device).to(bf), x.to(bf), mean, rstd, torch.empty(B, H, H, Cc, device=device, dtype=bf),
                torch.zeros(Cc, device=device), torch.zeros(Cc, device=device), torch.zeros(2 * B * 32, device=device))

    def run(s):
        dy, x, mean, rstd, dx, dg, db, ws = s
        _lib.check(L.mtus_groupnorm_act_bwd(ptr(dy), ptr(x), None, ptr(mean), ptr(rstd), ptr(g), ptr(b), ptr(dx), ptr(dg), ptr(db), ptr(ws),
                                            B, H * H, Cc, 32, act, BF16, _sp()), "gn_bwd")
    by = 3.0 * B * H * H * Cc * 2
    return graph_time(_ring(make, by), run), by
