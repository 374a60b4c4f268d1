"""GPU bring-up diagnostics (test infrastructure): runs every kernel group against a PyTorch reference in
its own subprocess (a faulting kernel must not poison the rest) and prints error tables.

    python tests/gpu_diag.py            # all groups
    python tests/gpu_diag.py gemm_tc    # one group, in-process
"""

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["elementwise", "gemm_simt", "gemm_tc", "attention", "fpn_ops", "model_fp32", "model_bf16_simt", "model_bf16_tc"]


def err(a, b):
    a, b = a.float(), b.float()
    d = (a - b).abs().max().item()
    rel = d / (b.abs().max().item() + 1e-12)
    return d, rel


def report(name, got, ref, tol):
    import torch
    d, rel = err(got, ref)
    bad = not (rel <= tol) or not torch.isfinite(got.float()).all().item()
    print(f"  {'FAIL' if bad else 'ok  '} {name:58s} max|d|={d:.3e} rel={rel:.3e} tol={tol:.0e}", flush=True)
    return not bad


def g_elementwise():
    import torch
    import torch.nn.functional as F
    import mtus_b200 as m
    from mtus_b200 import ops
    ok = True
    dev = "cuda"
    for dt, tol in ((torch.float32, 2e-5), (torch.bfloat16, 2e-2)):
        for rows, Cc in ((1000, 96), (3137, 128), (777, 512), (300, 768), (200, 1024), (100, 1536), (64, 3072)):
            x = torch.randn(rows, Cc, device=dev).to(dt)
            g = torch.randn(Cc, device=dev) * 0.5 + 1
            b = torch.randn(Cc, device=dev) * 0.1
            y, mean, rstd = ops.layernorm_fwd(x, g, b)
            xr = x.float().requires_grad_(True)
            gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
            yr = F.layer_norm(xr, (Cc,), gr, br, 1e-5)
            ok &= report(f"ln_fwd {dt} [{rows},{Cc}]", y, yr, tol)
            dy = torch.randn(rows, Cc, device=dev).to(dt)
            dres = torch.randn(rows, Cc, device=dev).to(dt)
            dx, dg, db = ops.layernorm_bwd(dy, x, g, mean, rstd, dres)
            yr.backward(dy.float())
            ok &= report(f"ln_bwd dx {dt} [{rows},{Cc}]", dx, xr.grad + dres.float(), tol)
            ok &= report(f"ln_bwd dgamma {dt} [{rows},{Cc}]", dg, gr.grad, max(tol, 2e-4))
            ok &= report(f"ln_bwd dbeta {dt} [{rows},{Cc}]", db, br.grad, max(tol, 2e-4))
        for (B, H, W, Cc) in ((2, 8, 8, 32), (3, 7, 7, 64), (2, 14, 14, 128), (1, 5, 9, 768)):
            x = torch.randn(B, H, W, Cc, device=dev).to(dt)
            g = torch.randn(4 * Cc, device=dev) * 0.5 + 1
            b = torch.randn(4 * Cc, device=dev) * 0.1
            y, mean, rstd = ops.patch_merge_ln_fwd(x, g, b)
            xr = x.float().requires_grad_(True)
            xp = F.pad(xr, (0, 0, 0, W % 2, 0, H % 2))
            Hp, Wp = xp.shape[1], xp.shape[2]
            xg = xp.reshape(B, Hp // 2, 2, Wp // 2, 2, Cc).permute(0, 1, 3, 4, 2, 5).flatten(3)
            yr = F.layer_norm(xg, (4 * Cc,), g, b, 1e-5)
            ok &= report(f"merge_ln_fwd {dt} {B,H,W,Cc}", y, yr, tol)
            dy = torch.randn_like(yr).to(dt)
            dx, dg, db = ops.patch_merge_ln_bwd(dy, x, g, mean, rstd)
            yr.backward(dy.float())
            ok &= report(f"merge_ln_bwd {dt} {B,H,W,Cc}", dx, xr.grad, tol)
        # ---- mixed-precision LayerNorm family (fp32 residual stream) ----
        for rows, Cc in ((1000, 96), (3137, 128), (300, 768), (200, 1024), (64, 3072)):
            xs = torch.randn(rows, Cc, device=dev) * 2 + 0.5                       # fp32 stream
            g = torch.randn(Cc, device=dev) * 0.5 + 1
            b = torch.randn(Cc, device=dev) * 0.1
            y, mean, rstd = ops.layernorm_fwd_mixed(xs, g, b, dt)
            xr = xs.clone().requires_grad_(True)
            gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
            yr = F.layer_norm(xr, (Cc,), gr, br, 1e-5)
            ok &= report(f"lnx_fwd f32->{dt} [{rows},{Cc}]", y, yr, tol)
            dy = torch.randn(rows, Cc, device=dev).to(dt)
            dres = torch.randn(rows, Cc, device=dev)
            rs = torch.rand(5, device=dev) + 0.5
            rps = (rows + 4) // 5
            dx, lp, cs, dg, db = ops.layernorm_bwd_mixed(dy, xs, g, mean, rstd, dres, lp_dtype=dt, rowscale=rs, rows_per_sample=rps)
            yr.backward(dy.float())
            ref_dx = xr.grad + dres
            ok &= report(f"lnx_bwd dx (fp32 stream)", dx, ref_dx, 2e-5)
            sc = rs[torch.arange(rows, device=dev) // rps][:, None]
            ok &= report(f"lnx_bwd dx_lp (scaled operand copy)", lp, ref_dx * sc, tol)
            ok &= report(f"lnx_bwd lp_colsum", cs, lp.float().sum(0), 2e-4)
            ok &= report(f"lnx_bwd dgamma", dg, gr.grad, max(tol, 2e-4))
            ok &= report(f"lnx_bwd dbeta", db, br.grad, max(tol, 2e-4))
            # patch-embed form: low-precision input, fp32 output; backward with dy = fp32 stream, only the operand copy out
            xl = (torch.randn(rows, Cc, device=dev) * 2).to(dt)
            y2, mean2, rstd2 = ops.layernorm_fwd_mixed(xl, g, b, torch.float32)
            xr2 = xl.float().requires_grad_(True)
            yr2 = F.layer_norm(xr2, (Cc,), g, b, 1e-5)
            ok &= report(f"lnx_fwd {dt}->f32", y2, yr2, 2e-5)
            dyf = torch.randn(rows, Cc, device=dev)
            _, lp2, cs2, _, _ = ops.layernorm_bwd_mixed(dyf, xl, g, mean2, rstd2, None, lp_dtype=dt, want_dx=False)
            yr2.backward(dyf)
            ok &= report(f"lnx_bwd (dy fp32, x {dt}) dx_lp", lp2, xr2.grad, tol)
            ok &= report(f"lnx_bwd (dy fp32) colsum", cs2, lp2.float().sum(0), 2e-4)
        for (B, H, W, Cc) in ((2, 8, 8, 32), (3, 7, 7, 64), (2, 14, 14, 128)):
            xs = torch.randn(B, H, W, Cc, device=dev)
            g = torch.randn(4 * Cc, device=dev) * 0.5 + 1
            b = torch.randn(4 * Cc, device=dev) * 0.1
            y, mean, rstd = ops.patch_merge_ln_fwd_mixed(xs, g, b, dt)
            xr = xs.clone().requires_grad_(True)
            xp = F.pad(xr, (0, 0, 0, W % 2, 0, H % 2))
            Hp, Wp = xp.shape[1], xp.shape[2]
            xg = xp.reshape(B, Hp // 2, 2, Wp // 2, 2, Cc).permute(0, 1, 3, 4, 2, 5).flatten(3)
            yr = F.layer_norm(xg, (4 * Cc,), g, b, 1e-5)
            ok &= report(f"merge_lnx_fwd f32->{dt} {B,H,W,Cc}", y, yr, tol)
            dy = torch.randn_like(yr).to(dt)
            dres = torch.randn(B, H, W, Cc, device=dev)
            rs = torch.rand(B, device=dev) + 0.5
            dx, lp, cs, dg, db = ops.patch_merge_ln_bwd_mixed(dy, xs, g, mean, rstd, dres, rowscale=rs)
            yr.backward(dy.float())
            ok &= report(f"merge_lnx_bwd dx", dx, xr.grad + dres, 2e-5)
            ok &= report(f"merge_lnx_bwd dx_lp", lp, (xr.grad + dres) * rs[:, None, None, None], tol)
            ok &= report(f"merge_lnx_bwd lp_colsum", cs, lp.float().sum((0, 1, 2)), 2e-4)
        gs = torch.randn(777, 256, device=dev)
        rs = torch.rand(7, device=dev) + 0.5
        yq, cs = ops.scale_cast_colsum(gs, dt, rs, 111)
        ok &= report(f"scale_cast_colsum {dt}", yq, gs * rs[torch.arange(777, device=dev) // 111][:, None], tol)
        ok &= report(f"scale_cast_colsum sums", cs, yq.float().sum(0), 2e-4)
        for (rows, Cc) in ((777, 256), (5000, 128), (33, 8), (100352 // 4, 264), (130, 520)):
            xs_ = torch.randn(rows, Cc, device=dev).to(dt)
            ok &= report(f"colsum {dt} {rows, Cc}", ops.colsum(xs_), xs_.double().sum(0).float(), 2e-4 if dt == torch.float32 else 1e-3)
        acc_ = torch.ones(136, device=dev)[4:]                  # 16-byte aligned but not 32: vector reductions; then a misaligned view
        ok &= report(f"colsum accumulates (offset view)", ops.colsum(torch.ones(300, 128, device=dev).to(dt), acc_[:128]), torch.full((128,), 301.0, device=dev), 1e-6)
        acc2_ = torch.zeros(137, device=dev)[1:]
        ok &= report(f"colsum scalar-atomic path (4-byte aligned out)", ops.colsum(torch.ones(300, 136, device=dev).to(dt), acc2_), torch.full((136,), 300.0, device=dev), 1e-6)
        xc = torch.randn(3, 50, 40, device=dev)
        ok &= report(f"convert f32->{dt} transpose", ops.convert(xc, dt, True), xc.transpose(1, 2), tol)
        ok &= report(f"convert {dt}->f32", ops.convert(xc.to(dt), torch.float32), xc.to(dt).float(), 0)
        x = torch.randn(3, 7, 9, 40, device=dev).to(dt)
        ok &= report(f"nhwc_to_nchw {dt}", ops.nhwc_to_nchw(x), x.permute(0, 3, 1, 2), 0)
        ok &= report(f"nchw_to_nhwc {dt}", ops.nchw_to_nhwc(x.permute(0, 3, 1, 2).contiguous()), x, 0)
    return ok


def _lin_cases():
    return [(256, 128, 128), (1000, 384, 128), (300, 96, 288), (777, 512, 2048), (1568, 2048, 512), (128, 64, 64), (100, 256, 1024),
            (20000, 384, 128), (6272, 512, 1536), (3000, 48, 64)]   # > 148 tiles: several tiles per persistent CTA


def _gemm_group(backend, dts):
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import torch.nn.functional as F
    from mtus_b200 import ops
    ok = True
    dev = "cuda"
    for dt, tol in dts:
        for (M, N, K) in _lin_cases():
            x = (torch.randn(M, K, device=dev) * 0.5).to(dt)
            w = (torch.randn(N, K, device=dev) * 0.05).to(dt)
            bias = torch.randn(N, device=dev) * 0.1
            res = torch.randn(M, N, device=dev).to(dt)
            rs = torch.rand(4, device=dev) + 0.5
            rps = (M + 3) // 4
            y = ops.linear_fwd(x, w, bias, backend=backend)
            yr = x.float() @ w.float().t() + bias
            ok &= report(f"linear_fwd {dt} M{M} N{N} K{K}", y, yr, tol)
            y2 = ops.linear_fwd(x, w, bias, res=res, rowscale=rs, rows_per_sample=rps, backend=backend)
            sc = rs[torch.arange(M, device=dev) // rps][:, None]
            ok &= report(f"linear_fwd+res+rowscale", y2, res.float() + sc * yr, tol)
            a, pre = ops.linear_fwd(x, w, bias, gelu=True, backend=backend)
            ok &= report(f"linear_fwd+gelu (act)", a, F.gelu(yr), tol)
            ok &= report(f"linear_fwd+gelu (pre)", pre, yr, tol)
            dy = (torch.randn(M, N, device=dev) * 0.5).to(dt)
            dx = ops.linear_dgrad(dy, w, backend=backend)
            ok &= report(f"linear_dgrad", dx, dy.float() @ w.float(), tol)
            hp = torch.randn(M, K, device=dev).to(dt)
            dx2 = ops.linear_dgrad(dy, w, gelu_pre=hp, backend=backend)
            hpf = hp.float().requires_grad_(True)
            F.gelu(hpf).backward(dy.float() @ w.float())
            ok &= report(f"linear_dgrad*gelu'", dx2, hpf.grad, tol)
            resf = torch.randn(M, N, device=dev)
            ys = ops.linear_fwd_stream(x, w, bias, res=resf, rowscale=rs, rows_per_sample=rps, backend=backend)
            ok &= report(f"linear_fwd_stream (fp32 out + fp32 residual)", ys, resf + sc * yr, max(tol * 0.05, 2e-5) if dt == torch.bfloat16 else tol)
            dx3, cs3 = ops.linear_dgrad(dy, w, gelu_pre=hp, backend=backend, with_colsum=True)
            # the fused column sum adds the fp32 (un-rounded) values: compare with the fp32 reference of dx, not with the
            # sum of the stored bf16 values (those differ by the accumulated rounding, ~2^-9 of the term scale)
            # (tcgen05 engine: fp32 sums, limited by the bf16 inputs of the reference product; SIMT engine in bf16: a column-sum
            # pass over the STORED values, i.e. sqrt(M) accumulated roundings of 2^-9 against a sum that is itself ~sqrt(M) terms)
            ok &= report(f"linear_dgrad*gelu' + colsum", cs3, hpf.grad.sum(0), 2e-4 if dt == torch.float32 else (2e-3 if backend == 2 else 1e-2))
            # the MLP pair that stores GELU' instead of the pre-activation (the executor's training path)
            a2, dact = ops.linear_fwd_gelu_dact(x, w, bias, backend=backend)
            yrg = yr.clone().requires_grad_(True)
            F.gelu(yrg).sum().backward()
            ok &= report(f"linear_fwd_gelu_dact (act)", a2, F.gelu(yr), tol)
            ok &= report(f"linear_fwd_gelu_dact (gelu')", dact, yrg.grad, tol)
            dmul = (torch.rand(M, K, device=dev) * 1.2 - 0.1).to(dt)
            dx4, cs4 = ops.linear_dgrad_dact(dy, w, dmul, backend=backend, with_colsum=True)
            ref4 = (dy.float() @ w.float()) * dmul.float()
            ok &= report(f"linear_dgrad_dact", dx4, ref4, tol)
            ok &= report(f"linear_dgrad_dact + colsum", cs4, ref4.sum(0), 2e-4 if dt == torch.float32 else (2e-3 if backend == 2 else 1e-2))
            dw, db = ops.linear_wgrad(dy, x, backend=backend)
            ok &= report(f"linear_wgrad dw", dw, dy.float().t() @ x.float(), max(tol, 1e-4))
            ok &= report(f"linear_wgrad db", db, dy.float().sum(0), max(tol, 1e-4))
        for (B, H, W, Cin, Cout) in ((2, 7, 7, 64, 128), (2, 14, 14, 256, 128), (1, 28, 28, 128, 128), (2, 56, 56, 64, 64), (3, 56, 56, 128, 128)):
            x = (torch.randn(B, H, W, Cin, device=dev) * 0.5).to(dt)
            w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
            wf, wd = ops.conv3x3_repack(w, dt)
            wq = wf.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).contiguous()   # as stored (rounded) weights
            xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            wr = wq.clone().requires_grad_(True)
            yr = F.conv2d(xr, wr, padding=1)
            y = ops.conv3x3_fwd(x, wf, backend=backend)
            ok &= report(f"conv3x3_fwd {dt} {B,H,W,Cin,Cout}", y, yr.permute(0, 2, 3, 1), tol)
            dy = (torch.randn(B, H, W, Cout, device=dev) * 0.5).to(dt)
            yr.backward(dy.float().permute(0, 3, 1, 2))
            dx = ops.conv3x3_dgrad(dy, wd, backend=backend)
            ok &= report(f"conv3x3_dgrad", dx, xr.grad.permute(0, 2, 3, 1), tol)
            dw = ops.conv3x3_wgrad(dy, x, backend=(0 if (backend == 2 and Cin % 64) else backend))   # tcgen05 conv wgrad: 64-wide tiles for Cin % 128 != 0
            ok &= report(f"conv3x3_wgrad", dw, wr.grad, max(tol, 1e-4))
    return ok


def g_gemm_simt():
    import torch
    return _gemm_group(1, ((torch.float32, 2e-5), (torch.bfloat16, 2e-2)))


def g_gemm_tc():
    import torch
    return _gemm_group(2, ((torch.bfloat16, 2e-2),))


def _attn_ref(qkv, table, qkv_bias, heads, win, shift):
    """Reference through the oracle's WindowAttention pieces (roll -> pad -> partition -> attention -> reverse)."""
    import torch
    import torch.nn.functional as F
    from oracle import swin as osw
    B, H, W, C3 = qkv.shape
    Cc = C3 // 3
    x = qkv
    if shift:
        x = torch.roll(x, (-shift, -shift), (1, 2))
    ph, pw = (win - H % win) % win, (win - W % win) % win
    if ph or pw:
        xb = qkv_bias.view(1, 1, 1, C3).expand(B, H + ph, W + pw, C3).clone()
        xb[:, :H, :W] = x
        x = xb
    Hp, Wp = H + ph, W + pw
    xw = osw.window_partition(x, (win, win)).view(-1, win * win, 3, heads, 32).permute(2, 0, 3, 1, 4)
    q, k, v = xw[0] * (32 ** -0.5), xw[1], xw[2]
    attn = q @ k.transpose(-2, -1)
    idx = osw.relative_position_index(win, win).to(qkv.device)
    attn = attn + table[idx.view(-1)].view(win * win, win * win, -1).permute(2, 0, 1).unsqueeze(0)
    if shift:
        blk = osw.SwinTransformerBlock.__new__(osw.SwinTransformerBlock)
        blk.input_resolution, blk.window_size, blk.shift_size, blk.window_area = (H, W), (win, win), (shift, shift), win * win
        mask = osw.SwinTransformerBlock.get_attn_mask(blk).to(qkv.device)
        nW = mask.shape[0]
        attn = (attn.view(-1, nW, heads, win * win, win * win) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, win * win, win * win)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, win, win, Cc)
    o = osw.window_reverse(o, (win, win), Hp, Wp)[:, :H, :W]
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    return o


def g_attention():
    import torch
    from mtus_b200 import ops
    ok = True
    dev = "cuda"
    cases = [(2, 14, 14, 2, 7, 0), (2, 14, 14, 2, 7, 3), (1, 56, 56, 4, 7, 3), (2, 7, 7, 8, 7, 0), (2, 4, 4, 1, 4, 0),
             (1, 24, 24, 3, 12, 6), (2, 16, 16, 2, 7, 3), (1, 9, 9, 1, 7, 3),
             # 65..144-token windows (attention_mma144.cu in bf16): padded map, 81-token windows, one unshifted window
             (2, 20, 20, 2, 12, 6), (1, 36, 36, 2, 9, 4), (2, 12, 12, 3, 12, 0), (1, 48, 48, 1, 12, 6)]
    for dt, tol in ((torch.float32, 5e-5), (torch.bfloat16, 3e-2)):
        for (B, H, W, heads, win, shift) in cases:
            Cc = heads * 32
            qkv = torch.randn(B, H, W, 3 * Cc, device=dev).to(dt)
            table = torch.randn((2 * win - 1) ** 2, heads, device=dev) * 0.5
            bias = torch.randn(3 * Cc, device=dev) * 0.5
            out = ops.window_attn_fwd(qkv, table, bias, heads, win, shift)
            qr = qkv.float().requires_grad_(True)
            tr, br = table.clone().requires_grad_(True), bias.clone().requires_grad_(True)
            ref = _attn_ref(qr, tr, br, heads, win, shift)
            ok &= report(f"attn_fwd {dt} B{B} {H}x{W} h{heads} w{win} s{shift}", out, ref, tol)
            dout = torch.randn(B, H, W, Cc, device=dev).to(dt)
            ref.backward(dout.float())
            dqkv, dtab, dbias, dcol = ops.window_attn_bwd(dout, qkv, out, table, bias, heads, win, shift, with_colsum=True)
            ok &= report(f"attn_bwd dqkv", dqkv, qr.grad, tol)
            ok &= report(f"attn_bwd colsum(dqkv) (qkv Linear bias grad)", dcol, qr.grad.sum((0, 1, 2)), max(tol, 2e-4))
            ok &= report(f"attn_bwd dtable", dtab, tr.grad, max(tol, 2e-4))
            if br.grad is not None:
                ok &= report(f"attn_bwd dqkv_bias (pad tokens)", dbias, br.grad, max(tol, 2e-4))
    return ok


def g_fpn_ops():
    import torch
    import torch.nn.functional as F
    from mtus_b200 import ops
    ok = True
    dev = "cuda"
    for dt, tol in ((torch.float32, 2e-5), (torch.bfloat16, 2e-2)):
        for (B, H, W, Cc) in ((2, 7, 7, 128), (3, 14, 14, 128), (2, 56, 56, 128), (1, 28, 28, 64)):
            x = (torch.randn(B, H, W, Cc, device=dev) * 2 + 0.3).to(dt)
            g = torch.randn(Cc, device=dev) * 0.5 + 1
            b = torch.randn(Cc, device=dev) * 0.1
            y, mean, rstd = ops.groupnorm_relu_fwd(x, g, b)                       # fused cluster kernel where a sample fits
            y2, mean2, rstd2 = ops.groupnorm_relu_fwd_two_pass(x, g, b)           # separate statistics / normalise kernels
            xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
            pre = F.group_norm(xr, 32, gr, br, 1e-5)
            yr = F.relu(pre)
            ok &= report(f"gn_relu_fwd {dt} {B,H,W,Cc}", y, yr.permute(0, 2, 3, 1), tol)
            ok &= report(f"gn_relu_fwd two-pass", y2, yr.permute(0, 2, 3, 1), tol)
            ok &= report(f"gn fused mean == two-pass mean", mean, mean2, 1e-5)
            ok &= report(f"gn fused rstd == two-pass rstd", rstd, rstd2, 1e-5)
            dy = torch.randn(B, H, W, Cc, device=dev).to(dt)
            # the ReLU gate is taken from the kernel's own output: an element whose pre-activation rounds to
            # +-1e-8 may land on either side of 0 and would flip a whole dy term (seen once per ~1e6 elements);
            # the gates must agree everywhere the pre-activation is not within rounding of zero
            gate = (y.float() > 0).permute(0, 3, 1, 2)
            flips = (gate != (pre > 0)) & (pre.abs() > 1e-2 if dt == torch.bfloat16 else pre.abs() > 1e-5)
            if flips.any():
                print(f"  FAIL gn_relu gate differs from the oracle at {int(flips.sum())} elements away from 0")
                ok = False
            (pre * gate).backward(dy.float().permute(0, 3, 1, 2))
            dx, dg, db = ops.groupnorm_relu_bwd(dy, x, y, mean, rstd, g)
            ok &= report(f"gn_relu_bwd dx", dx, xr.grad.permute(0, 2, 3, 1), max(tol, 1e-4))
            ok &= report(f"gn_relu_bwd dgamma", dg, gr.grad, max(tol, 3e-4))
            ok &= report(f"gn_relu_bwd dbeta", db, br.grad, max(tol, 3e-4))
            # gate recomputed from x (beta given, y not read): the same expression as the forward, so the same gate
            dx_r, dg_r, db_r = ops.groupnorm_relu_bwd(dy, x, None, mean, rstd, g, beta=b)
            # (the group sums are accumulated with atomics: two runs differ in the last bits, i.e. by one bf16 ulp of dx in bf16 mode)
            rt = 1e-4 if dt == torch.float32 else 5e-3
            ok &= report(f"gn_relu_bwd (recomputed gate) dx", dx_r, dx, rt)
            ok &= report(f"gn_relu_bwd (recomputed gate) dgamma", dg_r, dg, 1e-4)
            ok &= report(f"gn_relu_bwd (recomputed gate) dbeta", db_r, db, 1e-4)
            xr2 = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            ur = F.interpolate(xr2, scale_factor=2.0, mode="bilinear", align_corners=True)
            ok &= report(f"bilinear2x_fwd", ops.bilinear2x_fwd(x), ur.permute(0, 2, 3, 1), tol)
            du = torch.randn(B, 2 * H, 2 * W, Cc, device=dev).to(dt)
            ur.backward(du.float().permute(0, 3, 1, 2))
            ok &= report(f"bilinear2x_bwd", ops.bilinear2x_bwd(du), xr2.grad.permute(0, 2, 3, 1), tol)
            # BatchNorm2d + ReLU over NHWC rows (the GroupNorm kernels with one group per channel), training and eval
            for training in (True, False):
                xb = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
                gb, bb = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
                rm, rv = torch.randn(Cc, device=dev) * 0.1, torch.rand(Cc, device=dev) + 0.5
                preb = F.batch_norm(xb, rm.clone(), rv.clone(), gb, bb, training, 0.1, 1e-5)
                if training:
                    mean_k, rstd_k = ops.batchnorm_stats(x)
                else:
                    mean_k, rstd_k = rm.clone(), (rv + 1e-5).rsqrt()
                yk = ops.batchnorm_relu_fwd(x, mean_k, rstd_k, g, b)
                ok &= report(f"bn_relu_fwd training={training}", yk, F.relu(preb).permute(0, 2, 3, 1), tol)
                gate_b = (yk.float() > 0).permute(0, 3, 1, 2)
                (preb * gate_b).backward(dy.float().permute(0, 3, 1, 2))
                dxk, dgk, dbk = ops.batchnorm_relu_bwd(dy, x, yk, mean_k, rstd_k, g, training=training)
                ok &= report(f"bn_relu_bwd dx training={training}", dxk, xb.grad.permute(0, 2, 3, 1), max(tol, 1e-4))
                ok &= report(f"bn_relu_bwd dgamma", dgk, gb.grad, max(tol, 3e-4))
                ok &= report(f"bn_relu_bwd dbeta", dbk, bb.grad, max(tol, 3e-4))
            top = torch.randn(B, H, W, Cc, device=dev).to(dt)
            skip = torch.randn(B, 2 * H, 2 * W, Cc, device=dev).to(dt)
            refu = skip.float() + F.interpolate(top.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
            ok &= report(f"upsample_add_fwd", ops.upsample_add_fwd(skip, top), refu, tol)
            ok &= report(f"upsample_add_bwd", ops.upsample_add_bwd(du), F.avg_pool2d(du.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1) * 4, tol)
        # pointwise (1x1) convolution with a few output channels: the head tails
        for (B, H, W, K, N) in ((2, 56, 56, 128, 2), (3, 14, 14, 128, 5), (1, 7, 9, 64, 1), (2, 5, 5, 256, 8), (1, 3, 3, 8, 3)):
            xs = torch.randn(B, H, W, K, device=dev).to(dt)
            wq = torch.randn(N, K, device=dev) * 0.1
            bq = torch.randn(N, device=dev)
            xr3 = xs.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            wr3, br3 = wq.clone().requires_grad_(True), bq.clone().requires_grad_(True)
            yr3 = F.conv2d(xr3, wr3[:, :, None, None], br3)
            ok &= report(f"pointwise_conv_fwd {dt} {B,H,W,K,N}", ops.pointwise_conv_fwd(xs, wq, bq), yr3, max(tol, 1e-5))
            dy3 = torch.randn(B, N, H, W, device=dev)
            yr3.backward(dy3)
            dx3, dw3, db3 = ops.pointwise_conv_bwd(dy3, xs, wq)
            ok &= report(f"pointwise_conv_bwd dx", dx3, xr3.grad.permute(0, 2, 3, 1), max(tol, 1e-5))
            ok &= report(f"pointwise_conv_bwd dw", dw3, wr3.grad, 2e-4)
            ok &= report(f"pointwise_conv_bwd dbias", db3, br3.grad, 2e-4)
    return ok


FPN_TASK_TYPES = ("segmentation", "detection")


def _model_group(precision, gemm_env, tol, cos_min, variants, cos_fpn=None):
    """cos_min: per-tensor gradient cosine floor (north star: 0.999); cos_fpn: the floor for bf16 gradients of tasks that run
    through the FPN (measured 0.9976 at random init), None = same as cos_min."""
    import torch
    torch.backends.cudnn.allow_tf32 = False          # the oracle (and the PyTorch heads) must be true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    if gemm_env:
        os.environ["MTUS_GEMM"] = gemm_env
    import mtus_b200 as m
    from oracle.model import OracleMultiTaskModel
    ok = True
    for (enc, img, batch, tasks) in variants:
        cfg = m.make_config(enc, img, batch, mixed_precision=(precision == "bf16"), dropout=0.0,
                            tasks=[t for t in m.tasks_27() if t["task_id"] in (tasks or ["T1_fetal_planes"])])
        torch.manual_seed(0)
        gen = torch.Generator().manual_seed(1)
        x = torch.randn(batch, 3, img, img, generator=gen).cuda()
        if tasks is None:   # encoder only (odd-size merging / window clipping / padding cases the FPN cannot take)
            from oracle.model import OracleSwinEncoder
            oracle = OracleSwinEncoder(enc, img, drop_path_rate=0.0).cuda().eval()
            model = m.SwinTransformerEncoder(enc, pretrained=False, img_size=img, precision=precision).cuda().eval()
            model.load_state_dict(oracle.state_dict())
            fo, fm = oracle(x), model(x)
            for i, (a, b) in enumerate(zip(fm, fo)):
                ok &= report(f"{enc}@{img} {precision} feature[{i}] {tuple(b.shape)}", a, b, tol)
            sum(f.float().square().mean() for f in fo).backward()
            sum(f.float().square().mean() for f in fm).backward()
            ok &= _compare_grads(f"{enc}@{img} {precision} grads[encoder-only]", model, oracle, cos_min)
            continue
        oracle = OracleMultiTaskModel(cfg, drop_path_rate=0.0).cuda().eval()
        model = m.build_model(cfg, precision=precision).cuda().eval()
        model.load_state_dict(oracle.state_dict())
        fo = oracle.encoder(x)
        fm = model.encoder(x)
        for i, (a, b) in enumerate(zip(fm, fo)):
            ok &= report(f"{enc}@{img} {precision} feature[{i}] {tuple(b.shape)}", a, b, tol)
        for tid in tasks:
            oracle.zero_grad(set_to_none=True)
            model.zero_grad(set_to_none=True)
            yo = oracle(x, tid)
            ym = model(x, tid)
            ok &= report(f"{enc}@{img} {precision} output[{tid}]", ym, yo, tol)
            yo.float().square().mean().backward()
            ym.float().square().mean().backward()
            via_fpn = model.task_id_to_name[tid] in FPN_TASK_TYPES
            ok &= _compare_grads(f"{enc}@{img} {precision} grads[{tid}]", model, oracle, cos_fpn if (via_fpn and cos_fpn) else cos_min)
    return ok


def _compare_grads(label, model, oracle, cos_min):
    import torch
    ok = True
    worst, worst_name, n = 1.0, "", 0
    po = dict(oracle.named_parameters())
    for name, p in model.named_parameters():
        go = po[name].grad
        if go is None:
            if p.grad is not None and p.grad.abs().max() > 0:
                print(f"  FAIL {name}: oracle has no grad but kernel path produced one")
                ok = False
            continue
        if p.grad is None:
            print(f"  FAIL {name}: missing gradient")
            ok = False
            continue
        a, b = p.grad.float().flatten(), go.float().flatten()
        if b.norm() == 0 and a.norm() == 0:
            continue
        c = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        rn = (a.norm() / (b.norm() + 1e-30)).item()
        n += 1
        if c < worst:
            worst, worst_name = c, name
        if c < cos_min or not (0.9 < rn < 1.1):
            print(f"    low: {name} cos={c:.6f} norm ratio={rn:.4f}")
    bad = not (worst >= cos_min)
    print(f"  {'FAIL' if bad else 'ok  '} {label}: {n} tensors, min cosine {worst:.6f} ({worst_name})", flush=True)
    return ok and not bad


_VARIANTS_SMALL = [("swin_micro_patch4_window7_test", 56, 2, None),
                   ("swin_micro_patch4_window7_test", 112, 3, None),
                   ("swin_micro_patch4_window7_test", 64, 2, ["T2A_fetal_abdomen", "T1_fetal_planes"]),
                   ("swin_micro_patch4_window7_test", 128, 2, ["T4A_fetal_brain", "T5_fetal_femur"]),
                   ("swin_t", 224, 2, ["T2A_fetal_abdomen"])]


def g_model_fp32():
    return _model_group("fp32", None, 1e-4, 0.999, _VARIANTS_SMALL)


def g_model_bf16_simt():
    return _model_group("bf16", "simt", 2e-2, 0.999, _VARIANTS_SMALL[:4], cos_fpn=0.993)


def g_model_bf16_tc():
    return _model_group("bf16", "tc", 2e-2, 0.999, _VARIANTS_SMALL, cos_fpn=0.993)


def main():
    if len(sys.argv) > 1:
        name = sys.argv[1]
        t = time.time()
        ok = globals()["g_" + name]()
        import torch
        torch.cuda.synchronize()
        print(f"[{name}] {'PASS' if ok else 'FAIL'} in {time.time() - t:.1f}s", flush=True)
        sys.exit(0 if ok else 1)
    results = {}
    for g in GROUPS:
        print(f"=== {g} ===", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), g], timeout=600)
            results[g] = r.returncode
        except subprocess.TimeoutExpired:
            results[g] = "timeout"
    print("SUMMARY", results)


if __name__ == "__main__":
    main()
