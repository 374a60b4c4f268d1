#!/usr/bin/env python
"""Headline benchmark: Swin-B 27-task training throughput (img/s) on N B200s -- BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W]          our arm   (sm_100a kernels, bf16)
    python bench.py --impl reference [...]                        reference arm (CPU oracle port, fp32)

A "step" is one full training step of configs[1]: swin_b encoder, separate
FPNs, 27 heads, 224x224, 32 images per GPU, bf16 activations / fp32 master weights:
zero_grad -> forward (task drawn like MultiTaskUniformSampler) -> loss -> backward -> gradient
all-reduce (N > 1) -> clip 1.0 -> AdamW step (code/train.py:326,440-455 of the reference).

One JSON line is printed by rank 0:
  value   whole-job img/s with the step's inputs already resident in HBM (CUDA events, max over ranks)
  e2e     the same metric through the public API (DataParallelTrainer.step) with HOST inputs: pinned
          host -> device copy of images + labels and a device -> host read of the loss every step
  roofline      dominant kernel (tcgen05 GEMM) timed alone with CUDA events on the stage-3 Swin-B shapes
  cpu_baseline  the oracle (fp32 PyTorch restatement of the reference path) on the box's host cores
The reference arm times the oracle port on the host CPUs (the reference has no native code to
compile and its timm / smp dependencies are not installable offline).
"""

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Swin-B 27-task train img/s"
UNIT = "img/s"
WORKLOAD = "configs[1]: swin_b + separate FPNs + 27 heads, 224x224, batch 32/GPU, bf16, fwd+bwd+clip+AdamW"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "tc_burst": float(p["bf16_tflops"]),
                "tc_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tc_burst": 1590.0, "tc_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # "under load": the upper half of the samples (idle samples before/after the region are dropped)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _task_sequence(task_ids, n, seed=42):
    rng = random.Random(seed)            # code/data/dataset.py:145,171 -- one task per step, uniform over ids
    return [rng.choice(task_ids) for _ in range(n)]


# =================================================================================================
# reference arm / cpu_baseline: the oracle port on the host cores
# =================================================================================================
def _oracle_steps(per_step_batch, task_seq, threads=None):
    """Runs one oracle training step per entry of task_seq; returns seconds per step list."""
    import torch
    import mtus_b200 as m
    from oracle.model import OracleMultiTaskModel, synthetic_batch, train_step
    if threads:
        torch.set_num_threads(threads)
    cfg = m.swin_b_27task(batch_size=per_step_batch, mixed_precision=False)
    torch.manual_seed(0)
    model = OracleMultiTaskModel(cfg).train()
    enc = list(model.encoder.parameters())
    enc_ids = {id(p) for p in enc}
    rest = [p for p in model.parameters() if id(p) not in enc_ids]
    opt = torch.optim.AdamW([{"params": enc, "lr": 1e-5}, {"params": rest, "lr": 1e-4}], lr=1e-4, weight_decay=1e-4)
    tcfg = {t["task_id"]: t for t in cfg.get_task_configs()}
    gen = torch.Generator().manual_seed(1)
    out = []
    for tid in task_seq:
        x, y = synthetic_batch(tcfg[tid], per_step_batch, 224, gen)
        t0 = time.perf_counter()
        loss = train_step(model, opt, x, y, tid)
        float(loss)
        out.append(time.perf_counter() - t0)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    cores = os.cpu_count() or 1
    threads = torch.get_num_threads()
    task_ids = [t["task_id"] for t in __import__("mtus_b200").tasks_27()]
    # calibrate: one tiny step, then size the per-step sample so K+W steps end in about two minutes
    t_cal = _oracle_steps(2, ["T2A_fetal_abdomen"])[0] / 2.0
    budget = 120.0
    n = args.steps + args.warmup
    bsz = int(max(1, min(32, budget / (n * max(t_cal, 1e-3)))))
    seq = _task_sequence(task_ids, n)
    times = _oracle_steps(bsz, seq)
    timed = times[args.warmup:]
    total = sum(timed)
    value = bsz * len(timed) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / len(timed), 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{bsz} images per step (bounded sample of the 32-image step)",
                   "device": "host CPU"},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{len(timed)} training steps x {bsz} images, oracle fp32 PyTorch on {threads} threads "
                                   f"({cores} logical cores)"},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_sample():
    """cpu_baseline leg of our arm: ~10-30 s of oracle work on the host cores (rank 0, N = 1 only)."""
    import torch
    threads = torch.get_num_threads()
    task_ids = [t["task_id"] for t in __import__("mtus_b200").tasks_27()]
    t_cal = _oracle_steps(2, ["T2A_fetal_abdomen"])[0] / 2.0
    bsz = int(max(1, min(32, 15.0 / (4 * max(t_cal, 1e-3)))))
    seq = ["T2A_fetal_abdomen", "T1_fetal_planes", "T4A_fetal_brain", "T5_fetal_femur"]   # one of each task type
    times = _oracle_steps(bsz, [seq[0]] + seq)[1:]
    value = bsz * len(times) / sum(times)
    return {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"4 training steps (seg, cls, det, reg) x {bsz} images of the same swin_b 27-task model, "
                      f"oracle fp32 PyTorch, {threads} threads on {os.cpu_count()} logical cores"}


# =================================================================================================
# our arm
# =================================================================================================
def _kernel_rooflines(peaks, device):
    """Times the dominant kernel (tcgen05 GEMM) and the main bandwidth-bound kernels with CUDA events.

    Each kernel is launched back to back over a ring of distinct operand sets whose total size exceeds the 126 MB L2
    several times (every launch reads cold operands; no flush kernel sits inside the timed region), and the average
    launch duration is the event time divided by the number of launches."""
    import torch
    from mtus_b200 import ops, _lib
    res = {}

    def ring_time(make_set, run, bytes_per_set, launches=48):
        n_sets = max(3, int(400e6 // max(bytes_per_set, 1)) + 1)
        sets = [make_set() for _ in range(n_sets)]
        for i in range(min(n_sets, 6)):
            run(sets[i])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(launches):
            run(sets[i % n_sets])
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / launches

    # Swin-B stage 3 (18 of 24 blocks, 73 % of encoder FLOPs), B=32: M = 6272 tokens, C = 512
    M, Cc = 32 * 196, 512
    gemm = []
    for name, K, N, kind in (("qkv", Cc, 3 * Cc, "bias"), ("proj", Cc, Cc, "stream"), ("fc1", Cc, 4 * Cc, "gelu"), ("fc2", 4 * Cc, Cc, "stream")):
        def make_set(K=K, N=N, kind=kind):
            x = (torch.randn(M, K, device=device) * 0.5).bfloat16()
            w = (torch.randn(N, K, device=device) * 0.02).bfloat16()
            b = torch.zeros(N, device=device)
            r = torch.randn(M, N, device=device) if kind == "stream" else None
            return x, w, b, r

        def run(st, kind=kind):
            x, w, b, r = st
            if kind == "bias":
                ops.linear_fwd(x, w, b, backend=_lib.BACKEND_TCGEN05)
            elif kind == "gelu":
                ops.linear_fwd(x, w, b, gelu=True, backend=_lib.BACKEND_TCGEN05)
            else:   # proj / fc2 write the fp32 residual stream: y = res + x w^T + b
                ops.linear_fwd_stream(x, w, b, res=r, backend=_lib.BACKEND_TCGEN05)
        by = (M * K + N * K) * 2 + M * N * (8 if kind == "stream" else (4 if kind == "gelu" else 2))
        t = ring_time(make_set, run, by)
        gemm.append((name, 2.0 * M * K * N, t, by))
    fl = sum(g[1] for g in gemm)
    tt = sum(g[2] for g in gemm)
    ach = fl / tt / 1e12
    traffic = None
    try:   # dram bytes of the same four launches from the committed ncu --set full capture (profiles/)
        with open(os.path.join(ROOT, "profiles", "r1_gemm_tc2_traffic.json")) as f:
            traffic = json.load(f)["dram_bytes_per_launch_avg"]
    except Exception:
        pass
    res["roofline"] = {"bound": "tensor", "kernel": "gemm_tc2_kernel (persistent tcgen05.mma + TMA + TMEM), Swin-B stage-3 forward GEMMs "
                       "qkv / proj / fc1+GELU / fc2 at M=6272, one launch each", "achieved": round(ach, 1),
                       "peak": peaks["tc_sustained"], "unit": "TFLOP/s", "frac": round(ach / peaks["tc_sustained"], 4), "traffic": traffic,
                       "peak_source": f"{peaks['src']} (sustained cuBLAS bf16 figure: kernels timed back to back in a long loop, operands "
                                      "rotated through > 3x the L2 size)",
                       "algorithmic_flops_per_launch_avg": fl / 4, "avg_launch_us": round(tt / 4 * 1e6, 2),
                       "per_shape_tflops": {g[0]: round(g[1] / g[2] / 1e12, 1) for g in gemm}}
    extra = []
    # LayerNorm fwd (fp32 residual stream in, bf16 out), stage 1 shape [100352, 128]: algorithmic bytes rows*C*(4+2)
    rows, C1 = 32 * 3136, 128
    g_, b_ = torch.ones(C1, device=device), torch.zeros(C1, device=device)
    t = ring_time(lambda: torch.randn(rows, C1, device=device), lambda x: ops.layernorm_fwd_mixed(x, g_, b_, torch.bfloat16), rows * C1 * 6)
    by = rows * C1 * 6.0
    extra.append({"kernel": "lnv2_fwd_kernel [100352,128] fp32 -> bf16", "bound": "hbm", "achieved": round(by / t / 1e9, 1),
                  "peak": peaks["hbm"], "unit": "GB/s", "frac": round(by / t / 1e9 / peaks["hbm"], 4)})
    # window attention fwd, stage 1: qkv [32,56,56,384] in, out [32,56,56,128]: 4*C*s bytes per token
    tab = torch.zeros(169, 4, device=device)
    bias = torch.zeros(384, device=device)
    t = ring_time(lambda: torch.randn(32, 56, 56, 384, device=device).bfloat16(), lambda q: ops.window_attn_fwd(q, tab, bias, 4, 7, 3),
                  32 * 3136 * 128 * 8)
    by = 4.0 * 32 * 3136 * 128 * 2
    fl = 4.0 * 49 * 49 * 32 * (32 * 64 * 4)
    extra.append({"kernel": "window_attn_mma_fwd_kernel stage 1 (shifted) bf16", "bound": "hbm", "achieved": round(by / t / 1e9, 1),
                  "peak": peaks["hbm"], "unit": "GB/s", "frac": round(by / t / 1e9 / peaks["hbm"], 4),
                  "attn_only_tflops": round(fl / t / 1e12, 2),
                  "attn_only_frac_of_bf16_peak": round(fl / t / 1e12 / peaks["tc_sustained"], 4)})
    # window attention bwd, stage 1: q, k, v, o, dO in; dq, dk, dv out: 8*C*s bytes per token (+ lse, negligible)
    def make_bwd():
        q = torch.randn(32, 56, 56, 384, device=device).bfloat16()
        o, l = ops.window_attn_fwd(q, tab, bias, 4, 7, 3, return_lse=True)
        return q, o, l, torch.randn_like(o)
    t = ring_time(make_bwd, lambda s_: ops.window_attn_bwd(s_[3], s_[0], s_[1], tab, bias, 4, 7, 3, lse=s_[2], with_colsum=True),
                  32 * 3136 * 128 * 16, launches=24)
    by = 8.0 * 32 * 3136 * 128 * 2
    extra.append({"kernel": "window_attn_mma_bwd_kernel stage 1 (shifted) bf16", "bound": "hbm", "achieved": round(by / t / 1e9, 1),
                  "peak": peaks["hbm"], "unit": "GB/s", "frac": round(by / t / 1e9 / peaks["hbm"], 4),
                  "attn_only_tflops": round(2.5 * fl / t / 1e12, 2)})
    # LayerNorm bwd (bf16 dy, fp32 x, fp32 stream gradient in / out, bf16 operand copy out), stage 1: rows*C*(2+4+4+4+2)
    def make_lnb():
        x = torch.randn(rows, C1, device=device)
        y, mean, rstd = ops.layernorm_fwd_mixed(x, g_, b_, torch.bfloat16)
        return torch.randn(rows, C1, device=device).bfloat16(), x, mean, rstd, torch.randn(rows, C1, device=device)
    t = ring_time(make_lnb, lambda s_: ops.layernorm_bwd_mixed(s_[0], s_[1], g_, s_[2], s_[3], s_[4], lp_dtype=torch.bfloat16),
                  rows * C1 * 16, launches=24)
    by = rows * C1 * 16.0
    extra.append({"kernel": "lnv2_bwd_kernel [100352,128] bf16 dy, fp32 stream", "bound": "hbm", "achieved": round(by / t / 1e9, 1),
                  "peak": peaks["hbm"], "unit": "GB/s", "frac": round(by / t / 1e9 / peaks["hbm"], 4)})
    res["roofline_extra"] = extra
    return res


def run_native(args):
    import torch
    import torch.distributed as dist
    import mtus_b200 as m
    from mtus_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a B200; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    peaks = _peaks()
    L = _lib.lib()

    B, S = args.batch, 224
    cfg = m.swin_b_27task(batch_size=B, image_size=S, mixed_precision=True)
    torch.manual_seed(0)                       # identical replicas on every rank
    model = m.build_model(cfg, precision="bf16").to(dev).train()
    opt = m.build_flat_optimizer(model, cfg)
    loss_fns, loss_w = m.build_all_losses(cfg)
    trainer = m.DataParallelTrainer(model, opt, loss_fns, loss_w, gradient_clip=1.0)
    tcfg = {t["task_id"]: t for t in cfg.get_task_configs()}
    task_ids = list(tcfg.keys())
    n_total = args.warmup + args.steps
    seq = _task_sequence(task_ids, n_total)

    # synthetic data: one pinned host batch per task TYPE and per rank (fresh labels per type), reused across steps
    gen = torch.Generator().manual_seed(1234 + rank)
    host, devb = {}, {}
    for tid in task_ids:
        name = tcfg[tid]["task_name"] + str(tcfg[tid]["num_classes"])
        if name not in host:
            x, y = m.synthetic_batch(tcfg[tid], B, S, generator=gen)
            host[name] = (x.pin_memory(), y.pin_memory())
            devb[name] = (host[name][0].to(dev), host[name][1].to(dev))
    key = {tid: tcfg[tid]["task_name"] + str(tcfg[tid]["num_classes"]) for tid in task_ids}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one untimed step per task id first (cuDNN autotune of the PyTorch heads, optimizer state of every head,
    # allocator growth, NCCL buffers for every gradient size); the W warm-up steps of the contract follow inside timed()
    for tid0 in task_ids:
        trainer.step(devb[key[tid0]][0], devb[key[tid0]][1], tid0)

    def timed(step_fn):
        """W warm-up steps, then K timed steps bracketed by barrier + synchronize; device time, max over ranks.
        step_fn(i, last): i indexes the task sequence; last marks the final step of the (warm-up or timed) run."""
        for i in range(args.warmup):
            step_fn(i, i == args.warmup - 1)
        barrier()
        l0 = L.mtus_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.warmup, n_total):
            step_fn(i, i == n_total - 1)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), L.mtus_launch_count() - l0

    # ---- value: inputs already resident in HBM -------------------------------------------------
    def step_resident(i, last):
        tid = seq[i]
        x, y = devb[key[tid]]
        trainer.step(x, y, tid)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches = timed(step_resident)

    # ---- e2e: host inputs through the public API, loss read back every step -----------------------
    h2d = d2h = 0
    h2d_sum = h2d_n = 0

    # Every step's batch is copied from pinned host memory inside the timed region (K copies for K steps); the copy
    # of step i+1 is issued on DevicePrefetcher's side stream right after step i is enqueued, so it runs under step
    # i's compute instead of in front of step i+1 (the first step of a run has nothing to hide behind).
    prefetch = m.DevicePrefetcher(dev)

    def step_e2e(i, last):
        nonlocal h2d, d2h, h2d_sum, h2d_n
        tid = seq[i]
        x, y = host[key[tid]]
        if not prefetch.pending():
            prefetch.issue(x, y)
        xd, yd = prefetch.take()
        loss = trainer.step(xd, yd, tid)
        if not last:
            prefetch.issue(*host[key[seq[i + 1]]])
        v = loss.item()                          # device -> host read of the step's result
        h2d = x.numel() * x.element_size() + y.numel() * y.element_size()
        if i >= args.warmup:
            h2d_sum += h2d
            h2d_n += 1
        d2h = loss.element_size()
        return v

    ms_e2e, _ = timed(step_e2e)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    ws_gb = None
    try:
        import ctypes as C
        c = model.encoder.model._cfg(B, True)
        ws_gb = round(L.mtus_swin_workspace_bytes(C.byref(c)) / 1e9, 2)
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * B, "image_size": S, "parallelism": f"dp{world}",
                   "task_sequence": "random.Random(42).choice over the 27 task ids per step (MultiTaskUniformSampler)",
                   "l2": f"no explicit flush: each step streams a {ws_gb} GB activation workspace plus 0.35 GB of weights, "
                         "far beyond the 126 MB L2",
                   "detection_loss": "Detection (SURVEY 8d caveat: shipped YAML pairs the baseline head with the CenterNet loss)"},
        "e2e": {"value": round(e2e, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d_sum / max(h2d_n, 1)), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "train_flops_per_img": 105.9e9,
        "model_tflops": round(value * 105.9e9 / 1e12 / world, 1),
    }
    if world == 1:
        line.update(_kernel_rooflines(peaks, dev))
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
    else:
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU (the headline config uses 32)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
