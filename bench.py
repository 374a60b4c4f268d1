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
# --workload selects the other BASELINE.json configurations (the default line is configs[1] / configs[2])
WORKLOADS = {
    "swin_b_224": {"encoder": "swin_b", "image": 224, "batch": 32, "train": True, "name": WORKLOAD,
                   "metric": METRIC, "train_gf_per_img": 105.9},
    "swin_l_384": {"encoder": "swin_large_patch4_window12_384", "image": 384, "batch": 16, "train": True,
                   "name": "configs[3]: swin_large_patch4_window12_384 + separate FPNs + 27 heads, 384x384, batch 16/GPU, bf16, "
                           "fwd+bwd+clip+AdamW",
                   "metric": "Swin-L/384 window-12 27-task train img/s", "train_gf_per_img": None},
    "swin_b_512_infer": {"encoder": "swin_b", "image": 512, "batch": 128, "train": False,
                         "name": "configs[4]: swin_b 27-task inference (eval, no_grad), 512x512 (padded windows), batch 128, bf16",
                         "metric": "Swin-B 27-task inference img/s @512", "train_gf_per_img": None},
}
# the CPU arm (cpu_baseline and --impl reference) always times THIS sample: same model, same task sequence, fixed batch
CPU_SAMPLE_BATCH = 8


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


def _config_dict(wl, world, B, S):
    """The `config` object of the JSON line; both arms print the same one for the same workload and world size."""
    return {"workload": wl["name"], "global_batch": world * B, "image_size": S, "parallelism": f"dp{world}",
            "task_sequence": "random.Random(42).choice over the 27 task ids per step (MultiTaskUniformSampler)",
            "l2": "no explicit flush: each step streams a multi-GB activation workspace (workspace_gb) plus 0.35 GB of weights, "
                  "far beyond the 126 MB L2",
            "detection_loss": "Detection (SURVEY 8d caveat: shipped YAML pairs the baseline head with the CenterNet loss)"}


def _task_sequence(task_ids, n, seed=42):
    rng = random.Random(seed)            # code/data/dataset.py:145,171 -- one task per step, uniform over ids
    return [rng.choice(task_ids) for _ in range(n)]


# =================================================================================================
# reference arm / cpu_baseline: the oracle port on the host cores
# =================================================================================================
def _cpu_threads():
    """All host cores, regardless of what the launcher exported (torchrun sets OMP_NUM_THREADS=1)."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except Exception:
        pass
    torch.set_num_threads(n)
    return n


class _OracleRunner:
    """The oracle port of the reference's training step (code/train.py:326,440-455) on the host CPUs, fp32."""

    def __init__(self, batch):
        import torch
        import mtus_b200 as m
        from oracle.model import OracleMultiTaskModel
        self.batch = batch
        cfg = m.swin_b_27task(batch_size=batch, mixed_precision=False)
        torch.manual_seed(0)
        self.model = OracleMultiTaskModel(cfg).train()
        enc = list(self.model.encoder.parameters())
        enc_ids = {id(p) for p in enc}
        rest = [p for p in self.model.parameters() if id(p) not in enc_ids]
        self.opt = torch.optim.AdamW([{"params": enc, "lr": 1e-5}, {"params": rest, "lr": 1e-4}], lr=1e-4, weight_decay=1e-4)
        self.tcfg = {t["task_id"]: t for t in cfg.get_task_configs()}
        self.gen = torch.Generator().manual_seed(1)

    def step(self, tid):
        from oracle.model import synthetic_batch, train_step
        x, y = synthetic_batch(self.tcfg[tid], self.batch, 224, self.gen)
        t0 = time.perf_counter()
        loss = train_step(self.model, self.opt, x, y, tid)
        float(loss)
        return time.perf_counter() - t0


def _cpu_sample_text(n_steps, batch, threads):
    return (f"{n_steps} training steps x {batch} images of the swin_b 27-task model (same seeded task sequence as the GPU arm, "
            f"batch {batch} instead of 32), oracle fp32 PyTorch on {threads} threads ({os.cpu_count()} logical cores)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    threads = _cpu_threads()
    task_ids = [t["task_id"] for t in __import__("mtus_b200").tasks_27()]
    n = args.steps + args.warmup
    seq = _task_sequence(task_ids, n)
    batch = CPU_SAMPLE_BATCH
    runner = _OracleRunner(batch)
    t_first = runner.step(seq[0])                     # untimed: first-touch / oneDNN primitive creation
    if t_first * n > 400.0 and batch > 2:            # a slow host: keep the run within minutes, say so in `sample`
        batch = max(1, int(batch * 300.0 / (t_first * n)))
        runner = _OracleRunner(batch)
        runner.step(seq[0])
    times = [runner.step(tid) for tid in seq]
    timed = times[args.warmup:]
    total = sum(timed)
    value = batch * len(timed) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / len(timed), 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the SAME config object as our arm prints (the driver compares the two lines); what the CPU actually ran per step
        # is the bounded sample described next to it
        "config": _config_dict(WORKLOADS["swin_b_224"], world, WORKLOADS["swin_b_224"]["batch"], WORKLOADS["swin_b_224"]["image"]),
        "sample": f"{batch} images per step on the host CPU (bounded sample of the 32-image step), rank 0 only",
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": _cpu_sample_text(len(timed), batch, threads)},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_sample(seq):
    """cpu_baseline leg of our arm (rank 0, N = 1 only): the first steps of the SAME seeded task sequence, same fixed
    batch and thread policy as --impl reference, ~10-30 s of CPU work."""
    threads = _cpu_threads()
    runner = _OracleRunner(CPU_SAMPLE_BATCH)
    runner.step(seq[0])                               # untimed
    times, budget = [], 25.0
    for tid in seq:
        times.append(runner.step(tid))
        if sum(times) > budget or len(times) >= 12:
            break
    value = CPU_SAMPLE_BATCH * len(times) / sum(times)
    return {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": _cpu_sample_text(len(times), CPU_SAMPLE_BATCH, threads)}


# =================================================================================================
# our arm
# =================================================================================================
def _kernel_rooflines(peaks, device):
    """Roofline of the dominant kernel (tcgen05 GEMM) and of the bandwidth-bound kernels, timed host-free: each kernel is a
    CUDA graph of back-to-back launches over a ring of pre-allocated operand sets > 4x the L2 (mtus_b200/kbench.py), one
    graph launch between two CUDA events on the launching stream."""
    import torch
    from mtus_b200 import kbench
    res = {}
    hbm, tc = peaks["hbm"], peaks["tc_sustained"]

    # ---- dominant kernel: gemm_tc2_kernel on the Swin-B stage-3 shapes (18 of 24 blocks, 73 % of encoder FLOPs) ----
    M, Cc = 32 * 196, 512
    table, fwd = {}, []
    for direction in ("fwd", "dgrad", "wgrad"):
        for name, N, K, kind in kbench.linear_cases(M, Cc):
            t, fl, by = kbench.time_linear(M, N, K, kind, direction, device)
            table[f"{name}.{direction}"] = {"us": round(t * 1e6, 2), "tflops": round(fl / t / 1e12, 1), "frac": round(fl / t / 1e12 / tc, 3)}
            if direction == "fwd":
                fwd.append((fl, t, by))
    fl = sum(g[0] for g in fwd)
    tt = sum(g[1] for g in fwd)
    ach = fl / tt / 1e12
    traffic, traffic_src = None, None
    for fn in ("r2_gemm_tc2_traffic.json", "r1_gemm_tc2_traffic.json"):
        try:   # dram bytes of the same four launches from the committed ncu --set full capture (profiles/): STATIC, not this run
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                traffic = json.load(f)["dram_bytes_per_launch_avg"]
            traffic_src = f"static: profiles/{fn} (ncu --set full of these four launches inside a training step; not re-measured by this run)"
            break
        except Exception:
            pass
    all_fl = sum(2.0 * M * n * k for _, n, k, _ in kbench.linear_cases(M, Cc)) * 3
    all_t = sum(v["us"] for v in table.values()) * 1e-6
    res["roofline"] = {"bound": "tensor", "kernel": "gemm_tc2_kernel (persistent tcgen05.mma + TMA + TMEM), Swin-B stage-3 forward GEMMs "
                       "qkv / proj / fc1+GELU / fc2 at M=6272, one launch each", "achieved": round(ach, 1),
                       "peak": tc, "unit": "TFLOP/s", "frac": round(ach / tc, 4), "traffic": traffic, "traffic_source": traffic_src,
                       "peak_source": f"{peaks['src']} (sustained cuBLAS bf16 figure: kernels timed back to back in a long loop)",
                       "timing": "CUDA graph of back-to-back launches over operand rings > 4x L2, one graph launch between two events; no host "
                                 "work, no allocation, no fill kernels in the timed region",
                       "algorithmic_flops_per_launch_avg": fl / 4, "algorithmic_bytes_per_launch_avg": sum(g[2] for g in fwd) / 4,
                       "avg_launch_us": round(tt / 4 * 1e6, 2),
                       "per_shape": table,
                       "fwd_dgrad_wgrad_tflops": round(all_fl / all_t / 1e12, 1), "fwd_dgrad_wgrad_frac": round(all_fl / all_t / 1e12 / tc, 4)}

    extra = []

    def hbm_entry(kernel, t, by, **kw):
        e = {"kernel": kernel, "bound": "hbm", "achieved": round(by / t / 1e9, 1), "peak": hbm, "unit": "GB/s",
             "frac": round(by / t / 1e9 / hbm, 4), "us": round(t * 1e6, 2), "algorithmic_bytes": by}
        e.update(kw)
        extra.append(e)

    for rows, C1, tag in ((32 * 3136, 128, "stage 1"), (32 * 196, 512, "stage 3")):
        t, by = kbench.time_layernorm_fwd(rows, C1, device)
        hbm_entry(f"lnv2_fwd_kernel [{rows},{C1}] fp32 stream -> bf16 ({tag})", t, by)
        t, by = kbench.time_layernorm_bwd(rows, C1, device)
        hbm_entry(f"lnv2_bwd_kernel [{rows},{C1}] bf16 dy, fp32 stream in/out, bf16 copy out ({tag})", t, by)
    t, by = kbench.time_patch_merge_ln(32, 56, 128, device)
    hbm_entry("patch_merge_ln_fwd [32,56,56,128] fp32 -> [32,28,28,512] bf16 (gather + LayerNorm)", t, by)
    t, by = kbench.time_patch_merge_ln(32, 56, 128, device, backward=True)
    hbm_entry("patch_merge_ln_bwd [32,56,56,128] (scatter + LayerNorm backward)", t, by)
    t, by = kbench.time_fpn_upadd(32, 56, 256, device)
    hbm_entry("upadd_fwd FPN top-down p2 = skip + nearest_up(p3) [32,56,56,256] bf16", t, by)
    t, by = kbench.time_fpn_merge(32, 56, 128, 4, device)
    hbm_entry("merge_nhwc_fwd cat + Dropout2d scale -> [32,56,56,512] bf16", t, by)
    t, by = kbench.time_groupnorm_relu(32, 56, 128, device)
    hbm_entry("gn_fused_fwd_kernel GroupNorm(32)+ReLU [32,56,56,128] bf16 (one cluster kernel: statistics through distributed shared "
              "memory, one read + one write)", t, by)
    t, by = kbench.time_groupnorm_relu(32, 56, 128, device, fused=False)
    hbm_entry("gn_stats x2 + gn_relu_fwd, the two-pass fallback of the same op (3 reads + 1 write)", t, by,
              pair_moves_bytes=2.0 * by, frac_of_bytes_moved=round(2.0 * by / t / 1e9 / hbm, 4))
    t, by = kbench.time_groupnorm_bwd(32, 56, 128, device, act=0)
    hbm_entry("gn_fused_bwd_kernel GroupNorm(32)+ReLU backward [32,56,56,128] bf16 (one cluster kernel: x and dy staged once in shared "
              "memory, two reads + one write)", t, by)
    t, by = kbench.time_groupnorm_bwd(32, 56, 128, device, act=1)
    hbm_entry("gn_fused_bwd_kernel GroupNorm(32)+SiLU backward [32,56,56,128] bf16 (segmentation head)", t, by)
    t, by = kbench.time_colsum(32 * 3136, 256, device)
    hbm_entry("colsum_kernel [100352,256] bf16 (FPN lateral bias gradient; one wave, vector reductions)", t, by)
    t, by = kbench.time_patch_merge_ln(32, 14, 512, device, backward=True)
    hbm_entry("patch_merge_ln_bwd [32,14,14,512] (4C = 2048: wide-row kernel, partials reduced across a cluster)", t, by)
    t, by = kbench.time_bilinear(32, 28, 128, device)
    hbm_entry("bilinear2x_fwd [32,28,28,128] -> [32,56,56,128] bf16 (align_corners)", t, by)
    for (H, Cc_, heads, tag) in ((56, 128, 4, "stage 1"), (14, 512, 16, "stage 3")):
        t, by, fl_ = kbench.time_window_attn(32, H, Cc_, heads, 7, 3, device)
        hbm_entry(f"window attention fwd {tag} (shifted) bf16, stand-alone (q,k,v in / o out)", t, by,
                  attn_only_tflops=round(fl_ / t / 1e12, 2), attn_only_frac_of_bf16_peak=round(fl_ / t / 1e12 / tc, 4),
                  padded_mma_flops_factor=round((64 / 49) ** 2, 3))
        t, by, fl_ = kbench.time_window_attn(32, H, Cc_, heads, 7, 3, device, backward=True)
        hbm_entry(f"window attention bwd {tag} (shifted) bf16, stand-alone", t, by,
                  attn_only_tflops=round(fl_ / t / 1e12, 2), attn_only_frac_of_bf16_peak=round(fl_ / t / 1e12 / tc, 4))
    res["roofline_extra"] = extra
    torch.cuda.empty_cache()
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

    wl = WORKLOADS[args.workload]
    B, S = (args.batch or wl["batch"]), wl["image"]
    training = wl["train"]
    cfg = m.make_config(wl["encoder"], S, B, mixed_precision=True)
    torch.manual_seed(0)                       # identical replicas on every rank
    model = m.build_model(cfg, precision="bf16").to(dev)
    model = model.train() if training else model.eval()
    opt = m.build_flat_optimizer(model, cfg)
    loss_fns, loss_w = m.build_all_losses(cfg)
    trainer = m.DataParallelTrainer(model, opt, loss_fns, loss_w, gradient_clip=1.0)
    if not training:
        class _Infer:                          # configs[4]: model.eval() + no_grad forward through the task's head
            def step(self, x, y, tid):
                with torch.no_grad():
                    out = model(x, tid)
                return out.float().mean()      # a scalar to read back in the e2e leg
        trainer = _Infer()
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
        # device -> host read of the step's result, every step: through the trainer's side-stream read-back (the value leaves
        # the device as soon as the forward + loss are done instead of behind the whole backward); plain .item() for inference
        v = trainer.loss_item() if hasattr(trainer, "loss_item") else loss.item()
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
        "metric": wl["metric"], "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": _config_dict(wl, world, B, S), "workspace_gb": ws_gb,
        "e2e": {"value": round(e2e, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d_sum / max(h2d_n, 1)), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if wl["train_gf_per_img"]:
        line["train_flops_per_img"] = wl["train_gf_per_img"] * 1e9
        line["model_tflops"] = round(value * wl["train_gf_per_img"] * 1e9 / 1e12 / world, 1)
    if world == 1 and args.workload == "swin_b_224":
        if not args.no_rooflines:
            del trainer, opt, model
            torch.cuda.empty_cache()
            line.update(_kernel_rooflines(peaks, dev))
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(seq)
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the workload's own, 32 for the headline config)")
    ap.add_argument("--workload", default="swin_b_224", choices=sorted(WORKLOADS),
                    help="swin_b_224 = BASELINE configs[1]/[2] (default, the headline metric); swin_l_384 = configs[3]; "
                         "swin_b_512_infer = configs[4]")
    ap.add_argument("--no-rooflines", action="store_true")
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
