"""BASELINE.json configs[3] and configs[4] as smoke + throughput runs (not bench lines):
  configs[3]: swin_large_patch4_window12_384, 27-task training, bf16 (general attention engine: 144-token windows)
  configs[4]: swin_b, 512x512 (window 7 -> padded 133/70/35/21 maps), inference only, bf16, all 27 task ids
    python tools/run_configs.py [train_batch] [infer_batch]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mtus_b200 as m

dev = torch.device("cuda", 0)
tb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ib = int(sys.argv[2]) if len(sys.argv) > 2 else 128


def timed(fn, n):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


# ---- configs[3]: swin_l 384 window 12 training ----
cfg = m.make_config("swin_large_patch4_window12_384", 384, tb, mixed_precision=True)
torch.manual_seed(0)
model = m.build_model(cfg, precision="bf16").to(dev).train()
opt = m.build_flat_optimizer(model, cfg)
fns, w = m.build_all_losses(cfg)
tr = m.DataParallelTrainer(model, opt, fns, w)
tcfg = {t["task_id"]: t for t in cfg.get_task_configs()}
for tid in ("T2A_fetal_abdomen", "T1_fetal_planes"):
    x, y = m.synthetic_batch(tcfg[tid], tb, 384, generator=torch.Generator().manual_seed(0), device=dev)
    ms = timed(lambda: tr.step(x, y, tid), 5)
    print(f"configs[3] swin_l@384 w12 train B={tb} {tid}: {ms:.2f} ms/step, {tb / ms * 1e3:.1f} img/s, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
del model, opt, tr
torch.cuda.empty_cache()

# ---- configs[4]: swin_b 512x512 inference ----
cfg = m.make_config("swin_b", 512, ib, mixed_precision=True)
torch.manual_seed(0)
model = m.build_model(cfg, precision="bf16").to(dev).eval()
x = torch.randn(ib, 3, 512, 512, device=dev)
tids = [t["task_id"] for t in cfg.get_task_configs()]
with torch.no_grad():
    for tid in (tids[0], "T1_fetal_planes"):
        ms = timed(lambda: model(x, tid), 3)
        print(f"configs[4] swin_b@512 inference B={ib} {tid}: {ms:.2f} ms/batch, {ib / ms * 1e3:.1f} img/s")
    t0 = time.perf_counter()
    for tid in tids:
        out = model(x, tid)
    torch.cuda.synchronize()
    print(f"configs[4] all 27 task ids: {(time.perf_counter() - t0) * 1e3 / 27:.2f} ms/batch mean, out[-1] finite: {bool(torch.isfinite(out.float()).all())}")
