"""Per-task-type step times of the headline config (resident inputs), device time and host enqueue time.

    python tools/step_timing.py [steps]       (MTUS_WGRAD_STREAM=0 etc. for A/B runs)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mtus_b200 as m


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda", 0)
    B = 32
    cfg = m.swin_b_27task(batch_size=B)
    torch.manual_seed(0)
    model = m.build_model(cfg, precision="bf16").to(dev).train()
    opt = m.build_flat_optimizer(model, cfg)
    fns, w = m.build_all_losses(cfg)
    tr = m.DataParallelTrainer(model, opt, fns, w)
    tcfg = {t["task_id"]: t for t in cfg.get_task_configs()}
    reps = {}
    for tid, t in tcfg.items():
        reps.setdefault(t["task_name"], tid)
    torch.backends.cudnn.benchmark = True
    total = 0.0
    weights = {"segmentation": 12, "classification": 9, "detection": 3, "regression": 3}
    for name, tid in reps.items():
        x, y = m.synthetic_batch(tcfg[tid], B, 224, generator=torch.Generator().manual_seed(0), device=dev)
        for _ in range(4):
            tr.step(x, y, tid)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            tr.step(x, y, tid)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        hs = []
        for _ in range(3):                      # host enqueue time of ONE step into an empty queue
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            tr.step(x, y, tid)
            hs.append((time.perf_counter() - t2) * 1e3)
        torch.cuda.synchronize()
        total += ms * weights[name.lower()]
        print(f"{name:15s} {tid:28s} device {ms:7.3f} ms/step   host (queue full) {(t1 - t0) * 1e3 / n:7.3f}   host (single step, empty queue) {min(hs):7.3f} ms/step")
    print(f"27-task mean: {total / 27:.3f} ms/step -> {B / (total / 27) * 1e3:.1f} img/s")
    import ctypes as C
    from mtus_b200 import _lib
    h, ms_, n_ = C.c_int64(), C.c_int64(), C.c_int64()
    _lib.lib().mtus_graph_cache_stats(C.byref(h), C.byref(ms_), C.byref(n_))
    print(f"executor graph cache: {h.value} hits, {ms_.value} misses, {n_.value} entries")


if __name__ == "__main__":
    main()
