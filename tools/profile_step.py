"""One training step of the headline config between cudaProfilerStart/Stop (for ncu --profile-from-start off).

    python tools/profile_step.py [task_id] [batch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import mtus_b200 as m


def main():
    tid = sys.argv[1] if len(sys.argv) > 1 else "T2A_fetal_abdomen"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = torch.device("cuda", 0)
    cfg = m.swin_b_27task(batch_size=B)
    torch.manual_seed(0)
    model = m.build_model(cfg, precision="bf16").to(dev).train()
    opt = m.build_flat_optimizer(model, cfg)
    fns, w = m.build_all_losses(cfg)
    tr = m.DataParallelTrainer(model, opt, fns, w)
    tcfg = {t["task_id"]: t for t in cfg.get_task_configs()}
    x, y = m.synthetic_batch(tcfg[tid], B, 224, generator=torch.Generator().manual_seed(0), device=dev)
    for _ in range(3):
        tr.step(x, y, tid)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    e0.record()
    tr.step(x, y, tid)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"step {tid} B={B}: {e0.elapsed_time(e1):.3f} ms")


if __name__ == "__main__":
    main()
