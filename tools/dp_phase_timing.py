"""Device time of the phases of one data-parallel training step (forward+loss | backward incl. overlapped all-reduce |
finish() | clip+optimizer), rank 0, CUDA events.  Run under torchrun for N > 1, or plainly for N = 1:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/dp_phase_timing.py [task_id]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import mtus_b200 as m
from mtus_b200.losses import compute_task_loss

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
tids = sys.argv[1:] or ["T1_fetal_planes", "T2A_fetal_abdomen"]
B = 32
cfg = m.swin_b_27task(batch_size=B)
torch.manual_seed(0)
model = m.build_model(cfg, precision="bf16").to(dev).train()
opt = m.build_flat_optimizer(model, cfg)
fns, w = m.build_all_losses(cfg)
tr = m.DataParallelTrainer(model, opt, fns, w)
tcfg = {t["task_id"]: t for t in cfg.get_task_configs()}
for tid in tids:
    x, y = m.synthetic_batch(tcfg[tid], B, 224, generator=torch.Generator().manual_seed(rank), device=dev)
    name = model.task_id_to_name[tid]
    for _ in range(4):
        tr.step(x, y, tid)
    torch.cuda.synchronize()
    n = 15
    acc = [0.0] * 4
    for _ in range(n):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        if world > 1:
            dist.barrier()
        ev[0].record()
        out = model(x, task_id=tid)
        loss = compute_task_loss(fns, name, out, y) * w.get(name, 1.0)
        opt.zero_grad()
        ev[1].record()
        with tr.reducer:
            loss.backward()
        ev[2].record()
        tr.reducer.finish()
        ev[3].record()
        opt.step()
        ev[4].record()
        torch.cuda.synchronize()
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1]) / n
    if rank == 0:
        print(f"world {world} {tid}: forward+loss {acc[0]:.3f} | backward {acc[1]:.3f} | finish {acc[2]:.3f} | optimizer {acc[3]:.3f} | total {sum(acc):.3f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
