"""cProfile of the host side of one training step (queue drained between steps).  python tools/host_profile.py [task_id]"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mtus_b200 as m

tid = sys.argv[1] if len(sys.argv) > 1 else "T1_fetal_planes"
dev = torch.device("cuda", 0)
B = 32
cfg = m.swin_b_27task(batch_size=B)
torch.manual_seed(0)
model = m.build_model(cfg, precision="bf16").to(dev).train()
opt = m.build_flat_optimizer(model, cfg)
fns, w = m.build_all_losses(cfg)
tr = m.DataParallelTrainer(model, opt, fns, w)
tcfg = {t["task_id"]: t for t in cfg.get_task_configs()}
x, y = m.synthetic_batch(tcfg[tid], B, 224, generator=torch.Generator().manual_seed(0), device=dev)
for _ in range(4):
    tr.step(x, y, tid)
torch.cuda.synchronize()
pr = cProfile.Profile()
for _ in range(10):
    torch.cuda.synchronize()
    pr.enable()
    tr.step(x, y, tid)
    pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
