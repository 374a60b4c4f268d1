"""In-situ executor timing of one swin_l@384 (window 12) training step: MTUS_TIME_KERNELS=1 python tools/insitu_swinl.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mtus_b200 as m
dev = torch.device("cuda", 0)
tb = 8
cfg = m.make_config("swin_large_patch4_window12_384", 384, tb, mixed_precision=True)
torch.manual_seed(0)
model = m.build_model(cfg, precision="bf16").to(dev).train()
opt = m.build_flat_optimizer(model, cfg)
fns, w = m.build_all_losses(cfg)
tr = m.DataParallelTrainer(model, opt, fns, w)
tcfg = {t["task_id"]: t for t in cfg.get_task_configs()}
tid = "T1_fetal_planes"
x, y = m.synthetic_batch(tcfg[tid], tb, 384, generator=torch.Generator().manual_seed(0), device=dev)
for _ in range(3):
    tr.step(x, y, tid)
torch.cuda.synchronize()
