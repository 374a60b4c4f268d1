"""Cycle-stamp timeline of CTA 0 of one gemm_tc2_kernel launch (diagnostics; MTUS_T2_TRACE).
python tools/gemm_trace.py <stage> <qkv|proj|fc1|fc2> <fwd|dgrad|wgrad>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

dev = torch.device("cuda", 0)
buf = torch.zeros(20 * 512 + 148 * 32, dtype=torch.int64, device=dev)
os.environ["MTUS_T2_TRACE"] = hex(buf.data_ptr())
from mtus_b200 import kbench

stage, name, direction = int(sys.argv[1]), sys.argv[2], sys.argv[3]
_, M, Cc = kbench.swin_b_stage_shapes(32)[stage - 1]
N, K, kind = [(n, k, kd) for nm, n, k, kd in kbench.linear_cases(M, Cc) if nm == name][0]
t, fl, by = kbench.time_linear(M, N, K, kind, direction, dev)
torch.cuda.synchronize()
print(f"{name}.{direction} stage {stage}: {t * 1e6:.1f} us")
tr = buf.cpu()[:20 * 512].view(20, 512)
TAGS = {0: "entry", 14: "setup done", 1: "pdl_wait done", 2: "P: stage free", 3: "M: tile start (tempty)", 4: "M: full", 5: "M: commit tile",
        6: "E: tfull", 16: "E: x arrived", 7: "E: tmem ld done", 8: "E: math + staging done", 10: "E: fence done", 9: "E: previous stores read (thread 0)",
        17: "E: barrier passed", 11: "E: store issued", 12: "E: drained", 13: "exit", 15: "X: buffer free"}
ev = []
for w in range(20):
    for i in range(512):
        v = int(tr[w, i])
        if v == 0:
            break
        ev.append((v & 0xffffffffffff, w, (v >> 48) & 0xffff))
t0 = min(e[0] for e in ev)
only = {0, 1, 3, 4, 11}   # warps printed in full: producer, MMA, first epilogue warps
for c, w, tag in sorted(ev):
    if w in only or tag in (0, 1, 12, 13, 14):
        print(f"{c - t0:8d}  warp {w:2d}  {TAGS.get(tag, tag)}")
# per-CTA wall-clock stamps (globaltimer ns) of the last launch
full = buf.cpu()
g = full[20 * 512:20 * 512 + 148 * 32].view(148, 32)
names = {0: "entry", 1: "released", 2: "first k-block in", 3: "tile 1 committed", 4: "tile 2 committed", 5: "tile 3 committed", 10: "epilogue warps done", 11: "stores drained"}
live = [c for c in range(148) if int(g[c, 0]) != 0]
base = min(int(g[c, 0]) for c in live)
print("per-CTA stamps, ns after the first CTA's entry (min / median / max over CTAs):")
for k, nm in names.items():
    v = sorted(int(g[c, k]) - base for c in live if int(g[c, k]) != 0)
    if v:
        print(f"  {nm:22s} {v[0]:7d} {v[len(v) // 2]:7d} {v[-1]:7d}   (n={len(v)})")
late = sorted(live, key=lambda c: -int(g[c, 11]))[:6]
print("latest CTAs to drain:", [(c, int(g[c, 11]) - base) for c in late])
