"""Per-CTA phase times of pseudo_filter_kernel (OVDET_PSEUDO_DBG_PTR)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ovdet_b200 import synth
from ovdet_b200.utils.box_3d_utils import lift_filter_batch
dev = torch.device("cuda", 0)
S = 4096
bx, pool = synth.pseudo_label_scenes(S, P=256, pool=512, seed=5)
bx, pool = bx.to(dev), pool.to(dev)
lift_filter_batch(bx, pool); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    lift_filter_batch(bx, pool)
e1.record(); torch.cuda.synchronize()
print("kernel us / 4096 scenes", e0.elapsed_time(e1) / 10 * 1e3)
d = torch.zeros((S, 16), dtype=torch.int64, device=dev)
os.environ["OVDET_PSEUDO_DBG_PTR"] = str(d.data_ptr())
lift_filter_batch(bx, pool); torch.cuda.synchronize()
a = d.cpu().numpy().astype(np.float64)
for i, nm in [(1, "nms #1 (256 boxes, pick order)"), (2, "pool staging + slab masks"), (3, "pool match"), (4, "nms #2 (512 pool boxes)")]:
    print("%-32s %7.2f us" % (nm, np.median(a[:, i] - a[:, i - 1]) / 1e3))
print("CTA life %.1f us; picks per scene %.0f" % (np.median(a[:, 4] - a[:, 0]) / 1e3, a[:, 5].mean()))
