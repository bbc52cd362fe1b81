"""cProfile of the host side of one AP evaluation (reset + step + compute_metrics)."""
import os, sys, cProfile, pstats, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ovdet_b200.utils import ap_calculator as APC
S = int(sys.argv[1]) if len(sys.argv) > 1 else 631
dev = torch.device("cuda", 0)
out, tgt = bench.ap_inputs(S)
dv = {k: v.to(dev).contiguous() for k, v in {**out, **tgt}.items()}
calc = APC.APCalculator(bench._Cfg(), ap_iou_thresh=[0.25, 0.5], exact_eval=False)
def run():
    calc.reset()
    calc.step(dv["box_corners"], dv["sem_cls_prob"], dv["objectness_prob"], None, dv["gt_box_corners"], dv["gt_box_sem_cls_label"], dv["gt_box_present"])
    return calc.compute_metrics()
for _ in range(20):
    run()
t0 = time.perf_counter()
for _ in range(200):
    run()
print("wall us/eval", (time.perf_counter() - t0) / 200 * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    run()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
