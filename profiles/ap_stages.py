"""Stage timing of the AP evaluation (config 3): fused front end, exchange reducer, host glue.
    python profiles/ap_stages.py [n_scenes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from ovdet_b200.utils import ap_calculator as APC, eval_det as ED

S = int(sys.argv[1]) if len(sys.argv) > 1 else 5050
dev = torch.device("cuda", 0)
out, tgt = bench.ap_inputs(S)
dv = {k: v.to(dev).contiguous() for k, v in {**out, **tgt}.items()}
calc = APC.APCalculator(bench._Cfg(), ap_iou_thresh=[0.25, 0.5], exact_eval=False)

def run():
    calc.reset()
    calc.step(dv["box_corners"], dv["sem_cls_prob"], dv["objectness_prob"], None, dv["gt_box_corners"], dv["gt_box_sem_cls_label"], dv["gt_box_present"])
    return calc.compute_metrics()

for _ in range(3):
    m = run()
def ev(fn, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
cfg = calc.ap_config_dict
lists = calc._lists
def front():
    lists.reset()
    return ED.ap_front(dv["box_corners"], dv["sem_cls_prob"], dv["objectness_prob"], None, dv["gt_box_corners"], dv["gt_box_sem_cls_label"], dv["gt_box_present"], 20, [0.25, 0.5], cfg, lists, iou_ws=calc._iou_ws)
print("front us", ev(front))
rs = front()[0]
red = list(calc._reducers.values())[-1]
print("reduce us (cap %d)" % red.cap_total, ev(lambda: red.launch([rs], lists)))
t0 = time.perf_counter()
for _ in range(50):
    run()
torch.cuda.synchronize()
print("wall us per evaluation", (time.perf_counter() - t0) / 50 * 1e6)
# host-only cost: the same calls with the GPU work already queued is hard to isolate; time step() and compute_metrics() separately
t0 = time.perf_counter()
for _ in range(50):
    calc.reset()
    calc.step(dv["box_corners"], dv["sem_cls_prob"], dv["objectness_prob"], None, dv["gt_box_corners"], dv["gt_box_sem_cls_label"], dv["gt_box_present"])
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host us per reset+step (async)", (t1 - t0) / 50 * 1e6)
print("mAP", float(m[0.25]["mAP"]), float(m[0.5]["mAP"]))
if os.environ.get("APX_STAMPS"):
    d = torch.zeros((20, 32), dtype=torch.int64, device=dev)
    os.environ["OVDET_APX_DBG_PTR"] = str(d.data_ptr())
    red.launch([rs], lists); torch.cuda.synchronize()
    a = d.cpu().numpy().astype(np.float64)
    print("merge: gather issue %.2f, gather wait %.2f, rank+rest %.2f us; total entries %s" % (np.median(a[:, 6] - a[:, 2]) / 1e3, np.median(a[:, 7] - a[:, 6]) / 1e3, np.median(a[:, 3] - a[:, 7]) / 1e3, lists.tp_cnt.cpu().numpy()[:6]))
    for i, nm in [(1, "pdl wait"), (2, "counts (thread 0)"), (3, "gather+sort"), (4, "bin search"), (5, "table+zero")]:
        print("merge %-18s %6.2f us" % (nm, np.median(a[:, i] - a[:, i - 1]) / 1e3))
