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
calc.force_exchange = bool(os.environ.get("APX_FORCE"))   # one rank through the push / cluster-merge / ship code path

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
    # stamp ids (ap_compact.cu XSTAMPC): push 8 start, 9 after the dependency wait, 10 sorted, 11 flags raised;
    # cluster merge 0 start, 6 flags seen, 2 runs in shared memory, 3 ranked + stored, 5 bins done;
    # hist ship 12/13; final 16 start, 17 after the dependency wait, 18 histograms summed, 19 done
    t0 = a[:, 8].min()
    for i, nm in [(8, "push start"), (9, "push dep wait"), (10, "push sorted"), (11, "push flags"), (0, "merge start"), (6, "merge flags seen"),
                  (2, "merge runs in smem"), (3, "merge ranked"), (5, "merge bins"), (16, "final start"), (17, "final dep wait"),
                  (18, "final summed"), (19, "final done")]:
        col = a[:, i][a[:, i] > 0]
        if len(col):
            print("%-20s median %7.2f  max %7.2f us after the first push CTA" % (nm, (np.median(col) - t0) / 1e3, (col.max() - t0) / 1e3))
