"""Per-CTA phase times of ap_front2_kernel from globaltimer stamps (OVDET_APFRONT_DBG_PTR)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from ovdet_b200.utils import ap_calculator as APC, eval_det as ED
S = int(sys.argv[1]) if len(sys.argv) > 1 else 5050
dev = torch.device("cuda", 0)
out, tgt = bench.ap_inputs(S)
dv = {k: v.to(dev).contiguous() for k, v in {**out, **tgt}.items()}
cfg = APC.get_ap_config_dict(dataset_config=bench._Cfg(), remove_empty_box=False)
lists = ED.TpLists(20, dev)
f = lambda: ED.ap_front(dv["box_corners"], dv["sem_cls_prob"], dv["objectness_prob"], None, dv["gt_box_corners"], dv["gt_box_sem_cls_label"], dv["gt_box_present"], 20, [0.25, 0.5], cfg, lists)
f(); torch.cuda.synchronize()
d1 = torch.zeros((S, 16), dtype=torch.int64, device=dev)
os.environ["OVDET_APFRONT_DBG_PTR"] = str(d1.data_ptr())
f(); torch.cuda.synchronize()
a = d1.cpu().numpy().astype(np.float64)
t0 = a[:, 0]
prev = t0
for i, nm in [(7, "init+box"), (8, "probs+records"), (9, "GT"), (1, "barrier B1"), (10, "nms pairs"), (11, "nms rounds"), (2, "fixup"), (3, "enumerate"), (4, "clip"), (5, "claims")]:
    ok = a[:, i] > 0
    print("%-12s %7.2f us" % (nm, np.median((a[:, i] - prev)[ok]) / 1e3)); prev = np.where(ok, a[:, i], prev)
print("CTA life median %.2f us, p95 %.2f; kernel span %.1f us" % (np.median(prev - t0) / 1e3, np.percentile(prev - t0, 95) / 1e3, (prev.max() - t0.min()) / 1e3))
qn = d1.cpu().numpy()[:, 6]
print("candidates per scene: mean %.1f max %d" % (qn.mean(), qn.max()))
