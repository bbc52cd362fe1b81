"""Per-tile stage stamps of the persistent logits kernel (config 4) + graph-timed launch.\n    [OVDET_LOGITS_PREFETCH=1] python profiles/logits_stamps.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
dev = torch.device("cuda")
dbg = torch.zeros((148, 8, 8), dtype=torch.int64, device=dev)
os.environ["OVDET_LOGITS_DBG_PTR"] = str(dbg.data_ptr())
import ovdet_b200
from ovdet_b200 import synth
from ovdet_b200.models.model_3detr import clip_logits
x, t = synth.clip_logits_inputs(8192, 640, 1203)
xd, td = (x * 0.25).to(dev), t.to(dev)
for _ in range(3): clip_logits(xd, td)
torch.cuda.synchronize(); dbg.zero_(); torch.cuda.synchronize()
clip_logits(xd, td); torch.cuda.synchronize()
d = dbg.cpu().numpy().astype(np.float64); d = d[d[:, 7, 0] > 0]
t0 = d[:, 0, 0].min()
names = ["wait tmem_full", "TMEM pass (exp -> registers)", "merge+publish", "wait stats", "row sums", "rescale+store", "-"]
for it in range(4):
    v = d[:, it]
    ok = v[:, 7] > 0
    if not ok.any(): continue
    print("tile %d: n=%d start %.2f end %.2f |" % (it, ok.sum(), (v[ok, 0].mean() - t0) / 1e3, (v[ok, 7].mean() - t0) / 1e3),
          " ".join("%s %.2f" % (names[i], ((v[ok, i + 1] - v[ok, i]) / 1e3).mean()) for i in range(7)))
e = d[:, 7]
T0 = e[:, 0].min()
print("entry spread %.2f us; prologue mean %.2f us; kernel span (first entry -> last exit) %.2f us; mean CTA life %.2f" % (
    (e[:, 0].max() - T0) / 1e3, ((e[:, 2] - e[:, 0]) / 1e3).mean(), (e[:, 1].max() - T0) / 1e3, ((e[:, 1] - e[:, 0]) / 1e3).mean()))
print("first epilogue stamp relative to first entry: %.2f us" % ((d[:, 0, 0].min() - T0) / 1e3))
# back-to-back launches timed with events
def tm(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
os.environ.pop("OVDET_LOGITS_DBG_PTR")
print("graph-timed %.1f us" % tm(lambda: clip_logits(xd, td)))
