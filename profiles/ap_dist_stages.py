"""Per-rank stage timing of the scene-sharded AP evaluation under torchrun (front / reducer chain incl. waiting for the
peers / host), strong (5050 scenes in total) and weak (5050 per rank)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
from ovdet_b200.utils import ap_calculator as APC
from ovdet_b200 import dist as D
rank, world, local = bench.dist_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
S = 5050
out, tgt = bench.ap_inputs(S)
allin = {**out, **tgt}
for mode in ("strong", "weak"):
    if mode == "strong":
        lo, hi = D.shard_range(S, rank, world)
        dv = {k: v[lo:hi].to(dev).contiguous() for k, v in allin.items()}
    else:
        o2, t2 = bench.ap_inputs(S, seed=1000 + 7919 * rank)
        dv = {k: v.to(dev).contiguous() for k, v in {**o2, **t2}.items()}
    calc = APC.APCalculator(bench._Cfg(), ap_iou_thresh=[0.25, 0.5], exact_eval=False)
    def run(rec=None):
        t0 = time.perf_counter()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        calc.reset()
        e[0].record()
        calc.step(dv["box_corners"], dv["sem_cls_prob"], dv["objectness_prob"], None, dv["gt_box_corners"], dv["gt_box_sem_cls_label"], dv["gt_box_present"])
        e[1].record()
        t1 = time.perf_counter()
        m = calc.compute_metrics(distributed=True)
        t2 = time.perf_counter()
        e[2].record(); torch.cuda.synchronize()
        if rec is not None:
            rec.append((e[0].elapsed_time(e[1]) * 1e3, e[1].elapsed_time(e[2]) * 1e3, (t1 - t0) * 1e6, (t2 - t1) * 1e6, (t2 - t0) * 1e6))
        return m
    for _ in range(5):
        run()
    dist.barrier(); torch.cuda.synchronize()
    rec = []
    t0 = time.perf_counter()
    for _ in range(30):
        run(rec)
    wall = (time.perf_counter() - t0) / 30 * 1e6
    a = np.median(np.array(rec), 0)
    msg = "%s rank %d: gpu front %.0f us, gpu reduce(+peer wait) %.0f us | host step %.0f us, compute_metrics %.0f us, eval %.0f us, loop wall %.0f us" % (mode, rank, a[0], a[1], a[2], a[3], a[4], wall)
    gathered = [None] * world
    dist.all_gather_object(gathered, msg)
    if rank == 0:
        print("\n".join(gathered))
    # device-side stamps of one more evaluation (globaltimer of this GPU; all 20 classes)
    d = torch.zeros((20, 32), dtype=torch.int64, device=dev)
    os.environ["OVDET_APX_DBG_PTR"] = str(d.data_ptr())
    dist.barrier(); run(); torch.cuda.synchronize()
    os.environ.pop("OVDET_APX_DBG_PTR")
    a = d.cpu().numpy().astype(np.float64) / 1e3
    t0 = a[:, 8].min()
    f = lambda col, fn=np.median: fn(a[:, col]) - t0
    msg = ("%s rank %d stamps (us after push start): push pdl-wait done %.0f, sorted %.0f, pushed %.0f | merge start %.0f, flags in %.0f (max %.0f), counts %.0f, merged %.0f, bins %.0f, end %.0f | "
           "hist ship begin %.0f (max %.0f), shipped %.0f (max %.0f) | final start %.0f, after pdl %.0f, flags in %.0f (max %.0f), end %.0f (max %.0f)" %
           (mode, rank, f(9), f(10), f(11), f(0), f(6), f(6, np.max), f(2), f(3), f(4), f(5), f(12), f(12, np.max), f(13), f(13, np.max), f(16), f(17), f(18), f(18, np.max), f(19), f(19, np.max)))
    gathered = [None] * world
    dist.all_gather_object(gathered, msg)
    if rank == 0:
        print("\n".join(gathered[:3]))
    calc.close()
dist.barrier()
dist.destroy_process_group()
