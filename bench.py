#!/usr/bin/env python
"""bench.py -- headline benchmark of the detection-geometry hot path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm  (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU arm

Headline (BASELINE.json metric "3D GIoU pairs/s & AP-eval scenes/s"):
  step      one `generalized_box3d_iou` pass of a SUN RGB-D-shaped training step
            (BASELINE config 1): the 8 decoder layers x batch 8 = 64 box sets,
            128 queries x 64 padded GT, rotated boxes, reference default semantics
            (Cython path as shipped: fp64 clip, prefilter, K2<=4 column cap) = 524 288 pairs.
  value     pairs/s with inputs resident in HBM (CUDA events on the launch stream,
            rotating input sets larger than L2), max over ranks.
  e2e       the same call through the reference-facing API with HOST buffers:
            H2D of the corners from pinned memory, kernel, D2H of the [64,128,64] result, every step.
  ap_eval   BASELINE config 3: NMS + per-class AP over 5050 synthetic scenes (128 queries,
            20 classes, IoU 0.25/0.5), scene-sharded across the ranks with one NCCL
            all-gather of the (score, tp) records.
  extra     the other BASELINE configs (matcher step, logits GEMM, pseudo-label sweep), short runs.
Nothing here reads /root/reference.  oracle/ is used only for the cpu_baseline legs and
the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

L_LAYERS, B, Q, G, C_SUN = 8, 8, 128, 64, 20
GIOU_BYTES_PER_PAIR = (96 * (Q + G) + 8 + 4 * Q * G) / (Q * G)   # 6.2510 B/pair at config 1


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        hi = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(hi)) if hi else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def max_over_ranks(x, dev, world):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def barrier_sync(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def timed_region(fn, steps, world, dev):
    """EXACTLY `steps` calls of fn(i) bracketed by barrier + synchronize, CUDA events, max over ranks -> ms total."""
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    barrier_sync(world)
    return max_over_ranks(ms, dev, world)


def timed_graph(fn, steps, world, dev):
    """Same contract as timed_region, but the `steps` calls are captured once into a CUDA graph and the replay is
    timed, so the figure is the GPU's, not the Python launch loop's (each call is ~10-20 us of kernel against a
    similar amount of host-side shim work)."""
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for i in range(3):
            fn(i)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(steps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    return timed_region(lambda i: g.replay(), 1, world, dev)


# ----------------------------------------------------------------------------- GIoU (headline)
def giou_inputs(seed, heading=np.pi):
    from ovdet_b200 import synth
    out, tgt = synth.detection_batch(B=L_LAYERS * B, Q=Q, G=G, C=C_SUN, seed=seed, heading=heading)
    return out, tgt


def bench_giou(args, rank, world, dev, peaks):
    from ovdet_b200.utils.box_util import generalized_box3d_iou
    nsets = 48   # 48 x (3.1 MB in + 2.1 MB out) = 250 MB > 126 MB L2: every step misses L2
    sets = []
    for s in range(4):
        out, tgt = giou_inputs(100 + s)
        sets.append((out["box_corners"], tgt["gt_box_corners"], tgt["nactual_gt"]))
    dsets = []
    for s in range(nsets):
        c1, c2, nk = sets[s % 4]
        dsets.append((c1.to(dev).clone(), c2.to(dev).clone(), nk.to(dev).clone(),
                      torch.empty((L_LAYERS * B, Q, G), dtype=torch.float32, device=dev)))
    pairs = L_LAYERS * B * Q * G

    def step(i, **kw):
        c1, c2, nk, o = dsets[i % nsets]
        generalized_box3d_iou(c1, c2, nk, rotated_boxes=True, out=o, **kw)

    for i in range(args.warmup):
        step(i)
    ms_loop = timed_region(step, args.steps, world, dev)    # Python launch loop (host-bound at this kernel size)
    ms = timed_graph(step, args.steps, world, dev)          # the same K launches replayed from a CUDA graph
    value = world * pairs * args.steps / (ms * 1e-3)
    per_launch_s = ms * 1e-3 / args.steps
    algo_bytes = GIOU_BYTES_PER_PAIR * pairs
    achieved = algo_bytes / per_launch_s / 1e9
    res = {"value": value, "ms_per_step": ms / args.steps, "launches": args.steps, "ms_per_step_python_loop": ms_loop / args.steps,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm_gbs"], "traffic": load_traffic("giou3d_kernel<double"),
                        "kernel": "giou3d_kernel<double> (reference-default semantics)", "peak_source": peaks["source"],
                        "note": "pair maths is ALU/issue-bound (SURVEY 8d); see profiles/ for pipe utilisation"}}
    # intended semantics (no K2 cap, fp32 clip = the torch path): every prefilter-passing pair is clipped
    for i in range(3):
        step(i, mode="tensor", k2_cap=0)
    ms2 = timed_graph(lambda i: step(i, mode="tensor", k2_cap=0), args.steps, world, dev)
    res["variants"] = {"tensor_nocap_pairs_per_s": world * pairs * args.steps / (ms2 * 1e-3)}
    for i in range(3):
        step(i, mode="tensor", k2_cap=0, prefilter=False)
    ms3 = timed_graph(lambda i: step(i, mode="tensor", k2_cap=0, prefilter=False), max(args.steps // 4, 5), world, dev)
    res["variants"]["exact_noprefilter_pairs_per_s"] = world * pairs * max(args.steps // 4, 5) / (ms3 * 1e-3)

    # the |heading| <= 0.5 variant (many more pairs pass the axis-aligned prefilter and are clipped)
    o5, t5 = giou_inputs(200, heading=0.5)
    h5 = (o5["box_corners"].to(dev), t5["gt_box_corners"].to(dev), t5["nactual_gt"].to(dev), torch.empty((L_LAYERS * B, Q, G), device=dev))
    for kw, name in ((dict(), "heading05_default_pairs_per_s"), (dict(mode="tensor", k2_cap=0), "heading05_tensor_nocap_pairs_per_s")):
        f5 = lambda i: generalized_box3d_iou(h5[0], h5[1], h5[2], rotated_boxes=True, out=h5[3], **kw)
        for i in range(3):
            f5(i)
        ms5 = timed_graph(f5, max(args.steps // 4, 5), world, dev)
        res["variants"][name] = world * pairs * max(args.steps // 4, 5) / (ms5 * 1e-3)
    # config 2 shape: ScanNet, 256 queries, axis-aligned boxes (no clipping at all: one fp32 result per ~70 flops)
    from ovdet_b200 import synth as _synth
    o2, t2 = _synth.detection_batch(B=L_LAYERS * B, Q=256, G=G, C=18, seed=300, room="scannet", heading=0.0)
    a2 = (o2["box_corners"].to(dev), t2["gt_box_corners"].to(dev), t2["nactual_gt"].to(dev), torch.empty((L_LAYERS * B, 256, G), device=dev))
    f2_ = lambda i: generalized_box3d_iou(a2[0], a2[1], a2[2], rotated_boxes=False, out=a2[3])
    for i in range(3):
        f2_(i)
    msa = timed_graph(f2_, max(args.steps // 4, 5), world, dev)
    res["variants"]["axis_aligned_c2_pairs_per_s"] = world * L_LAYERS * B * 256 * G * max(args.steps // 4, 5) / (msa * 1e-3)
    # launch floor of this timing method: the same entry point on a 1x1x1 problem (one CTA, ~no work)
    t1 = (torch.zeros((1, 1, 8, 3), device=dev), torch.zeros((1, 1, 8, 3), device=dev), torch.ones((1,), dtype=torch.int64, device=dev),
          torch.empty((1, 1, 1), device=dev))
    msf = timed_graph(lambda i: generalized_box3d_iou(t1[0], t1[1], t1[2], out=t1[3]), args.steps, world, dev)
    res["variants"]["launch_floor_us"] = msf * 1e3 / args.steps
    # training path (needs_grad=True): torch-path forward + sparse backward, upstream gradient at one GT per query
    c1g, c2g, nkg, _ = dsets[0]
    wg = torch.zeros((L_LAYERS * B, Q, G), device=dev)
    wg.scatter_(2, torch.randint(0, G, (L_LAYERS * B, Q, 1), device=dev), 1.0)
    xg = c1g.clone().requires_grad_(True)

    def train_step(i):
        xg.grad = None
        (generalized_box3d_iou(xg, c2g, nkg, rotated_boxes=True, needs_grad=True) * wg).sum().backward()

    for i in range(3):
        train_step(i)
    msg = timed_region(train_step, max(args.steps // 4, 5), world, dev)
    res["variants"]["train_fwd_bwd_us"] = msg * 1e3 / max(args.steps // 4, 5)
    # throughput regime: 4096 box sets (33.5 M pairs, 240 MB in+out) in one launch, torch-path semantics
    big = 4096
    rep = big // (L_LAYERS * B)
    c1b, c2b, nkb = dsets[0][0].repeat(rep, 1, 1, 1), dsets[0][1].repeat(rep, 1, 1, 1), dsets[0][2].repeat(rep)
    ob = torch.empty((big, Q, G), dtype=torch.float32, device=dev)
    for kw, name in ((dict(), "large_batch_default"), (dict(mode="tensor", k2_cap=0), "large_batch_tensor_nocap")):
        f = lambda i: generalized_box3d_iou(c1b, c2b, nkb, rotated_boxes=True, out=ob, **kw)
        f(0)
        msb = timed_region(f, 10, world, dev)
        pb = big * Q * G
        res["variants"][name] = {"pairs_per_s": world * pb * 10 / (msb * 1e-3), "ms": msb / 10,
                                 "hbm_frac": GIOU_BYTES_PER_PAIR * pb / (msb * 1e-3 / 10) / 1e9 / peaks["hbm_gbs"]}
    del c1b, c2b, ob

    # ---- e2e: host buffers in pinned memory, H2D + kernel + D2H every step (the reference call ends in .cpu())
    c1, c2, nk = sets[0]
    hsets = []
    for s in range(4):
        a, b_, n_ = sets[s]
        hsets.append((a.clone().pin_memory(), b_.clone().pin_memory(), n_.clone().pin_memory(),
                      torch.empty((L_LAYERS * B, Q, G), dtype=torch.float32).pin_memory()))

    def estep(i):
        a, b_, n_, o = hsets[i % 4]
        generalized_box3d_iou(a, b_, n_, rotated_boxes=True, out=o)

    for i in range(max(3, args.warmup)):
        estep(i)
    esteps = max(args.steps, 5)     # every call ends in a stream synchronise inside the entry point: wall clock is exact
    barrier_sync(world)
    t0 = time.perf_counter()
    for i in range(esteps):
        estep(i)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0, dev, world)
    h2d = c1.numel() * 4 + c2.numel() * 4 + nk.numel() * 8
    d2h = L_LAYERS * B * Q * G * 4
    res["e2e"] = {"value": world * pairs * esteps / dt, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "ms_per_step": dt * 1e3 / esteps, "steps": esteps, "api": "generalized_box3d_iou(cpu pinned tensors) -> ovdet_giou3d_host_f32"}
    return res


def load_traffic(kernel):
    """dram bytes per launch from the committed ncu summary (profiles/ncu_summary.json), else null."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k, v in d.items():           # keys carry the template arguments: first launch whose name starts with `kernel`
                if k.startswith(kernel):
                    return v.get("dram_bytes_per_launch")
            return None
        except Exception:
            return None
    return None


def cpu_giou_baseline(steps=None, budget_s=10.0):
    """The reference's own compiled Cython loop (oracle/_ref) + restated torch glue, as shipped, on one decoder
    layer's batch (8 x 128 x 64 = 65 536 nominal pairs); falls back to the C port when _ref is absent.
    Bounded sample: repeated for ~budget_s seconds of CPU work (or exactly `steps` calls), median per call."""
    import oracle
    out, tgt = giou_inputs(100)
    c1, c2, nk = out["box_corners"][:B], tgt["gt_box_corners"][:B], tgt["nactual_gt"][:B]
    use_ref = oracle.ref_box_intersection() is not None
    fn = (lambda: oracle.generalized_box3d_iou_ref_cython(c1, c2, nk, True, False)) if use_ref else \
         (lambda: oracle.generalized_box3d_iou(c1, c2, nk, True, False, mode="cython", k2_cap=4))
    fn()
    ts = []
    t_start = time.perf_counter()
    while (len(ts) < steps) if steps is not None else (time.perf_counter() - t_start < budget_s and len(ts) < 20000):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    steps = len(ts)
    t = float(np.median(ts))
    res = {"value": B * Q * G / t, "unit": "pairs/s", "cores": 1, "kind": "reference" if use_ref else "port",
           "sample": "one decoder layer (8x128x64 = 65 536 nominal pairs) of the step, as shipped (K2<=4 cap); "
                     + ("hot loop = reference box_intersection.pyx compiled in oracle/_ref, torch glue restated" if use_ref
                        else "oracle C port (oracle/_ref absent)") + "; median of %d calls (%.1f s of CPU work)" % (steps, sum(ts)),
           "ms_per_sample": t * 1e3, "torch_threads": torch.get_num_threads(), "host_cpus": os.cpu_count()}
    if use_ref:  # the intended (no-cap) semantics through the unmodified extension, 4 GT columns per call
        t0 = time.perf_counter()
        oracle.generalized_box3d_iou_ref_cython(c1[:2], c2[:2], nk[:2], True, False, lift_k2_cap=True)
        res["nocap_pairs_per_s"] = 2 * Q * G / (time.perf_counter() - t0)
    return res


# ----------------------------------------------------------------------------- AP evaluation (config 3)
class _Cfg:
    num_semcls = C_SUN


def ap_inputs(n_scenes, seed=7):
    from ovdet_b200 import synth
    outs, tgts = [], []
    chunk = 505
    for s0 in range(0, n_scenes, chunk):
        n = min(chunk, n_scenes - s0)
        o, t = synth.detection_batch(B=n, Q=Q, G=G, C=C_SUN, seed=seed + s0, heading=np.pi, max_gt=12)
        outs.append(o); tgts.append(t)
    cat = lambda key, src: torch.cat([d[key] for d in src], 0)
    return ({k: cat(k, outs) for k in ("box_corners", "sem_cls_prob", "objectness_prob")},
            {k: cat(k, tgts) for k in ("gt_box_corners", "gt_box_sem_cls_label", "gt_box_present")})


def bench_ap(args, rank, world, dev, peaks):
    from ovdet_b200.utils import ap_calculator as APC
    from ovdet_b200 import dist as D
    S = 5050
    out, tgt = ap_inputs(S)
    lo, hi = D.shard_range(S, rank, world)
    dv = {k: v[lo:hi].to(dev).contiguous() for k, v in {**out, **tgt}.items()}
    hv = {k: v[lo:hi].contiguous().pin_memory() for k, v in {**out, **tgt}.items()}
    calc = APC.APCalculator(_Cfg(), ap_iou_thresh=[0.25, 0.5], exact_eval=False)

    def run(src):
        calc.reset()
        calc.step(src["box_corners"], src["sem_cls_prob"], src["objectness_prob"], None, src["gt_box_corners"],
                  src["gt_box_sem_cls_label"], src["gt_box_present"])
        return calc.compute_metrics(distributed=world > 1)

    steps = max(3, min(args.steps // 10, 20))
    for _ in range(3):
        m = run(dv)
    ms = timed_region(lambda i: run(dv), steps, world, dev)
    val = S * steps / (ms * 1e-3)
    rec_bytes = S * Q * C_SUN * 5 * 2 * 5   # (4 B score + 1 B tp) x read+write x (prep + 4 radix passes), SURVEY 8d stream
    res = {"value": val, "unit": "scenes/s", "ms_per_step": ms / steps, "steps": steps, "scaling": "strong",
           "mAP_0.25": float(m[0.25]["mAP"]), "mAP_0.5": float(m[0.5]["mAP"]),
           "record_stream_GBs": rec_bytes / (ms * 1e-3 / steps) / 1e9,
           "workload": "5050 scenes x 128 queries x 20 classes, NMS 0.25 + AP@0.25/0.5, %d scenes on this rank" % (hi - lo)}
    # e2e: host tensors (pinned) -> device every step, metrics dict read back
    def erun(i):
        src = {k: v.to(dev, non_blocking=True) for k, v in hv.items()}
        return run(src)
    erun(0)
    barrier_sync(world)
    t0 = time.perf_counter()
    for i in range(3):
        erun(i)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0, dev, world)
    res["e2e"] = {"value": S * 3 / dt, "unit": "scenes/s", "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in hv.values())),
                  "d2h_bytes_per_step": 2 * 2 * C_SUN * 8}
    if world > 1:
        # weak scaling: 5050 scenes PER RANK (different seeds), same exchange -- the regime where sharding pays;
        # the strong-scaling figure above is latency-limited (the whole 5050-scene evaluation is ~1 ms on one GPU)
        out_w, tgt_w = ap_inputs(S, seed=1000 + 7919 * rank)
        dw = {k: v.to(dev).contiguous() for k, v in {**out_w, **tgt_w}.items()}
        for _ in range(2):
            run(dw)
        msw = timed_region(lambda i: run(dw), steps, world, dev)
        res["weak"] = {"value": world * S * steps / (msw * 1e-3), "unit": "scenes/s", "ms_per_step": msw / steps,
                       "scaling": "weak", "workload": "5050 scenes per rank"}
    return res


def cpu_ap_baseline(n_scenes=1024):
    import oracle
    out, tgt = ap_inputs(n_scenes)
    t0 = time.perf_counter()
    oracle.ap_metrics(out["box_corners"], out["sem_cls_prob"], out["objectness_prob"], tgt["gt_box_corners"],
                      tgt["gt_box_sem_cls_label"], tgt["gt_box_present"], C_SUN)
    dt = time.perf_counter() - t0
    return {"value": n_scenes / dt, "unit": "scenes/s", "cores": 1, "kind": "port",
            "sample": "%d of the 5050 scenes through the oracle's C/numpy port of parse_predictions + eval_det (the Python "
                      "reference itself, Pool(10), measured 9.5 scenes/s at survey time, SURVEY.md 6)" % n_scenes}


# ----------------------------------------------------------------------------- extras
def bench_extras(args, rank, world, dev, peaks):
    from ovdet_b200 import synth
    from ovdet_b200.criterion import Matcher
    from ovdet_b200.models.model_3detr import clip_logits
    from ovdet_b200.utils.box_3d_utils import lift_filter_batch
    ex = {}
    n = max(20, args.steps // 4)
    # config 2: ScanNet-shaped matcher step (GIoU + L1 centre + class + objectness cost, then LSAP), 8 layers batched
    out, tgt = synth.detection_batch(B=L_LAYERS * B, Q=256, G=G, C=18, seed=3, room="scannet", heading=0.0)
    o = {k: v.to(dev) for k, v in out.items()}
    t = {k: v.to(dev) for k, v in tgt.items()}
    m = Matcher(1, 0, 2, 0)
    f = lambda i: m.match_from_boxes(o, t, rotated_boxes=False, return_assignments=False)
    for i in range(5):
        f(i)
    ms = timed_graph(f, n, world, dev)
    ex["matcher_scannet"] = {"pairs_per_s": world * L_LAYERS * B * 256 * G * n / (ms * 1e-3), "ms_per_step": ms / n,
                             "workload": "8 layers x B8 x 256 queries x 64 GT, fused cost kernel + on-device LSAP"}
    # config 4: open-vocab logits
    x, tx = synth.clip_logits_inputs(8192, 640, 1203)
    xd, td = (x * 0.25).to(dev), tx.to(dev)
    f = lambda i: clip_logits(xd, td)
    for i in range(5):
        f(i)
    ms = timed_graph(f, n, world, dev)
    fl = 2 * 8192 * 640 * 1203
    ach = fl / (ms * 1e-3 / n) / 1e12
    ex["clip_logits"] = {"tflops": ach, "ms_per_step": ms / n, "frac_of_bf16_peak": ach / peaks["bf16_tflops"],
                         "roofline": {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                      "frac": ach / peaks["bf16_tflops"], "traffic": load_traffic("clip_logits_persistent_kernel")},
                         "workload": "8192x640 @ 1203x640^T bf16 -> softmax probs bf16 + objectness (includes bf16 cast-free path)"}
    # the reference's only native ABI, box_intersection(rect1, rect2, nonrot, nums_k2, inter_areas, approximate) on HOST
    # numpy buffers (box_intersection.pyx:166-171): the drop-in against the reference's own compiled extension, same call
    from ovdet_b200.utils.box_intersection import box_intersection as bi_ours
    o1, t1 = synth.detection_batch(B=B, Q=Q, G=G, C=C_SUN, seed=100, heading=np.pi)
    r1 = np.ascontiguousarray(o1["box_corners"][:, :, [3, 2, 1, 0]][..., [0, 2]].numpy(), dtype=np.float32)
    r2 = np.ascontiguousarray(t1["gt_box_corners"][:, :, [3, 2, 1, 0]][..., [0, 2]].numpy(), dtype=np.float32)
    lt = np.maximum(r1[:, :, None, 1], r2[:, None, :, 1]); rb = np.minimum(r1[:, :, None, 3], r2[:, None, :, 3])
    wh = np.clip(rb - lt, 0, None)
    nonrot = np.ascontiguousarray(wh[..., 0] * wh[..., 1], dtype=np.float32)
    nk32 = t1["nactual_gt"].numpy().astype(np.int32)
    areas = np.zeros_like(nonrot)
    for _ in range(5):
        bi_ours(r1, r2, nonrot, nk32, areas, True)
    t0 = time.perf_counter()
    for _ in range(50):
        bi_ours(r1, r2, nonrot, nk32, areas, True)
    t_ours = (time.perf_counter() - t0) / 50
    abi = {"ours_us_per_call": t_ours * 1e6, "calls": 50, "nominal_pairs": B * Q * G,
           "workload": "box_intersection on host numpy buffers, one decoder layer (8x128x64), as shipped (k2 < 4), approximate=True"}
    if not args.skip_cpu and rank == 0:
        import oracle
        ref = oracle.ref_box_intersection()
        if ref is not None:
            a2 = np.zeros_like(nonrot)
            ref(r1, r2, nonrot, nk32, a2, True)
            t0 = time.perf_counter()
            for _ in range(20):
                ref(r1, r2, nonrot, nk32, a2, True)
            abi["reference_cython_us_per_call"] = (time.perf_counter() - t0) / 20 * 1e6
            abi["identical_output"] = bool(np.array_equal(a2, areas))
    ex["box_intersection_abi"] = abi
    # class-wise 3D NMS alone (10.9 ms/scene in the reference at K = 256)
    from ovdet_b200.utils.nms import nms_batch
    gN = torch.Generator().manual_seed(11)
    SN, KN = 4096, 256
    cN, sN, _ = synth.sample_boxes(gN, (SN, KN), "scannet", 0.0)
    bN = torch.cat([(cN - sN / 2).double(), (cN + sN / 2).double(), torch.rand((SN, KN, 1), generator=gN).double(),
                    torch.randint(0, 18, (SN, KN, 1), generator=gN).double()], -1).to(dev)
    f = lambda i: nms_batch(bN, 0.25, samecls=True, want_order=False)
    for i in range(2):
        f(i)
    ms = timed_graph(f, 5, world, dev)
    ex["nms3d_samecls"] = {"scenes_per_s": world * SN * 5 / (ms * 1e-3), "ms_per_step": ms / 5,
                           "workload": "%d scenes/rank x 256 boxes x 18 classes, nms_3d_faster_samecls thr 0.25 (keep mask)" % SN}
    # config 5: pseudo-label sweep, scene-sharded, no data-path collective
    S5 = 4096
    bx, pool = synth.pseudo_label_scenes(S5, P=256, pool=512, seed=5)
    bxd, pd = bx.to(dev), pool.to(dev)
    f = lambda i: lift_filter_batch(bxd, pd)
    for i in range(2):
        f(i)
    k = 5
    ms = timed_graph(f, k, world, dev)
    ex["pseudo_label"] = {"scenes_per_s": world * S5 * k / (ms * 1e-3), "ms_per_step": ms / k,
                          "workload": "%d scenes/rank x 256 proposals x 512 pool boxes: NMS 0.7 -> IoU>=0.3 match -> size NMS" % S5}
    return ex


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    # all the host threads the reference can use: torchrun exports OMP_NUM_THREADS=1, which would cripple its torch glue
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    base = cpu_giou_baseline(steps=max(1, args.steps))
    line = {"impl": "reference", "metric": "3D GIoU pairs/s (SUN RGB-D-shaped step) & AP-eval scenes/s", "value": base["value"],
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": base["ms_per_sample"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "generalized_box3d_iou, rotated, reference default (Cython path as shipped); each step = "
                                   "one decoder layer 8x128x64 of the 8-layer step (bounded sample)", "B": B, "Q": Q, "G": G},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ap_eval": cpu_ap_baseline(48)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline legs (profiling runs)")
    ap.add_argument("--only", default="", help="comma list of sections: giou,ap,extras")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import ovdet_b200  # noqa: F401
    peaks = load_peaks()
    only = set(x for x in args.only.split(",") if x)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    g = bench_giou(args, rank, world, dev, peaks)
    clocks = sampler.stop() if rank == 0 else None
    apres = bench_ap(args, rank, world, dev, peaks) if (not only or "ap" in only) else None
    extras = bench_extras(args, rank, world, dev, peaks) if (not only or "extras" in only) else None
    if rank == 0:
        cpu = None if args.skip_cpu else cpu_giou_baseline()
        if apres is not None and not args.skip_cpu:
            apres["cpu_baseline"] = cpu_ap_baseline()
        line = {
            "metric": "3D GIoU pairs/s (SUN RGB-D-shaped step) & AP-eval scenes/s",
            "value": g["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": g["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "timing": "K launches captured in one CUDA graph, replay timed with CUDA events (ms_per_step_python_loop = same launches from Python)",
            "config": {"workload": "generalized_box3d_iou over one SUN RGB-D-shaped training step: 8 decoder layers x batch 8 = 64 "
                                   "box sets, 128 queries x 64 padded GT (nactual~U{1..64}), rotated, reference default semantics "
                                   "(Cython path as shipped); 524 288 pairs/step",
                       "B": L_LAYERS * B, "Q": Q, "G": G, "parallelism": "replicas x%d (path does not shard, SURVEY 8e)" % world,
                       "l2": "48 rotating input/output sets = 250 MB > 126 MB L2"},
            "ms_per_step_python_loop": g["ms_per_step_python_loop"],
            "roofline": g["roofline"], "cpu_baseline": cpu, "e2e": g["e2e"], "gpu_launches": g["launches"], "clocks": clocks,
            "variants": g["variants"], "ap_eval": apres, "extra": extras,
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
