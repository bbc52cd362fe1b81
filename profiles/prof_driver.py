"""Launch every hot kernel exactly once at its BASELINE-config size, for `ncu` (no warm-up: ncu replays each launch).

    ncu --set full --clock-control none --import-source on -k regex:"^(ap|apx|clip|giou|lsap|nms|points|pseudo)" -o gpurun_out/prof_all python profiles/prof_driver.py
    ncu -i gpurun_out/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv
    python profiles/summarize_ncu.py gpurun_out/prof_all_raw.csv profiles/ncu_summary.json gpurun_out/prof_labels.json

The labels of the launches, in order, go to gpurun_out/prof_labels.json; summarize_ncu.py keys the summary by them
(bench.py looks its `roofline.traffic` / issue figures up by label).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import ovdet_b200  # noqa: F401
from ovdet_b200 import synth
from ovdet_b200.criterion import Matcher
from ovdet_b200.models.model_3detr import clip_logits
from ovdet_b200.utils import ap_calculator as APC, eval_det as ED
from ovdet_b200.utils.box_3d_utils import lift_filter_batch
from ovdet_b200.utils.box_util import generalized_box3d_iou
from ovdet_b200.utils.nms import nms_batch
from ovdet_b200.utils.points_in_box import points_in_boxes_count

dev = torch.device("cuda")
only = set(sys.argv[1:])
want = lambda k: not only or k in only
labels = []

if want("giou"):
    out, tgt = bench.giou_inputs(100)   # config 1 x 8 decoder layers
    c1, c2, nk = out["box_corners"].to(dev), tgt["gt_box_corners"].to(dev), tgt["nactual_gt"].to(dev)
    generalized_box3d_iou(c1, c2, nk); labels.append("giou3d_default")                                              # reference default (Cython semantics)
    generalized_box3d_iou(c1, c2, nk, mode="tensor", k2_cap=0); labels.append("giou3d_tensor_nocap")                # torch path, no cap
    generalized_box3d_iou(c1, c2, nk, mode="tensor", k2_cap=0, prefilter=False); labels.append("giou3d_exact_noprefilter")   # every valid pair clipped
    o1, t1 = bench.giou_inputs(100, nb=bench.B)
    generalized_box3d_iou(o1["box_corners"].to(dev), t1["gt_box_corners"].to(dev), t1["nactual_gt"].to(dev)); labels.append("giou3d_config1_b8")
if want("matcher"):
    out, tgt = synth.detection_batch(B=64, Q=256, G=64, C=18, seed=3, room="scannet", heading=0.0)   # config 2 x 8 layers
    o = {k: v.to(dev) for k, v in out.items()}
    t = {k: v.to(dev) for k, v in tgt.items()}
    Matcher(1, 0, 2, 0).match_from_boxes(o, t, rotated_boxes=False, return_assignments=False)
    labels += ["matcher_cost", "lsap"]
if want("ap"):
    S = 5050                                                                        # config 3
    out, tgt = bench.ap_inputs(S)
    dv = {k: v.to(dev).contiguous() for k, v in {**out, **tgt}.items()}
    calc = APC.APCalculator(bench._Cfg(), ap_iou_thresh=[0.25, 0.5], exact_eval=False)
    calc.step(dv["box_corners"], dv["sem_cls_prob"], dv["objectness_prob"], None, dv["gt_box_corners"], dv["gt_box_sem_cls_label"], dv["gt_box_present"])
    calc.compute_metrics()
    labels += ["ap_front2", "apx_merge", "apx_hist", "apx_final"]
if want("logits"):
    x, tx = synth.clip_logits_inputs(8192, 640, 1203)                               # config 4
    clip_logits((x * 0.25).to(dev), tx.to(dev)); labels.append("clip_logits")
if want("pseudo"):
    bx, pool = synth.pseudo_label_scenes(4096, P=256, pool=512, seed=5)             # config 5 scene shape
    lift_filter_batch(bx.to(dev), pool.to(dev)); labels.append("pseudo_filter")
if want("points"):
    out, tgt = synth.detection_batch(B=8, Q=128, G=64, C=20, seed=4, heading=np.pi)
    pc = synth.scene_points(out["box_corners"], n_points=20000, seed=1)
    points_in_boxes_count(pc.to(dev), out["box_corners"].to(dev)); labels.append("points_in_boxes")
if want("nms"):
    gN = torch.Generator().manual_seed(11)
    cN, sN, _ = synth.sample_boxes(gN, (4096, 256), "scannet", 0.0)
    bN = torch.cat([(cN - sN / 2).double(), (cN + sN / 2).double(), torch.rand((4096, 256, 1), generator=gN).double(),
                    torch.randint(0, 18, (4096, 256, 1), generator=gN).double()], -1).to(dev)
    nms_batch(bN, 0.25, samecls=True, want_order=False); labels.append("nms_samecls")
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(labels, open(os.path.join(ROOT, "gpurun_out", "prof_labels.json"), "w"))
print("prof_driver done", labels)
