"""Launch every hot kernel exactly once at its BASELINE-config size, for `ncu` (no warm-up: ncu replays each launch).

    ncu --set full --clock-control none --import-source on -k regex:ovdet -o gpurun_out/prof_all python profiles/prof_driver.py
"""
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
from ovdet_b200.utils.points_in_box import points_in_boxes_count

dev = torch.device("cuda")
only = set(sys.argv[1:])
want = lambda k: not only or k in only

if want("giou"):
    out, tgt = bench.giou_inputs(100)   # config 1 x 8 decoder layers
    c1, c2, nk = out["box_corners"].to(dev), tgt["gt_box_corners"].to(dev), tgt["nactual_gt"].to(dev)
    generalized_box3d_iou(c1, c2, nk)                                              # reference default (Cython semantics)
    generalized_box3d_iou(c1, c2, nk, mode="tensor", k2_cap=0)                     # torch path, no cap
    generalized_box3d_iou(c1, c2, nk, mode="tensor", k2_cap=0, prefilter=False)    # every pair clipped
if want("matcher"):
    out, tgt = synth.detection_batch(B=64, Q=256, G=64, C=18, seed=3, room="scannet", heading=0.0)   # config 2 x 8 layers
    o = {k: v.to(dev) for k, v in out.items()}
    t = {k: v.to(dev) for k, v in tgt.items()}
    Matcher(1, 0, 2, 0).match_from_boxes(o, t, rotated_boxes=False, return_assignments=False)
if want("ap"):
    S = 5050                                                                        # config 3
    out, tgt = bench.ap_inputs(S)
    dv = {k: v.to(dev).contiguous() for k, v in {**out, **tgt}.items()}
    cfg = APC.get_ap_config_dict(dataset_config=bench._Cfg(), remove_empty_box=False)
    _, keep, cls, clsp = APC.parse_predictions_device(dv["box_corners"], dv["sem_cls_prob"], dv["objectness_prob"], cfg)
    rs, rt, npos = ED.ap_match(dv["box_corners"], dv["sem_cls_prob"], dv["objectness_prob"], keep, dv["gt_box_corners"],
                               dv["gt_box_sem_cls_label"], dv["gt_box_present"], 20, [0.25, 0.5])
    ED.ap_reduce_compact(rs, rt, npos, 2, cap=2048)
    if "sort" in only:
        ED.ap_reduce(rs, rt, npos, 2)
if want("logits"):
    x, tx = synth.clip_logits_inputs(8192, 640, 1203)                               # config 4
    clip_logits((x * 0.25).to(dev), tx.to(dev))
if want("pseudo"):
    bx, pool = synth.pseudo_label_scenes(4096, P=256, pool=512, seed=5)             # config 5 scene shape
    lift_filter_batch(bx.to(dev), pool.to(dev))
if want("points"):
    out, tgt = synth.detection_batch(B=8, Q=128, G=64, C=20, seed=4, heading=np.pi)
    pc = synth.scene_points(out["box_corners"], n_points=20000, seed=1)
    points_in_boxes_count(pc.to(dev), out["box_corners"].to(dev))
torch.cuda.synchronize()
print("prof_driver done")
