"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run once in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

It imports the reference's modules from /root/reference unchanged (with
``sys.modules`` stubs for the absent third-party imports), uses the
reference's own Cython extension compiled into oracle/_ref by oracle/build.py,
feeds them the seeded synthetic inputs of the package's ``synth`` module and
stores inputs + reference outputs as small .npz files.  Nothing here is needed
at test time: the tests read only the committed .npz files (the GPU box has no
/root/reference).  The reference ships no tests or fixtures of its own
(SURVEY.md 4); the known-answer polygons come from the demo inputs at
utils/box_ops3d.py:740-766 evaluated with the importable twins in utils/box_util.py.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("OVDET_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

import ovdet_b200  # noqa: E402  (alias module at the repo root)
from ovdet_b200 import synth  # noqa: E402
from oracle import build as obuild  # noqa: E402


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    mod("detectron2")
    mod("detectron2.structures", Boxes=_Dummy, Instances=_Dummy)
    mod("detectron2.modeling")
    mod("detectron2.modeling.meta_arch", CLIPFastRCNN=_Dummy)
    mod("detectron2.config", get_cfg=lambda: None)
    mod("detectron2.checkpoint", DetectionCheckpointer=_Dummy)
    mod("third_party")
    mod("third_party.pointnet2")
    mod("third_party.pointnet2.pointnet2_modules", PointnetSAModuleVotes=_Dummy)
    mod("third_party.pointnet2.pointnet2_utils", furthest_point_sample=lambda *a, **k: None)
    mod("plyfile", PlyData=_Dummy, PlyElement=_Dummy)
    mod("trimesh")
    mod("imageio", imread=lambda *a, **k: None)
    mod("torch_ema", ExponentialMovingAverage=_Dummy)


def import_reference():
    install_stubs()
    sys.path.insert(0, REF)
    # make the compiled Cython extension visible as utils.box_intersection
    so = obuild.build_ref()
    import utils  # the reference's package (namespace)
    spec = importlib.util.spec_from_file_location("utils.box_intersection", so)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    sys.modules["utils.box_intersection"] = m
    utils.box_intersection = m
    import utils.box_util as box_util
    import utils.nms as nms
    import utils.eval_det as eval_det
    import utils.ap_calculator as ap_calculator
    import utils.label_formatter as label_formatter
    import criterion
    spec = importlib.util.spec_from_file_location("ref_box_3d_utils",
                                                  os.path.join(REF, "3DOVDet_tools/utils/box_3d_utils.py"))
    tools = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tools)
    assert box_util.box_intersection is not None
    return box_util, nms, eval_det, ap_calculator, label_formatter, criterion, tools


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrs.items()})


def golden_giou_grad(box_util):
    """Autograd of the reference's TorchScript-path GIoU (utils/box_util.py:517-618, the needs_grad=True branch at
    :725-730) w.r.t. the predicted box PARAMETERS and corners, with the sparse upstream gradient loss_giou produces
    (criterion.py:274-296): one matched query per ground-truth box."""
    out = {}
    for tag, heading, rotated in (("rot", 0.6, True), ("axis", 0.0, False)):
        g = torch.Generator().manual_seed(41 if rotated else 42)
        B, K1, K2 = 2, 24, 6
        ctr2, size2, ang2 = synth.sample_boxes(g, (B, K2), "sunrgbd", heading)
        c2 = box_util.get_3d_box_batch_tensor(size2, ang2, ctr2)
        # predictions = jittered copies of the GT boxes (so matched pairs overlap) + random boxes
        rep = (K1 + K2 - 1) // K2
        ctr1 = (ctr2.repeat(1, rep, 1)[:, :K1] + 0.15 * torch.randn(B, K1, 3, generator=g)).clone()
        size1 = (size2.repeat(1, rep, 1)[:, :K1] * (1 + 0.2 * torch.rand(B, K1, 3, generator=g))).clone()
        ang1 = (ang2.repeat(1, rep)[:, :K1] + (0.3 * torch.randn(B, K1, generator=g) if rotated else 0)).clone()
        nk = torch.tensor([K2, K2 - 2], dtype=torch.int64)
        w = torch.zeros(B, K1, K2)
        for b in range(B):
            for j in range(K2):          # query j + r*K2 is a jittered copy of GT j
                w[b, j, j] = float(torch.rand((), generator=g)) + 0.5
                w[b, j + K2, j] = float(torch.rand((), generator=g)) + 0.5
                w[b, j + 2 * K2, (j + 1) % K2] = -float(torch.rand((), generator=g)) - 0.5   # a non-matching pair too
                w[b, j + 3 * K2, j] = 1.0
        ctr1.requires_grad_(True); size1.requires_grad_(True); ang1.requires_grad_(True)
        corners1 = box_util.get_3d_box_batch_tensor(size1, ang1, ctr1)
        corners1.retain_grad()
        giou = box_util.generalized_box3d_iou(corners1, c2, nk, rotated_boxes=rotated, needs_grad=True)
        (giou * w).sum().backward()
        out.update({f"{tag}_center1": ctr1.detach().numpy(), f"{tag}_size1": size1.detach().numpy(),
                    f"{tag}_angle1": ang1.detach().numpy(), f"{tag}_corners1": corners1.detach().numpy(),
                    f"{tag}_corners2": c2.numpy(), f"{tag}_nums_k2": nk.numpy(), f"{tag}_w": w.numpy(),
                    f"{tag}_giou": giou.detach().numpy(), f"{tag}_grad_corners1": corners1.grad.numpy(),
                    f"{tag}_grad_center1": ctr1.grad.numpy(), f"{tag}_grad_size1": size1.grad.numpy(),
                    f"{tag}_grad_angle1": ang1.grad.numpy()})
    save("giou_grad.npz", **out)


def golden_holes():
    """Fixtures for the two API corners added in round 2, from the reference's own tools modules: the ``lhs`` re-pick of
    3DOVDet_tools/utils/box_3d_utils.py:113-116 and the axis-aligned evaluation of
    3DOVDet_tools/utils/evaluation/eval_det.py:86 (get_iou_func=get_iou -> evaluation/box_util.py:287-309)."""
    sys.path.insert(0, os.path.join(REF, "3DOVDet_tools/utils"))
    import evaluation.eval_det as ted
    spec = importlib.util.spec_from_file_location("ref_box_3d_utils2", os.path.join(REF, "3DOVDet_tools/utils/box_3d_utils.py"))
    tools = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tools)
    rng = np.random.default_rng(7)
    out = {}
    for tag, cw in (("plain", False), ("cls", True)):
        K = 96
        c = rng.uniform(0, 4, (K, 3)); sz = rng.uniform(0.5, 2.0, (K, 3))
        b = np.concatenate([c - sz / 2, c + sz / 2, rng.permutation(K)[:, None] / K + 0.001, rng.integers(0, 4, (K, 1)).astype(float)], 1)
        picked = tools.nms_3d_faster(b.copy(), 0.1, class_wise=cw, lhs=True)
        out["lhs_boxes_" + tag] = b
        out["lhs_pick_" + tag] = np.array([int(np.where((b == row).all(1))[0][0]) for row in picked])
    S = 12
    pred, gt = {}, {}
    for s in range(S):
        ng = int(rng.integers(0, 6))
        g = [np.concatenate([rng.uniform(0, 5, 3), rng.uniform(0.5, 2, 3)]) for _ in range(ng)]
        gt[s] = g
        p = []
        for gb in g:
            for _ in range(int(rng.integers(1, 4))):
                p.append((np.concatenate([gb[:3] + rng.normal(0, 0.25, 3), gb[3:] * rng.uniform(0.7, 1.3, 3)]), float(rng.uniform(0.05, 1))))
        for _ in range(int(rng.integers(0, 8))):
            p.append((np.concatenate([rng.uniform(0, 5, 3), rng.uniform(0.5, 2, 3)]), float(rng.uniform(0.05, 1))))
        pred[s] = p
    for thr in (0.25, 0.5):
        rec, prec, ap = ted.eval_det_cls(pred, gt, thr, False, ted.get_iou)
        out[f"aabb_rec_{thr}"] = rec; out[f"aabb_prec_{thr}"] = prec; out[f"aabb_ap_{thr}"] = np.array(ap)
    out["aabb_ap07"] = np.array(ted.eval_det_cls(pred, gt, 0.25, True, ted.get_iou)[2])
    K = max(len(v) for v in pred.values()); G = max(len(v) for v in gt.values())
    pb = np.zeros((S, K, 7)); pn = np.zeros(S, np.int64); gb = np.zeros((S, max(G, 1), 6)); gn = np.zeros(S, np.int64)
    for s in range(S):
        pn[s] = len(pred[s]); gn[s] = len(gt[s])
        for k, (b, sc) in enumerate(pred[s]):
            pb[s, k, :6] = b; pb[s, k, 6] = sc
        for j, b in enumerate(gt[s]):
            gb[s, j] = b
    out.update(aabb_pred=pb, aabb_npred=pn, aabb_gt=gb, aabb_ngt=gn)
    a = np.concatenate([rng.uniform(0, 3, 3), rng.uniform(0.5, 2, 3)])
    b2 = np.concatenate([a[:3] + 0.3, a[3:] * 1.1])
    out.update(calc_iou_a=a, calc_iou_b=b2, calc_iou=np.array(ted.get_iou(a, b2)))
    save("holes.npz", **out)


def golden_project(ref_root):
    """project_box_3d_cuda + SUNRGBD_Calibration_cuda (utils/image_util.py) and the image clip of criterion.py:387-391."""
    spec = importlib.util.spec_from_file_location("ref_image_util", os.path.join(ref_root, "utils", "image_util.py"))
    iu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(iu)
    g = torch.Generator().manual_seed(17)
    B, Q = 3, 40
    ctr = torch.stack([torch.rand(B, Q, generator=g) * 4 - 2, torch.rand(B, Q, generator=g) * 4 + 1.5, torch.rand(B, Q, generator=g) * 2 - 1], -1)
    size = torch.rand(B, Q, 3, generator=g) * 0.8 + 0.15          # half extents, as the reference passes them
    ang = torch.rand(B, Q, generator=g) * 6.28 - 3.14
    out, outc, Rs, Ks, whs = [], [], [], [], []
    for b in range(B):
        tilt = float(torch.rand((), generator=g)) * 0.3 - 0.15
        Rt = torch.tensor([[1, 0, 0], [0, np.cos(tilt), -np.sin(tilt)], [0, np.sin(tilt), np.cos(tilt)]], dtype=torch.float32)
        K = torch.tensor([[529.5 + 10 * b, 0, 365.0], [0, 529.5 + 5 * b, 265.0], [0, 0, 1]], dtype=torch.float32)
        calib = iu.SUNRGBD_Calibration_cuda(Rt, K)
        bx = iu.project_box_3d_cuda(calib, ctr[b], size[b], ang[b])
        w, h = 730, 530
        mx = torch.broadcast_to(torch.tensor([[w, h, w, h]]), bx.size())
        bc = torch.minimum(torch.clamp_min(bx, 0), mx)
        out.append(bx.numpy()); outc.append(bc.numpy()); Rs.append(Rt.numpy()); Ks.append(K.numpy()); whs.append([w, h])
    save("project.npz", center=ctr.numpy(), size=size.numpy(), angle=ang.numpy(), Rtilt=np.stack(Rs), K=np.stack(Ks),
         image_wh=np.array(whs, np.float32), boxes=np.stack(out), boxes_clipped=np.stack(outc))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "holes":     # only the lhs-NMS / axis-aligned-evaluation fixture
        golden_holes()
        return
    box_util, nms, eval_det, apc, lf, criterion, tools = import_reference()
    if len(sys.argv) <= 1:
        golden_holes()
    if len(sys.argv) > 1 and sys.argv[1] == "project":   # regenerate only the projection fixture
        golden_project(REF)
        return
    torch.manual_seed(0)
    np.random.seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == "grad":      # regenerate only the backward fixture
        golden_giou_grad(box_util)
        return
    golden_giou_grad(box_util)

    # ---------------- known-answer vectors ----------------
    sub_poly = [(0, 0), (300, 0), (300, 300), (0, 300)]
    clip_poly = [(150, 150), (300, 300), (150, 450), (0, 300)]
    ip = np.array(box_util.polygon_clip(sub_poly, clip_poly), np.float64)
    a0 = box_util.poly_area(ip[:, 0], ip[:, 1])
    rect1 = [(50, 0), (50, 300), (300, 300), (300, 0)]
    _, a1 = box_util.convex_hull_intersection(rect1, clip_poly)
    r1 = [(0.30026005199835404, 8.9408694211408424), (-1.1571105364358421, 9.4686676477075533),
          (0.1777082043006144, 13.154404877812102), (1.6350787927348105, 12.626606651245391)]
    r1 = [r1[0], r1[3], r1[2], r1[1]]
    r2 = [(0.23908745901608636, 8.8551095691132886), (-1.2771419487733995, 9.4269062966181956),
          (0.13138836963152717, 13.161896351296868), (1.647617777421013, 12.590099623791961)]
    r2 = [r2[0], r2[3], r2[2], r2[1]]
    ip2, a2 = box_util.convex_hull_intersection(r1, r2)

    def cube(cx, cy, cz, s=1.0):
        return box_util.get_3d_box((s, s, s), 0.0, (cx, cy, cz))

    unit = cube(0, 0, 0)
    # axis-aligned fp32 box against itself (a ROTATED box against itself makes the reference divide by ~0
    # and Qhull raise "Points cannot contain NaN" -- documented degenerate case)
    ident = synth.params_to_corners(torch.tensor([[0.3, -0.2, 0.1]]), torch.tensor([[1.1, 0.7, 0.9]]),
                                    torch.tensor([0.0]))[0].numpy()
    hand = {
        "identical_f32": np.array(box_util.box3d_iou(ident.astype(np.float64), ident.astype(np.float64))),
        "touching": np.array(box_util.box3d_iou(unit, cube(1.0, 0, 0))),
        "offset": np.array(box_util.box3d_iou(unit, cube(0.5, 0.25, 0.0))),
    }
    vrec = np.array([.25, .5, .5, .75, 1.0])
    vprec = np.array([1, 1, 2 / 3, .75, .8])
    nms_in = np.array([[0, 0, 0, 1, 1, 1, .9], [.1, 0, 0, 1.1, 1, 1, .8], [2, 2, 2, 3, 3, 3, .7], [0, 0, 0, 1, 1, .5, .6]])
    save("kat.npz",
         sub_poly=np.array(sub_poly, np.float64), clip_poly=np.array(clip_poly, np.float64), inter0=ip, area0=a0,
         rect1=np.array(rect1, np.float64), area1=a1, r1=np.array(r1), r2=np.array(r2),
         inter2=np.array(ip2, np.float64), area2=a2,
         unit=unit, ident=ident, touch=cube(1.0, 0, 0), off=cube(0.5, 0.25, 0.0),
         iou_identical_f32=hand["identical_f32"], iou_touching=hand["touching"], iou_offset=hand["offset"],
         voc_rec=vrec, voc_prec=vprec, voc_ap=eval_det.voc_ap(vrec, vprec), voc_ap07=eval_det.voc_ap(vrec, vprec, True),
         nms_in=nms_in, nms_pick=np.array(nms.nms_3d_faster(nms_in, 0.25)))

    # ---------------- GIoU, all reference variants ----------------
    for tag, heading, seed in (("pi", np.pi, 1), ("half", 0.5, 2)):
        out, tgt = synth.detection_batch(B=2, Q=32, G=16, C=20, seed=seed, heading=heading, max_gt=16)
        c1, c2, nk = out["box_corners"], tgt["gt_box_corners"], tgt["nactual_gt"]
        res = {}
        res["cython_shipped"] = box_util.generalized_box3d_iou_cython(c1, c2, nk, True, False).numpy()
        res["cython_shipped_inter"] = box_util.generalized_box3d_iou_cython(c1, c2, nk, True, True).numpy()
        res["tensor"] = box_util.generalized_box3d_iou_tensor(c1, c2, nk, True, False).numpy()
        res["tensor_inter"] = box_util.generalized_box3d_iou_tensor(c1, c2, nk, True, True).numpy()
        res["nonrot"] = box_util.generalized_box3d_iou_cython(c1, c2, nk, False, False).numpy()
        res["dispatch_default"] = box_util.generalized_box3d_iou(c1, c2, nk).numpy()
        # exact pairwise IoU (box3d_iou) for the valid columns
        ex = np.zeros((2, 32, 16))
        for b in range(2):
            for i in range(32):
                for j in range(int(nk[b])):
                    ex[b, i, j] = box_util.box3d_iou(c1[b, i].numpy().astype(np.float64), c2[b, j].numpy().astype(np.float64))[0]
        res["exact_iou"] = ex
        save(f"giou_{tag}.npz", corners1=c1.numpy(), corners2=c2.numpy(), nums_k2=nk.numpy(), **res)

    # ---------------- box_intersection (Cython ABI) directly ----------------
    out, tgt = synth.detection_batch(B=2, Q=16, G=8, C=20, seed=3, heading=0.5, max_gt=8)
    c1, c2, nk = out["box_corners"], tgt["gt_box_corners"], tgt["nactual_gt"]
    rect1 = c1[:, :, [3, 2, 1, 0]][..., [0, 2]].contiguous().numpy()
    rect2 = c2[:, :, [3, 2, 1, 0]][..., [0, 2]].contiguous().numpy()
    nonrot = np.ones((2, 16, 8), np.float32)
    nonrot[:, ::3, :] = 0
    ia_a = np.zeros((2, 16, 8), np.float32)
    ia_e = np.zeros((2, 16, 8), np.float32)
    import utils.box_intersection as bi
    bi.box_intersection(rect1, rect2, nonrot, nk.numpy().astype(np.int32), ia_a, True)
    bi.box_intersection(rect1, rect2, nonrot, nk.numpy().astype(np.int32), ia_e, False)
    save("box_intersection.npz", rect1=rect1, rect2=rect2, nonrot=nonrot, nums_k2=nk.numpy().astype(np.int32),
         approx=ia_a, exact=ia_e)

    # ---------------- Matcher ----------------
    for tag, room, heading, Q, w in (("sunrgbd", "sunrgbd", np.pi, 48, (1, 5, 3, 5)), ("scannet", "scannet", 0.0, 64, (1, 0, 2, 0))):
        out, tgt = synth.detection_batch(B=3, Q=Q, G=16, C=18, seed=5, room=room, heading=heading, max_gt=16)
        tgt["nactual_gt"][1] = 0  # empty-GT sample (criterion.py:78,86)
        tgt["gt_box_present"][1] = 0
        gious = box_util.generalized_box3d_iou_tensor(out["box_corners"], tgt["gt_box_corners"], tgt["nactual_gt"],
                                                      heading > 0, False)
        cd = torch.cdist(out["center_normalized"], tgt["gt_box_centers_normalized"], p=1)
        o = dict(out)
        o["gious"] = gious
        o["center_dist"] = cd
        # Matcher(cost_class, cost_objectness, cost_giou, cost_center)
        m = criterion.Matcher(w[0], w[1], w[2], w[3])
        r = m(o, tgt)
        cost = (w[0] * -torch.gather(o["sem_cls_prob"], 2, tgt["gt_box_sem_cls_label"].unsqueeze(1).expand(3, Q, 16))
                + w[1] * -o["objectness_prob"].unsqueeze(-1) + w[3] * cd + w[2] * -gious)
        rows = np.full((3, 16), -1, np.int64)
        cols = np.full((3, 16), -1, np.int64)
        for b, a in enumerate(r["assignments"]):
            if len(a):
                rows[b, :len(a[0])] = a[0].numpy()
                cols[b, :len(a[1])] = a[1].numpy()
        save(f"matcher_{tag}.npz", corners1=out["box_corners"].numpy(), corners2=tgt["gt_box_corners"].numpy(),
             nactual=tgt["nactual_gt"].numpy(), sem_cls_prob=out["sem_cls_prob"].numpy(),
             objectness=out["objectness_prob"].numpy(), center_q=out["center_normalized"].numpy(),
             center_g=tgt["gt_box_centers_normalized"].numpy(), labels=tgt["gt_box_sem_cls_label"].numpy(),
             weights=np.array(w, np.float32), rotated=np.array(heading > 0),
             gious=gious.numpy(), center_dist=cd.numpy(), cost=cost.numpy(),
             per_prop_gt_inds=r["per_prop_gt_inds"].numpy(), proposal_matched_mask=r["proposal_matched_mask"].numpy(),
             assign_rows=rows, assign_cols=cols)

    # ---------------- NMS ----------------
    g = torch.Generator().manual_seed(7)
    c, s, _ = synth.sample_boxes(g, (3, 96), "sunrgbd", 0.0)
    s = s * 1.6
    score = torch.rand((3, 96), generator=g).double()
    cls = torch.randint(0, 4, (3, 96), generator=g).double()
    b8 = torch.cat([(c - s / 2).double(), (c + s / 2).double(), score[..., None], cls[..., None]], -1).numpy()
    picks = {}
    for i in range(3):
        picks[f"p3_{i}"] = np.array(nms.nms_3d_faster(b8[i, :, :7], 0.25))
        picks[f"p3old_{i}"] = np.array(nms.nms_3d_faster(b8[i, :, :7], 0.25, True))
        picks[f"p3c_{i}"] = np.array(nms.nms_3d_faster_samecls(b8[i], 0.25))
        picks[f"p3cold_{i}"] = np.array(nms.nms_3d_faster_samecls(b8[i], 0.1, True))
        b5 = b8[i][:, [0, 2, 3, 5, 6]]
        picks[f"p2_{i}"] = np.array(nms.nms_2d_faster(b5, 0.25))
        picks[f"p2old_{i}"] = np.array(nms.nms_2d_faster(b5, 0.25, True))
        t = np.concatenate([b8[i], np.prod(b8[i][:, 3:6] - b8[i][:, :3], -1, keepdims=True)], 1)
        picks[f"tools_cw_{i}"] = tools.nms_3d_faster(t.copy(), 0.7, class_wise=True)
        picks[f"tools_size_{i}"] = tools.nms_3d_faster(t.copy(), 0.0, use_size_score=True, class_wise=True, size_typ="Volume")
    save("nms.npz", boxes=b8, **picks)

    # ---------------- AP pipeline (parse_predictions + APCalculator) ----------------
    class _Cfg:
        num_semcls = 10

    S, Q, G, C = 24, 64, 16, 10
    out, tgt = synth.detection_batch(B=S, Q=Q, G=G, C=C, seed=11, heading=np.pi, max_gt=8)
    cfg = apc.get_ap_config_dict(dataset_config=_Cfg(), remove_empty_box=False)
    calc = apc.APCalculator(_Cfg(), ap_iou_thresh=[0.25, 0.5], exact_eval=False, ap_config_dict=cfg)
    dummy_pc = torch.zeros((S, 8, 3))
    preds = apc.parse_predictions(out["box_corners"], out["sem_cls_prob"], out["objectness_prob"], dummy_pc, cfg)
    calc.step(out["box_corners"], out["sem_cls_prob"], out["objectness_prob"], dummy_pc,
              tgt["gt_box_corners"], tgt["gt_box_sem_cls_label"], tgt["gt_box_present"])
    metrics = calc.compute_metrics()
    ap_arrs = {}
    for thr, d in metrics.items():
        for k, v in d.items():
            ap_arrs[f"m{thr}|{k}"] = np.float64(v)
    # per-class PR curves straight from eval_det_cls
    pred_c, gt_c = {}, {}
    for i in range(S):
        for cl, bb, sc in calc.pred_map_cls[i]:
            pred_c.setdefault(cl, {}).setdefault(i, []).append((bb, sc))
            gt_c.setdefault(cl, {}).setdefault(i, [])
        for cl, bb in calc.gt_map_cls[i]:
            gt_c.setdefault(cl, {}).setdefault(i, []).append(bb)
    for cl in (0, 3, 7):
        for thr in (0.25, 0.5):
            rec, prec, ap = eval_det.eval_det_cls(pred_c[cl], gt_c[cl], thr)
            ap_arrs[f"rec_c{cl}_t{thr}"] = rec
            ap_arrs[f"prec_c{cl}_t{thr}"] = prec
            ap_arrs[f"ap_c{cl}_t{thr}"] = np.float64(ap)
    n_pred = np.array([len(p) for p in preds])
    kept = np.zeros((S, Q), np.uint8)
    for i in range(S):
        for cl, bb, sc in preds[i]:
            if cl == 0:
                j = np.where((out["box_corners"][i].numpy() == bb).all((1, 2)))[0][0]
                kept[i, j] = 1
    # variants of parse_predictions
    var = {}
    for name, kw in (("nms3d_nocls", dict(cls_nms=False)), ("nms2d", dict(use_3d_nms=False)),
                     ("no_pcp", dict(per_class_proposal=False)),
                     ("clsconf", dict(per_class_proposal=False, use_cls_confidence_only=True)),
                     ("nonms", dict(no_nms=True)), ("old", dict(use_old_type_nms=True))):
        cfg2 = apc.get_ap_config_dict(dataset_config=_Cfg(), remove_empty_box=False, **kw)
        p2 = apc.parse_predictions(out["box_corners"], out["sem_cls_prob"], out["objectness_prob"], dummy_pc, cfg2)
        var[f"v_{name}_n"] = np.array([len(p) for p in p2])
        var[f"v_{name}_cls"] = np.concatenate([np.array([t[0] for t in p], np.int64) for p in p2])
        var[f"v_{name}_score"] = np.concatenate([np.array([t[2] for t in p], np.float32) for p in p2])
    save("ap.npz", box_corners=out["box_corners"].numpy(), sem_cls_prob=out["sem_cls_prob"].numpy(),
         objectness=out["objectness_prob"].numpy(), gt_corners=tgt["gt_box_corners"].numpy(),
         gt_labels=tgt["gt_box_sem_cls_label"].numpy(), gt_present=tgt["gt_box_present"].numpy(),
         n_pred=n_pred, kept=kept, **ap_arrs, **var)

    # ---------------- exact box3d_iou on random pairs ----------------
    g = torch.Generator().manual_seed(13)
    c, s, a = synth.sample_boxes(g, (400,), "sunrgbd", np.pi)
    c2 = c + torch.randn((400, 3), generator=g) * 0.4
    s2 = s * (torch.randn((400, 3), generator=g) * 0.2 + 1).clamp(0.5, 1.5)
    a2 = a + torch.randn((400,), generator=g) * 0.3
    A = synth.params_to_corners(c, s, a).numpy()
    Bc = synth.params_to_corners(c2, s2, a2).numpy()
    r = np.array([box_util.box3d_iou(A[i].astype(float), Bc[i].astype(float)) for i in range(400)])
    save("box3d_iou.npz", a=A, b=Bc, iou=r[:, 0], iou2d=r[:, 1])

    # ---------------- AABB 1xN IoU + lift pipeline ----------------
    bx, pool = synth.pseudo_label_scenes(4, P=96, pool=128, seed=17)
    bx, pool = bx.numpy(), pool.numpy()
    iou_vv = np.stack([lf.box_3d_iou(bx[0, i, :6], pool[0]) for i in range(16)])
    cs = bx[0, :16, :6].copy()
    cs[:, 3:6] -= cs[:, :3]
    cs[:, :3] += cs[:, 3:6] / 2
    pcs = pool[0].copy()
    pcs[:, 3:6] -= pcs[:, :3]
    pcs[:, :3] += pcs[:, 3:6] / 2
    iou_cs = np.stack([tools.box_3d_iou(cs[i], pcs, typ="cs") for i in range(16)])
    lift = {}
    for sidx in range(4):
        boxes = tools.nms_3d_faster(bx[sidx].copy(), 0.7, class_wise=True)
        box_pool = pool[sidx].copy()
        labels = -100 * np.ones(box_pool.shape[0])
        tmp_score = np.zeros(box_pool.shape[0])
        for box in boxes:  # 3DOVDet_tools/scannet/lift_boxes.py:151-158, verbatim semantics
            iou = tools.box_3d_iou(box, box_pool)
            if iou.max() < 0.3:
                continue
            index = np.argmax(iou)
            if box[-2] > tmp_score[index]:
                labels[index] = box[-1]
                tmp_score[index] = box[-2]
        scale = box_pool[:, 3:6] - box_pool[:, 0:3]
        box_pool = np.concatenate([box_pool[:, :6], np.stack(
            [tmp_score, labels, np.prod(scale, axis=-1), 2 * np.sum(scale * np.roll(scale, 1, axis=-1), axis=-1)], axis=1)], axis=-1)
        b2 = box_pool[labels != -100]
        if b2.shape[0]:
            b2 = tools.nms_3d_faster(b2, 0, use_size_score=True, class_wise=True, size_typ="Volume")
        lift[f"nms1_{sidx}"] = boxes
        lift[f"final_{sidx}"] = b2
    save("lift.npz", boxes=bx, pool=pool, iou_vv=iou_vv, iou_cs=iou_cs, **lift)

    # ---------------- remove_empty_box (Delaunay point-in-hull) ----------------
    class _Cfg2:
        num_semcls = 10

    out, tgt = synth.detection_batch(B=3, Q=48, G=8, C=10, seed=23, heading=np.pi, max_gt=8)
    pc = synth.scene_points(out["box_corners"], n_points=3000, seed=5)
    cfg = apc.get_ap_config_dict(dataset_config=_Cfg2(), remove_empty_box=True)
    preds = apc.parse_predictions(out["box_corners"], out["sem_cls_prob"], out["objectness_prob"], pc, cfg)
    counts = np.zeros((3, 48), np.int32)
    for i in range(3):
        for j in range(48):
            box3d = apc.flip_axis_to_depth(out["box_corners"][i, j].numpy())
            pin, _ = box_util.extract_pc_in_box3d(pc[i].numpy(), box3d)
            counts[i, j] = len(pin)
    kept = np.zeros((3, 48), np.uint8)
    for i in range(3):
        for cl, bb, sc in preds[i]:
            if cl == 0:
                kept[i, np.where((out["box_corners"][i].numpy() == bb).all((1, 2)))[0][0]] = 1
    save("empty.npz", box_corners=out["box_corners"].numpy(), sem_cls_prob=out["sem_cls_prob"].numpy(),
         objectness=out["objectness_prob"].numpy(), point_cloud=pc.numpy(), counts=counts, kept=kept,
         n_pred=np.array([len(p) for p in preds]))

    # ---------------- LabelFormatter.gen_pseudo (file based) ----------------
    import tempfile
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "labels")); os.makedirs(os.path.join(tmp, "out"))
    g = torch.Generator().manual_seed(29)
    scenes = ["sceneA", "sceneB"]
    raw, boxes_all = [], []
    for si, name in enumerate(scenes):
        c, s_, _ = synth.sample_boxes(g, (12,), "scannet", 0.0)
        lab = torch.randint(0, 18, (12,), generator=g).double()
        pts, pl = [], []
        for k in range(12):   # points inside box k mostly carry its label (or a wrong one for every third box)
            t = torch.rand((60, 3), generator=g) - 0.5
            pts.append(c[k] + t * s_[k])
            true = lab[k] if k % 3 else (lab[k] + 1) % 18
            l = torch.where(torch.rand(60, generator=g) < 0.7, true, torch.randint(0, 25, (60,), generator=g).double())
            pl.append(l)
        noise = torch.rand((500, 3), generator=g) * 8
        pts.append(noise); pl.append(torch.randint(0, 25, (500,), generator=g).double())
        arr = torch.cat([torch.cat(pts).double(), torch.cat(pl)[:, None]], 1).numpy()
        np.save(os.path.join(tmp, "labels", name + ".npy"), arr)
        raw.append(arr)
        rows = torch.cat([c.double(), s_.double(), lab[:, None], torch.rand((12, 2), generator=g).double(),
                          torch.full((12, 1), float(si)).double()], 1).numpy()
        boxes_all.append(rows)
    fmt = lf.LabelFormatter(tmp, os.path.join(tmp, "out"), os.path.join(tmp, "labels"), scenes)
    fmt.pseudo_boxes = np.concatenate(boxes_all, 0)
    res = {}
    for si, name in enumerate(scenes):
        nb = fmt.gen_pseudo(si)
        res[f"nbox_{si}"] = np.array(nb)
        res[f"bbox_{si}"] = np.load(os.path.join(tmp, "out", name + "_bbox.npy"))
        res[f"raw_{si}"] = raw[si]
    save("labelfmt.npz", pseudo_boxes=fmt.pseudo_boxes, **res)


if __name__ == "__main__":
    main()
