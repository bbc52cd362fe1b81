"""Randomised shapes and mode switches against the oracle (GPU): GIoU (every mode / cap / prefilter combination, ragged
tiles, empty GT), the NMS family, the AP evaluation end to end and the matcher + LSAP.  Complements the fixed cases of
test_gpu_parity.py; each seed draws ~230 cases and runs in a few seconds."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ovdet_b200  # noqa: E402,F401
from ovdet_b200 import synth  # noqa: E402
from ovdet_b200.utils import box_util as BU, nms as NMS, ap_calculator as APC  # noqa: E402
from ovdet_b200.criterion import Matcher  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Cfg:
    def __init__(s, n):
        s.num_semcls = n


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_against_oracle(seed):
    rng = np.random.default_rng(seed)
    n_g = n_n = n_a = n_m = 0
    # ---- GIoU
    for it in range(120):
        B = int(rng.integers(1, 5)); Q = int(rng.integers(1, 180)); G = int(rng.integers(1, 140))
        heading = float(rng.choice([0.0, 0.3, 1.0, np.pi]))
        out, tgt = synth.detection_batch(B=B, Q=Q, G=G, seed=int(rng.integers(1 << 30)), heading=heading, max_gt=G,
                                         room=str(rng.choice(["sunrgbd", "scannet"])))
        c1, c2, nk = out["box_corners"], tgt["gt_box_corners"], tgt["nactual_gt"].clone()
        if rng.random() < 0.3: nk[int(rng.integers(B))] = 0
        rot = bool(rng.random() < 0.8)
        mode = str(rng.choice(["tensor", "cython"])); cap = int(rng.choice([0, 1, 4, 7])); pre = bool(rng.random() < 0.7); inter = bool(rng.random() < 0.2)
        use_nk = rng.random() < 0.85
        want = oracle.generalized_box3d_iou(c1, c2, nk if use_nk else None, rot, inter, mode=mode, prefilter=pre, k2_cap=cap or None)
        got = BU.generalized_box3d_iou(c1.to(DEV), c2.to(DEV), nk.to(DEV) if use_nk else None, rot, inter, mode=mode, prefilter=pre, k2_cap=cap).cpu().numpy()
        if not np.allclose(got, want, rtol=1e-5, atol=1e-6, equal_nan=True):
            pytest.fail(str(("GIOU MISMATCH", dict(B=B, Q=Q, G=G, heading=heading, rot=rot, mode=mode, cap=cap, pre=pre, inter=inter, use_nk=use_nk), np.abs(got - want).max())))
        n_g += 1
    # ---- NMS
    for it in range(60):
        K = int(rng.integers(1, 700)); ncls = int(rng.integers(1, 40)); S = 2
        g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
        c, s, _ = synth.sample_boxes(g, (S, K), "scannet", 0.0)
        s = s * float(rng.uniform(0.8, 2.5))
        score = torch.rand((S, K), generator=g).double()
        cls = torch.randint(0, ncls, (S, K), generator=g).double()
        bx = torch.cat([(c - s / 2).double(), (c + s / 2).double(), score[..., None], cls[..., None]], -1)
        thr = float(rng.choice([0.0, 0.1, 0.25, 0.7])); old = bool(rng.random() < 0.3); same = bool(rng.random() < 0.7)
        keep, order, npick = NMS.nms_batch(bx.to(DEV), thr, old_type=old, samecls=same, want_order=bool(rng.random() < 0.5) or True)
        for i in range(S):
            want = oracle.nms_3d_faster_samecls(bx[i].numpy(), thr, old) if same else oracle.nms_3d_faster(bx[i, :, :7].numpy(), thr, old)
            if order[i, :int(npick[i])].cpu().tolist() != want:
                pytest.fail(str(("NMS MISMATCH", dict(K=K, ncls=ncls, thr=thr, old=old, same=same))))
        n_n += 1
    # ---- AP end to end
    for it in range(25):
        S = int(rng.integers(1, 12)); Q = int(rng.choice([16, 64, 128, 200])); G = int(rng.choice([8, 64])); C = int(rng.choice([3, 10, 20]))
        out, tgt = synth.detection_batch(B=S, Q=Q, G=G, C=C, seed=int(rng.integers(1 << 30)), heading=float(rng.choice([0.0, np.pi])), max_gt=int(rng.integers(1, G + 1)))
        thrs = tuple(sorted(rng.choice([0.1, 0.25, 0.5, 0.75], size=int(rng.integers(1, 4)), replace=False).tolist()))
        calc = APC.APCalculator(Cfg(C), ap_iou_thresh=list(thrs), exact_eval=False)
        calc.step(out["box_corners"].to(DEV), out["sem_cls_prob"].to(DEV), out["objectness_prob"].to(DEV), None,
                  tgt["gt_box_corners"].to(DEV), tgt["gt_box_sem_cls_label"].to(DEV), tgt["gt_box_present"].to(DEV))
        got = calc.compute_metrics()
        want, _ = oracle.ap_metrics(out["box_corners"], out["sem_cls_prob"], out["objectness_prob"], tgt["gt_box_corners"],
                                    tgt["gt_box_sem_cls_label"], tgt["gt_box_present"], C, ap_iou_thresh=thrs)
        for thr in thrs:
            for k, v in want[thr].items():
                a, b_ = float(got[thr][k]), float(v)
                if not (abs(a - b_) < 1e-9 or (np.isnan(a) and np.isnan(b_))):
                    pytest.fail(str(("AP MISMATCH", dict(S=S, Q=Q, G=G, C=C, thrs=thrs, key=k, got=a, want=b_))))
        n_a += 1
    # ---- matcher
    for it in range(25):
        B = int(rng.integers(1, 6)); Q = int(rng.choice([32, 128, 256, 300])); G = int(rng.choice([8, 64, 100])); C = 18
        out, tgt = synth.detection_batch(B=B, Q=Q, G=G, C=C, seed=int(rng.integers(1 << 30)), heading=float(rng.choice([0.0, np.pi])), max_gt=G)
        w = [float(x) for x in rng.uniform(0, 5, size=4)]
        m = Matcher(*w)
        o = {k: v.to(DEV) for k, v in out.items()}; t = {k: v.to(DEV) for k, v in tgt.items()}
        rot = bool(rng.random() < 0.5)
        res = m.match_from_boxes(o, t, rotated_boxes=rot)
        cost = res["final_cost"].cpu().numpy()
        _, inds_w, mask_w = oracle.matcher_assign(cost, tgt["nactual_gt"].numpy())
        if not np.array_equal(res["per_prop_gt_inds"].cpu().numpy(), inds_w) or not np.array_equal(res["proposal_matched_mask"].cpu().numpy(), mask_w):
            pytest.fail(str(("MATCHER MISMATCH", dict(B=B, Q=Q, G=G, w=w, rot=rot))))
        n_m += 1
    assert (n_g, n_n, n_a, n_m) == (120, 60, 25, 25)


@pytest.mark.parametrize("seed", [21])
def test_fuzz_pseudo_label_filter(seed):
    """lift_boxes' per-scene NMS -> pool match -> size NMS on random proposal / pool sizes and thresholds vs the oracle."""
    from ovdet_b200.utils import box_3d_utils as B3
    rng = np.random.default_rng(seed)
    for it in range(40):
        P = int(rng.integers(1, 257)); M = int(rng.integers(1, 513)); C = int(rng.integers(1, 19))
        bx, pool = synth.pseudo_label_scenes(2, P=P, pool=M, C=C, seed=int(rng.integers(1 << 30)))
        nms_t = float(rng.choice([0.25, 0.5, 0.7])); match_t = float(rng.choice([0.0, 0.1, 0.3, 0.6])); size_t = float(rng.choice([0.0, 0.1]))
        for s in range(2):
            want = oracle.lift_filter_scene(bx[s].numpy(), pool[s].numpy(), nms_t, match_t, size_t)
            got = B3.lift_filter_scene(bx[s].numpy(), pool[s].numpy(), nms_t, match_t, size_t)
            assert got.shape == want.shape, (P, M, C, nms_t, match_t, size_t)
            np.testing.assert_allclose(got, want, rtol=0, atol=0)


@pytest.mark.parametrize("seed", [31])
def test_fuzz_clip_logits_shapes(seed):
    """Random (M, K, N) for the logits kernels: ragged M, N from a single tile to eight, every cluster size the
    co-residency model can pick, both kernels (persistent for >= 2 rounds of M-tiles), with and without normalisation."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_gpu_parity as T
    rng = np.random.default_rng(seed)
    for it in range(12):
        M = int(rng.choice([1, 77, 500, 1500, 3000, 6000])) + int(rng.integers(0, 50))
        K = int(rng.choice([64, 128, 640]))
        N = int(rng.choice([2, 19, 21, 100, 257, 600, 1203, 2048])) + (int(rng.integers(0, 7)) if it % 2 else 0)
        N = min(max(N, 2), 2048)
        l2 = bool(rng.random() < 0.3)
        T._check_clip_logits(M, K, N, l2, (1.0 / 0.07) if l2 else float(rng.choice([0.5, 1.0])))


@pytest.mark.parametrize("seed", [41])
def test_fuzz_ap_configs_over_virtual_ranks(seed):
    """Random parse_predictions configurations (NMS variant, record layout, thresholds) through the fused front end and
    the exchange reducer for 1..8 virtual ranks (stage by stage on one GPU), against the oracle's parse_predictions +
    eval_det of all scenes; every rank must report the oracle's numbers."""
    from test_gpu_parity import _apx_virtual_ranks
    rng = np.random.default_rng(seed)
    for it in range(14):
        S = int(rng.integers(1, 30)); Q = int(rng.choice([16, 64, 128, 160])); G = int(rng.choice([4, 24, 64])); C = int(rng.choice([2, 7, 20]))
        nranks = int(rng.choice([1, 2, 3, 5, 8]))
        thrs = tuple(sorted(rng.choice([0.1, 0.25, 0.5, 0.75], size=int(rng.integers(1, 4)), replace=False).tolist()))
        kw = dict(nms_iou=float(rng.choice([0.1, 0.25, 0.5])), conf_thresh=float(rng.choice([0.0, 0.05, 0.3])),
                  cls_nms=bool(rng.random() < 0.6), use_3d_nms=bool(rng.random() < 0.7), use_old_type_nms=bool(rng.random() < 0.3),
                  no_nms=bool(rng.random() < 0.1))
        if rng.random() < 0.4:
            kw["per_class_proposal"] = False
            kw["use_cls_confidence_only"] = bool(rng.random() < 0.5)
        out, tgt = synth.detection_batch(B=S, Q=Q, G=G, C=C, seed=int(rng.integers(1 << 30)), heading=float(rng.choice([0.0, np.pi])),
                                         max_gt=int(rng.integers(1, G + 1)))
        cfg = APC.get_ap_config_dict(dataset_config=Cfg(C), remove_empty_box=False, **kw)
        want, _ = oracle.ap_metrics(out["box_corners"], out["sem_cls_prob"], out["objectness_prob"], tgt["gt_box_corners"],
                                    tgt["gt_box_sem_cls_label"], tgt["gt_box_present"], C, ap_iou_thresh=thrs,
                                    config=oracle.default_ap_config(C, **kw))
        res = _apx_virtual_ranks(out, tgt, C, thrs, nranks, cap_total=int(rng.choice([1024, 4096])), rounds=1, cfg=cfg)
        for r, (ap, recall, ndet, ovf, _, _) in enumerate(res):
            assert ovf == 0
            for ti, thr in enumerate(thrs):
                for c in range(C):
                    k = "%d Average Precision" % c
                    if k not in want[thr]:      # a class nobody predicted and no GT has: the reference has no entry for it
                        continue
                    a, b_ = float(ap[ti, c]), float(want[thr][k])
                    if not (abs(a - b_) < 1e-9 or (np.isnan(a) and np.isnan(b_))):
                        pytest.fail(str(("AP MISMATCH", dict(it=it, rank=r, S=S, Q=Q, G=G, C=C, nranks=nranks, thrs=thrs, kw=kw, cls=c, got=a, want=b_))))
