"""Host-side logic that needs no GPU: result unpacking of the compact AP reducer, capacity arithmetic, metric
formatting (same keys / order / value types as the reference's APCalculator.compute_metrics, ap_calculator.py:370-395),
flag words of the GIoU dispatcher, scene sharding."""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ovdet_b200  # noqa: E402,F401
from ovdet_b200 import _capi as C  # noqa: E402
from ovdet_b200 import dist as D  # noqa: E402
from ovdet_b200.utils import ap_calculator as APC, box_util as BU, eval_det as ED  # noqa: E402


def test_pow2_at_least():
    assert ED._pow2_at_least(1) == 1024 and ED._pow2_at_least(1025) == 2048
    assert ED._pow2_at_least(200, 32) == 256 and ED._pow2_at_least(32, 32) == 32 and ED._pow2_at_least(33, 32) == 64


def test_unpack_compact_layouts():
    nthr, Cn = 2, 5
    ap = np.arange(nthr * Cn, dtype=np.float64).reshape(nthr, Cn) / 10
    rc = ap + 0.5
    nd = np.arange(Cn, dtype=np.int64) * 7
    # single-rank byte layout: ap | recall | n_det (int64) | overflow (int32)
    raw = np.concatenate([ap.ravel().view(np.uint8), rc.ravel().view(np.uint8), nd.view(np.uint8),
                          np.array([3, 0, 0, 0], np.int32).view(np.uint8)])
    a, r, ovf, n, mx = ED.unpack_compact(torch.from_numpy(raw.copy()), nthr, Cn, with_max=True)
    np.testing.assert_array_equal(a, ap); np.testing.assert_array_equal(r, rc); np.testing.assert_array_equal(n, nd)
    assert ovf == 3 and mx == -1
    # distributed fp64 layout: ap | recall | overflow | n_det | max count
    f = np.concatenate([ap.ravel(), rc.ravel(), [2.0], nd.astype(np.float64), [311.0]])
    a, r, ovf, n, mx = ED.unpack_compact(torch.from_numpy(f), nthr, Cn, with_max=True)
    np.testing.assert_array_equal(a, ap); np.testing.assert_array_equal(n, nd)
    assert ovf == 2 and mx == 311
    assert len(ED.unpack_compact(torch.from_numpy(f), nthr, Cn)) == 4


def test_metric_formatting_matches_reference_layout():
    class Cfg:
        num_semcls = 4
    calc = APC.APCalculator(Cfg(), exact_eval=False, class2type_map={0: "bed", 1: "chair", 2: "sofa", 3: "table"})
    ap = np.array([0.5, np.nan, 0.25, 1.0]); rc = np.array([0.9, 0.0, 0.4, 1.0])
    fast = calc._format_rows(ap, rc)
    slow = calc._format({k: ap[k] for k in range(4)}, {k: rc[k] for k in range(4)})
    assert isinstance(fast, OrderedDict) and list(fast.keys()) == list(slow.keys())
    assert list(fast.keys())[:5] == ["bed Average Precision", "chair Average Precision", "sofa Average Precision",
                                     "table Average Precision", "mAP"]
    for k in slow:
        assert type(fast[k]) is type(slow[k])
        assert fast[k] == slow[k] or (np.isnan(fast[k]) and np.isnan(slow[k]))
    assert fast["mAP"] == np.float32((0.5 + 0 + 0.25 + 1.0) / 4)   # NaN -> 0 before the fp32 mean (ap_calculator.py:385-386)
    text = calc.metrics_to_str({0.25: fast, 0.5: fast})
    assert "mAP0.25, mAP0.50: 43.75, 43.75" in text and "bed Average Precision: 50.00" in text


def test_metric_formatting_all_thresholds_equals_per_threshold():
    """compute_metrics / replay format all thresholds at once (_format_all): keys, order, values and value types of the
    reference's per-threshold _format (utils/ap_calculator.py:377-395)."""
    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 8, 9, 20, 37):
        class Cfg:
            num_semcls = n
        calc = APC.APCalculator(Cfg(), ap_iou_thresh=[0.25, 0.5, 0.75], exact_eval=False)
        for _ in range(40):
            ap = rng.random((3, n)) * rng.choice([1.0, 1e-3]); rc = rng.random((3, n))
            ap[rng.random((3, n)) < 0.15] = np.nan
            got = calc._format_all(ap, rc)
            assert isinstance(got, OrderedDict) and list(got.keys()) == [0.25, 0.5, 0.75]
            for ti, thr in enumerate([0.25, 0.5, 0.75]):
                want = calc._format({k: ap[ti, k] for k in range(n)}, {k: rc[ti, k] for k in range(n)})
                assert isinstance(got[thr], OrderedDict) and list(got[thr].keys()) == list(want.keys())
                for k in want:
                    assert type(got[thr][k]) is type(want[k]), (k, type(got[thr][k]), type(want[k]))
                    assert got[thr][k] == want[k] or (np.isnan(got[thr][k]) and np.isnan(want[k])), (n, thr, k)


def test_giou_flag_words():
    assert BU.giou_flags(True, False, "cython", True, "aabb") == (C.GIOU_ROTATED | C.GIOU_PREFILTER | C.GIOU_CLIP_F64)
    assert BU.giou_flags(False, True, "tensor", False, "aabb") == C.GIOU_INTER_ONLY
    assert BU.giou_flags(True, False, "tensor", True, "hull") & C.GIOU_ENCL_HULL


def test_shard_range_partitions():
    for n in (0, 1, 7, 5050):
        for w in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
