"""Drop-in for the reference's ``utils/label_formatter.py`` (pseudo-label generation).

``step`` packs (centre, size, label, score, objectness, scan_idx) rows (:81-106), ``compute`` applies the
per-class thresholds (:117-132), ``gen_pseudo`` keeps a box iff the mode of the LSeg point labels inside
its extent equals its label (:134-167).  The per-box point crop + label vote, which the reference runs
as numpy masks over the whole cloud per box in a multiprocessing pool, is one kernel (csrc/points.cu);
file formats (``<scan>.npy`` in, ``<scan>_bbox.npy`` fp64 [N,7] out) are unchanged."""
import os

import numpy as np
import torch

from .. import _capi as C
from .box_3d_utils import box_3d_iou  # noqa: F401  (same function, label_formatter.py:10-64)


def box_label_mode(points, labels, boxes, ignore_label=-100):
    """points [N,3], labels [N], boxes [M,>=6] (centre, size) -> (mode int32 [M], count int32 [M])."""
    dev = torch.device("cuda")
    p = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float64), device=dev)
    l = torch.as_tensor(np.ascontiguousarray(labels, dtype=np.float64), device=dev)
    b = torch.as_tensor(np.ascontiguousarray(boxes, dtype=np.float64), device=dev)
    M = b.shape[0]
    mode = torch.empty((M,), dtype=torch.int32, device=dev)
    cnt = torch.empty((M,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_box_label_mode(C.ptr(p), C.ptr(l), p.shape[0], C.ptr(b), b.shape[1], M, float(ignore_label),
                                             C.ptr(mode), C.ptr(cnt), C.stream(dev)))
    return mode.cpu().numpy(), cnt.cpu().numpy()


class LabelFormatter():
    def __init__(self, box_path, output_path, label_path, scene_list) -> None:
        self.boxes = []
        self.pseudo_box_dir = box_path
        self.output_path = output_path
        self.scene_list = scene_list
        self.raw_label_path = os.path.join(label_path, "{}.npy")
        self.IGNORE_LABEL = -100
        self.nyu40ids = np.array([3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39])
        self.nyu40id2class = {nyu40id: i for i, nyu40id in enumerate(list(self.nyu40ids))}

    def step(self, outputs, batch_data_label):
        """label_formatter.py:81-106: rows = centre(3) size(3) label score objectness scan_idx."""
        sem_cls_prob = outputs["sem_cls_prob"]
        obj_prob = outputs["objectness_prob"]
        B, Q, _ = sem_cls_prob.size()
        center = outputs["center_unnormalized"]
        size = outputs["size_unnormalized"]
        score, label = torch.max(sem_cls_prob.float(), dim=-1)
        scan = torch.repeat_interleave(batch_data_label["scan_idx"][:, None], Q, dim=1).to(score.device)
        boxes = torch.cat([center, size, torch.stack([label.to(score.dtype), score, obj_prob.to(score.dtype),
                                                      scan.to(score.dtype)], -1)], dim=-1).view(B * Q, 10)
        self.boxes.append(boxes.cpu().numpy())

    def compute(self, k, th_s, th_o):
        """label_formatter.py:117-132 (the top-k of :128-130 is commented out in the reference)."""
        self.boxes = np.concatenate(self.boxes, 0)
        out = []
        for label in range(18):
            boxes = self.boxes[self.boxes[:, 6] == label]
            out.append(boxes[np.logical_and(boxes[:, 7] >= th_s, boxes[:, 8] >= th_o)])
        self.pseudo_boxes = np.concatenate(out, 0)

    def gen_pseudo(self, idx):
        """label_formatter.py:134-167."""
        scan_name = self.scene_list[idx]
        raw_pc_data = np.load(self.raw_label_path.format(scan_name))
        point_clouds = raw_pc_data[:, :3]
        sem_seg_labels = self.project_label(raw_pc_data[:, 3], True)
        instance_bboxes = np.zeros((0, 7))
        mask = self.pseudo_boxes[:, -1] == idx
        numBox = int(mask.sum())
        if numBox > 0:
            boxes = self.pseudo_boxes[mask]
            assert (boxes[:, 6] >= 0).all()
            mode, cnt = box_label_mode(point_clouds, sem_seg_labels, boxes, self.IGNORE_LABEL)
            keep = (cnt > 0) & (mode == boxes[:, 6])
            filtered = boxes[keep]
            if len(filtered) > 0:
                instance_bboxes = np.concatenate([instance_bboxes[:, :7], filtered[:, :7]], 0)
            numBox = len(filtered)
        np.save(os.path.join(self.output_path, scan_name) + "_bbox.npy", instance_bboxes)
        return numBox

    def save(self, distributed=False):
        """label_formatter.py:169-174.  The reference maps gen_pseudo over the scans with ``mp.Pool(cpu_count())``; here each
        scan is one kernel launch, and with ``distributed=True`` the scans are cut across the ranks (``dist.shard_range``,
        no data-path collective) and the kept-box counts are summed with one all-reduce."""
        from ..dist import shard_range, all_reduce_count, is_distributed
        n = len(self.scene_list)
        lo, hi = shard_range(n) if (distributed and is_distributed()) else (0, n)
        l = sum(self.gen_pseudo(i) for i in range(lo, hi))
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        return all_reduce_count(l, dev) if distributed else l

    def process(self, k, th_s, th_o, distributed=False):
        """label_formatter.py:176-179 -- the entry point generate_pseudo_label.py:209 calls with
        (args.topk, args.conf_thresh, args.obj_thresh).  Returns the number of boxes acquired (the reference only prints it)."""
        self.compute(k, th_s, th_o)
        l = self.save(distributed=distributed)
        print("Done! Acquired {} boxes.".format(l))
        return l

    def crop_pc(self, pc, box):
        """label_formatter.py:183-188 (host mask; kept for interface parity)."""
        mask1 = np.prod(pc >= box[0:3] - box[3:6] / 2, axis=-1, keepdims=False)
        mask2 = np.prod(pc <= box[0:3] + box[3:6] / 2, axis=-1, keepdims=False)
        return (mask1 * mask2).astype("bool")

    def project_label(self, semantic_labels, PSEUDO_FLAG):
        """label_formatter.py:190-205."""
        if not PSEUDO_FLAG:
            sem_seg_labels = np.ones_like(semantic_labels) * self.IGNORE_LABEL
            for _c in self.nyu40ids:
                sem_seg_labels[semantic_labels == _c] = self.nyu40id2class[_c]
        else:
            sem_seg_labels = semantic_labels
            sem_seg_labels[semantic_labels >= 18] = self.IGNORE_LABEL
        return sem_seg_labels
