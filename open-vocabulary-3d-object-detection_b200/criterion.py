"""Drop-in for ``criterion.Matcher`` (criterion.py:18-92).

The reference builds the cost with four elementwise kernels, copies it to the
host, loops over the batch calling scipy's LSAP and copies indices back
(criterion.py:58-86) -- twice per decoder layer with the GIoU call before it.
Here ``forward`` is two launches on the current stream: the fused cost kernel
(csrc/giou3d.cu epilogue, or the elementwise kernel when ``gious`` is supplied)
and one-warp-per-sample LSAP (csrc/lsap.cu).  ``match_from_boxes`` fuses the GIoU
itself, so a decoder layer costs two kernels and no host round trip.
"""
import torch
import torch.nn as nn

from . import _capi as C
from .utils.box_util import DEFAULT_K2_CAP, giou_flags


def matcher_cost(sem_cls_prob, objectness_prob, gt_labels, weights, center_dist=None, gious=None,
                 center_q=None, center_g=None, corners1=None, corners2=None, nactual_gt=None,
                 flags=0, k2_cap=0, want_gious=False):
    """weights = (cost_class, cost_objectness, cost_center, cost_giou).  Returns (cost [B,Q,G], gious or None)."""
    C.require_cuda(sem_cls_prob)
    dev = sem_cls_prob.device
    f32 = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()
    prob, obj = f32(sem_cls_prob), f32(objectness_prob)
    cd, gi, cq, cg, c1, c2 = f32(center_dist), f32(gious), f32(center_q), f32(center_g), f32(corners1), f32(corners2)
    lab = gt_labels.detach().to(device=dev, dtype=torch.int64).contiguous()
    nk = None if nactual_gt is None else nactual_gt.detach().to(device=dev, dtype=torch.int64).contiguous()
    B, Q, Cn = prob.shape
    G = lab.shape[1]
    cost = torch.empty((B, Q, G), dtype=torch.float32, device=dev)
    gout = torch.empty((B, Q, G), dtype=torch.float32, device=dev) if (want_gious and gi is None) else None
    wc, wo, wce, wg = [float(w) for w in weights]
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_matcher_cost_f32(C.ptr(prob), C.ptr(obj), C.ptr(cd), C.ptr(cq), C.ptr(cg), C.ptr(gi),
                                               C.ptr(c1), C.ptr(c2), C.ptr(lab), C.ptr(nk), B, Q, G, Cn, wc, wo, wce, wg,
                                               int(flags), int(k2_cap), C.ptr(gout), C.ptr(cost), C.stream(dev)))
    return cost, (gi if gi is not None else gout)


def lsap(cost, nactual_gt):
    """Per-sample assignment on cost[b, :, :nactual_gt[b]].  Returns
    (per_prop_gt_inds int64 [B,Q], proposal_matched_mask fp32 [B,Q], col_to_row int32 [B,G]).
    A sample whose cost slab holds NaN / -inf, or that has no feasible assignment, is not solved: its mask is -1
    everywhere and its col_to_row row -2 (scipy raises ValueError there; ``Matcher`` does too as soon as the
    assignments are read on the host -- the stream-ordered outputs cannot raise by themselves)."""
    C.require_cuda(cost)
    dev = cost.device
    cost = cost.detach().to(torch.float32).contiguous()
    B, Q, G = cost.shape
    nk = nactual_gt.detach().to(device=dev, dtype=torch.int64).contiguous()
    inds = torch.empty((B, Q), dtype=torch.int64, device=dev)
    mask = torch.empty((B, Q), dtype=torch.float32, device=dev)
    c2r = torch.empty((B, G), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_lsap_f32(C.ptr(cost), C.ptr(nk), B, Q, G, C.ptr(inds), C.ptr(mask), C.ptr(c2r), C.stream(dev)))
    return inds, mask, c2r


def _assignments(c2r, nactual_gt, device):
    """The reference's per-sample [rows, cols] LongTensors (criterion.py:79-86; scipy
    returns rows ascending).  One small D2H of the [B,G] column->row table."""
    c2r_h = c2r.cpu()
    if (c2r_h == -2).any():
        # what scipy.optimize.linear_sum_assignment does at criterion.py:79 on such a cost matrix
        raise ValueError("matrix contains invalid numeric entries (sample %d): NaN / -inf in the matcher cost, or an "
                         "infeasible assignment" % int((c2r_h == -2).any(1).nonzero()[0]))
    n_h = nactual_gt.cpu().tolist()
    out = []
    for b, n in enumerate(n_h):
        if n > 0:
            rows = c2r_h[b, :n].to(torch.int64)
            cols = torch.arange(n, dtype=torch.int64)
            ok = rows >= 0
            rows, cols = rows[ok], cols[ok]
            order = torch.argsort(rows)
            out.append([rows[order].to(device), cols[order].to(device)])
        else:
            out.append([])
    return out


class Matcher(nn.Module):
    def __init__(self, cost_class, cost_objectness, cost_giou, cost_center):
        super().__init__()
        self.cost_class = cost_class
        self.cost_objectness = cost_objectness
        self.cost_giou = cost_giou
        self.cost_center = cost_center

    def _weights(self):
        return (self.cost_class, self.cost_objectness, self.cost_center, self.cost_giou)

    @torch.no_grad()
    def forward(self, outputs, targets, return_assignments=True):
        """criterion.py:33-92: needs outputs["sem_cls_prob","objectness_prob","center_dist","gious"],
        targets["gt_box_sem_cls_label","nactual_gt"]."""
        cost, _ = matcher_cost(outputs["sem_cls_prob"], outputs["objectness_prob"], targets["gt_box_sem_cls_label"],
                               self._weights(), center_dist=outputs["center_dist"], gious=outputs["gious"])
        return self._solve(cost, targets["nactual_gt"], return_assignments)

    @torch.no_grad()
    def match_from_boxes(self, outputs, targets, rotated_boxes=True, needs_grad=False, return_assignments=True,
                         k2_cap=None, prefilter=True):
        """GIoU (criterion.py:348-355) + L1 centre distance (:357-359) + cost (:40-63) in one
        kernel, then LSAP.  Also stores outputs["gious"] like single_output_forward does."""
        mode = "tensor" if needs_grad else "cython"
        if k2_cap is None:
            k2_cap = DEFAULT_K2_CAP if mode == "cython" else 0
        flags = giou_flags(rotated_boxes, False, mode, prefilter, "aabb")
        prob = outputs["sem_cls_prob"]
        C.require_cuda(prob)
        dev = prob.device
        f32 = lambda t: C.as_input(t, torch.float32, dev)
        prob, obj, cq, cg = f32(prob), f32(outputs["objectness_prob"]), f32(outputs["center_normalized"]), f32(targets["gt_box_centers_normalized"])
        c1, c2 = f32(outputs["box_corners"]), f32(targets["gt_box_corners"])
        lab = C.as_input(targets["gt_box_sem_cls_label"], torch.int64, dev)
        nk = C.as_input(targets["nactual_gt"], torch.int64, dev)
        B, Q, Cn = prob.shape
        G = lab.shape[1]
        # one allocation for the three fp32 [B,Q,G]-sized / [B,Q] outputs, one for the integer ones
        fbuf = torch.empty((2 * B * Q * G + B * Q,), dtype=torch.float32, device=dev)
        cost, gious, mask = fbuf[:B * Q * G].view(B, Q, G), fbuf[B * Q * G:2 * B * Q * G].view(B, Q, G), fbuf[2 * B * Q * G:].view(B, Q)
        inds = torch.empty((B, Q), dtype=torch.int64, device=dev)
        c2r = torch.empty((B, G), dtype=torch.int32, device=dev)
        wc, wo, wce, wg = [float(w) for w in self._weights()]
        with C.on_device(dev):
            C.check(C.lib().ovdet_matcher_step_f32(prob.data_ptr(), obj.data_ptr(), cq.data_ptr(), cg.data_ptr(), c1.data_ptr(), c2.data_ptr(),
                                                   lab.data_ptr(), nk.data_ptr(), B, Q, G, Cn, wc, wo, wce, wg, int(flags), int(k2_cap),
                                                   gious.data_ptr(), cost.data_ptr(), inds.data_ptr(), mask.data_ptr(), c2r.data_ptr(), C.stream(dev)))
        outputs["gious"] = gious
        ret = {"per_prop_gt_inds": inds, "proposal_matched_mask": mask, "final_cost": cost}
        if return_assignments:
            ret["assignments"] = _assignments(c2r, nk, dev)
        return ret

    def _solve(self, cost, nactual_gt, return_assignments):
        inds, mask, c2r = lsap(cost, nactual_gt)
        ret = {"per_prop_gt_inds": inds, "proposal_matched_mask": mask, "final_cost": cost}
        if return_assignments:
            ret["assignments"] = _assignments(c2r, nactual_gt, cost.device)
        return ret
