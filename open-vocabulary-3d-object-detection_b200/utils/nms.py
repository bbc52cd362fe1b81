"""Drop-in for the reference's ``utils/nms.py`` (greedy NMS, fp64) on the GPU.

Same signatures and return values (``list[int]`` of picked indices in pick
order, utils/nms.py:43,79,120); the work runs in csrc/nms.cu (bitonic sort +
ballot bitmask + ordered warp scan).  ``nms_batch`` is the batched, device-side
form used by the AP and pseudo-label paths (one CTA per scene, no host sync).
Scores must be tie-free for the order to be defined (numpy's default argsort,
utils/nms.py:51,90,131, is not stable)."""
import numpy as np
import torch

from .. import _capi as C


def nms_batch(boxes, overlap_threshold, old_type=False, dims=3, samecls=False, counts=None, vol_eps=0.0, want_order=True, lhs=False):
    """boxes: CUDA fp64 [S,K,ncols] (2*dims coords, score, [cls]).  Returns
    (keep uint8 [S,K], pick_order int32 [S,K] (-1 padded) or None, npick int32 [S]).  ``want_order=False`` lets the
    class-wise kernel skip the global score sort (the keep mask is the same)."""
    C.require_cuda(boxes)
    b = boxes.detach().to(torch.float64).contiguous()
    S, K, ncols = b.shape
    keep = torch.empty((S, K), dtype=torch.uint8, device=b.device)
    order = torch.empty((S, K), dtype=torch.int32, device=b.device) if want_order else None
    npick = torch.zeros((S,), dtype=torch.int32, device=b.device)
    cnt = None if counts is None else torch.as_tensor(counts).to(device=b.device, dtype=torch.int32).contiguous()
    flags = (C.NMS_2D if dims == 2 else 0) | (C.NMS_SAMECLS if samecls else 0) | (C.NMS_OLD_TYPE if old_type else 0) | (C.NMS_LHS if lhs else 0)
    with torch.cuda.device(b.device):
        C.check(C.lib().ovdet_nms_f64(C.ptr(b), C.ptr(cnt), S, K, ncols, float(overlap_threshold), float(vol_eps), flags,
                                      C.ptr(keep), C.ptr(order), C.ptr(npick), C.stream(b.device)))
    return keep, order, npick


def _single(boxes, thr, old_type, dims, samecls):
    boxes = np.asarray(boxes)
    if boxes.shape[0] == 0:
        return []
    b = torch.as_tensor(np.ascontiguousarray(boxes, dtype=np.float64), device="cuda")[None]
    _, order, npick = nms_batch(b, thr, old_type, dims, samecls)
    n = int(npick[0].item())
    return [int(i) for i in order[0, :n].cpu().numpy()]


def nms_2d_faster(boxes, overlap_threshold, old_type=False):
    """utils/nms.py:43-76; boxes [K,5] = x1,y1,x2,y2,score."""
    return _single(boxes, overlap_threshold, old_type, 2, False)


def nms_3d_faster(boxes, overlap_threshold, old_type=False):
    """utils/nms.py:79-117; boxes [K,7] = x1,y1,z1,x2,y2,z2,score."""
    return _single(boxes, overlap_threshold, old_type, 3, False)


def nms_3d_faster_samecls(boxes, overlap_threshold, old_type=False):
    """utils/nms.py:120-162; boxes [K,8] = x1..z2,score,cls."""
    return _single(boxes, overlap_threshold, old_type, 3, True)
