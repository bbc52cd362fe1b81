"""remove_empty_box of the reference's parse_predictions (utils/ap_calculator.py:70-84): a predicted
box is kept only if at least 5 scene points lie inside it; if a scene keeps no box, the box with the
highest objectness is kept (:83-84).  The reference does K Delaunay builds + find_simplex over
20-40k points per scene on the CPU; here it is one streaming kernel (csrc/points.cu)."""
import torch

from .. import _capi as C


def points_in_boxes_count(point_cloud, corners):
    """point_cloud [B,N,>=3] (depth frame), corners [B,K,8,3] (upright camera frame) -> int32 [B,K]."""
    C.require_cuda(corners)
    dev = corners.device
    pc = point_cloud.detach().to(device=dev, dtype=torch.float32).contiguous()
    cr = corners.detach().to(torch.float32).contiguous()
    B, N, stride = pc.shape
    K = cr.shape[1]
    counts = torch.empty((B, K), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_points_in_boxes_count(C.ptr(pc), B, N, stride, C.ptr(cr), K, C.ptr(counts), C.stream(dev)))
    return counts


def nonempty_box_mask(corners, point_cloud, objectness_probs, min_points=5):
    """uint8 [B,K] mask as ap_calculator.py:70-84 builds it (device-side, no host sync)."""
    counts = points_in_boxes_count(point_cloud, corners)
    mask = counts >= min_points
    none = ~mask.any(dim=1)
    best = objectness_probs.to(mask.device).argmax(dim=1)
    rows = torch.arange(mask.shape[0], device=mask.device)
    mask[rows, best] = mask[rows, best] | none
    return mask.to(torch.uint8)
