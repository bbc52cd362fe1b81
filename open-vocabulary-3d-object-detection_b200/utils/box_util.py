"""Drop-in for the hot functions of the reference's ``utils/box_util.py``.

``generalized_box3d_iou`` keeps the reference signature (utils/box_util.py:717-724)
and dispatch semantics, but every variant runs as ONE sm_100a kernel
(csrc/giou3d.cu) instead of ~30 torch kernels, 2B+1 host syncs, two PCIe round
trips and a Cython triple loop.

Reference quirks are explicit switches, defaulting to the reference's behaviour:

* ``needs_grad=False`` (the ``no_grad`` Cython path, :731-737) -> fp64 clip,
  fp32 polygon/area, and the shipped ``K2 = rect2.shape[2]`` bug
  (utils/box_intersection.pyx:180): only GT columns < 4 are clipped.  Override
  with ``k2_cap=0`` (no cap) or set ``ovdet_b200.utils.box_util.DEFAULT_K2_CAP``.
* ``needs_grad=True`` (the TorchScript path, :725-730) -> all-fp32 clip, no cap.  When ``corners1`` requires
  grad the call is differentiable (``_GIoU3D``: forward kernel + the sparse backward kernel of csrc/giou3d_bwd.cu,
  SURVEY.md 8f-4) and, like the reference's ``with torch.enable_grad()``, tracks gradients even inside an outer
  ``no_grad``.  No gradient w.r.t. ``corners2`` (the ground truth carries none in criterion.py:274-296).
* ``prefilter`` (default True): skip pairs whose axis-aligned BEV overlap is 0
  (:587-588), which is wrong for rotated boxes but is what the reference does.
* ``enclosing``: "aabb" (live reference, :466-514) or "hull" (utils/box_ops3d.py:533-571).
"""
import numpy as np
import torch

from .. import _capi as C

DEFAULT_K2_CAP = 4  # the reference as shipped (utils/box_intersection.pyx:180)


def _check_shapes(corners1, corners2):
    # utils/box_util.py:639-645
    assert len(corners1.shape) == 4
    assert len(corners2.shape) == 4
    assert corners1.shape[2] == 8
    assert corners1.shape[3] == 3
    assert corners1.shape[0] == corners2.shape[0]
    assert corners1.shape[2] == corners2.shape[2]
    assert corners1.shape[3] == corners2.shape[3]


def giou_flags(rotated_boxes, return_inter_vols_only, mode, prefilter, enclosing):
    flags = 0
    if rotated_boxes:
        flags |= C.GIOU_ROTATED
    if prefilter:
        flags |= C.GIOU_PREFILTER
    if return_inter_vols_only:
        flags |= C.GIOU_INTER_ONLY
    if mode == "cython":
        flags |= C.GIOU_CLIP_F64
    else:
        assert mode == "tensor", mode
    if enclosing == "hull":
        flags |= C.GIOU_ENCL_HULL
    else:
        assert enclosing == "aabb", enclosing
    return flags


class _GIoU3D(torch.autograd.Function):
    """Differentiable fp32 torch-path GIoU (utils/box_util.py:517-618): forward = the GIoU kernel, backward = the sparse
    per-pair backward kernel (csrc/giou3d_bwd.cu).  Gradient w.r.t. corners1 only (the GT boxes carry none in
    criterion.py:274-296)."""

    @staticmethod
    def forward(ctx, corners1, corners2, nums_k2, rotated_boxes, prefilter):
        c1 = corners1.detach().to(torch.float32).contiguous()
        c2 = corners2.detach().to(torch.float32).contiguous()
        nk = None if nums_k2 is None else nums_k2.detach().to(device=c1.device, dtype=torch.int64).contiguous()
        B, K1, K2 = c1.shape[0], c1.shape[1], c2.shape[1]
        flags = giou_flags(rotated_boxes, False, "tensor", prefilter, "aabb")
        out = torch.empty((B, K1, K2), dtype=torch.float32, device=c1.device)
        with torch.cuda.device(c1.device):
            C.check(C.lib().ovdet_giou3d_f32(C.ptr(c1), C.ptr(c2), C.ptr(nk), B, K1, K2, 0, flags, C.ptr(out), C.stream(c1.device)))
        ctx.save_for_backward(c1, c2, nk if nk is not None else torch.empty(0, device=c1.device))
        ctx.has_nk, ctx.flags = nk is not None, flags
        return out

    @staticmethod
    def backward(ctx, grad_out):
        c1, c2, nk = ctx.saved_tensors
        B, K1, K2 = c1.shape[0], c1.shape[1], c2.shape[1]
        go = grad_out.detach().to(torch.float32).contiguous()
        g1 = torch.empty_like(c1)
        with torch.cuda.device(c1.device):
            C.check(C.lib().ovdet_giou3d_backward_f32(C.ptr(c1), C.ptr(c2), C.ptr(nk) if ctx.has_nk else None, C.ptr(go), B, K1, K2,
                                                      ctx.flags, C.ptr(g1), C.stream(c1.device)))
        return g1, None, None, None, None


def generalized_box3d_iou(corners1, corners2, nums_k2, rotated_boxes=True, return_inter_vols_only=False,
                          needs_grad=False, *, mode=None, prefilter=True, k2_cap=None, enclosing="aabb", out=None):
    """[B,K1,8,3] x [B,K2,8,3] -> [B,K1,K2] fp32 on ``corners1.device``
    (utils/box_util.py:717-737).  CUDA tensors run stream-ordered with no host
    sync; CPU tensors go through the host-buffer C entry point (H2D, kernel, D2H)."""
    _check_shapes(corners1, corners2)
    if needs_grad and (corners1.requires_grad or corners2.requires_grad):
        # the reference's TorchScript path runs under `with torch.enable_grad()` (:725-730): differentiable w.r.t. the
        # predicted corners even when the caller sits inside no_grad
        if corners2.requires_grad:
            raise NotImplementedError("no gradient w.r.t. corners2 (ground truth) is produced")
        if return_inter_vols_only or enclosing != "aabb" or (mode not in (None, "tensor")) or not corners1.is_cuda or k2_cap:
            raise NotImplementedError("the backward exists for the fp32 torch-path GIoU on CUDA tensors only")
        nk_t = None if nums_k2 is None else torch.as_tensor(nums_k2)
        with torch.enable_grad():
            return _GIoU3D.apply(corners1, corners2, nk_t, bool(rotated_boxes), bool(prefilter))
    if mode is None:
        mode = "tensor" if needs_grad else "cython"
    if k2_cap is None:
        k2_cap = DEFAULT_K2_CAP if mode == "cython" else 0
    flags = giou_flags(rotated_boxes, return_inter_vols_only, mode, prefilter, enclosing)
    B, K1, K2 = corners1.shape[0], corners1.shape[1], corners2.shape[1]
    dev = corners1.device
    c1 = C.as_input(corners1, torch.float32, dev)
    L = C.lib()
    if c1.is_cuda:
        C.require_cuda(corners2)
        c2 = C.as_input(corners2, torch.float32, dev)
        nk = None
        if nums_k2 is not None:
            nk = C.as_input(nums_k2 if isinstance(nums_k2, torch.Tensor) else torch.as_tensor(nums_k2), torch.int64, dev)
            assert nk.numel() == B
        if out is None:
            out = torch.empty((B, K1, K2), dtype=torch.float32, device=dev)
        else:  # caller-owned result buffer
            assert out.shape == (B, K1, K2) and out.dtype == torch.float32 and out.is_contiguous() and out.device == dev
        with C.on_device(dev):
            C.check(L.ovdet_giou3d_f32(c1.data_ptr(), c2.data_ptr(), None if nk is None else nk.data_ptr(), B, K1, K2, int(k2_cap), flags,
                                       out.data_ptr(), C.stream(dev)))
        return out
    # host buffers: everything the C entry point sees must live on the host and outlive the call (locals, not temporaries)
    cpu = torch.device("cpu")
    c2 = C.as_input(corners2, torch.float32, cpu)
    nk = None
    if nums_k2 is not None:
        nk = C.as_input(nums_k2 if isinstance(nums_k2, torch.Tensor) else torch.as_tensor(nums_k2), torch.int64, cpu)
        assert nk.numel() == B
    if out is None:
        out = torch.empty((B, K1, K2), dtype=torch.float32)
    else:  # e.g. pinned host memory
        assert out.shape == (B, K1, K2) and out.dtype == torch.float32 and out.is_contiguous() and not out.is_cuda
    C.check(L.ovdet_giou3d_host_f32(c1.data_ptr(), c2.data_ptr(), None if nk is None else nk.data_ptr(), B, K1, K2, int(k2_cap), flags,
                                    out.data_ptr()))
    return out


def generalized_box3d_iou_tensor(corners1, corners2, nums_k2, rotated_boxes=True, return_inter_vols_only=False):
    """utils/box_util.py:517-618 (values only)."""
    return generalized_box3d_iou(corners1, corners2, nums_k2, rotated_boxes, return_inter_vols_only,
                                 mode="tensor", k2_cap=0)


def generalized_box3d_iou_cython(corners1, corners2, nums_k2, rotated_boxes=True, return_inter_vols_only=False):
    """utils/box_util.py:624-714, as shipped."""
    return generalized_box3d_iou(corners1, corners2, nums_k2, rotated_boxes, return_inter_vols_only, mode="cython")


def box3d_iou_batch(dets, gts, nd=None, ng=None, want_2d=False):
    """Exact fp64 IoU matrices: dets [S,D,8,3] x gts [S,G,8,3] -> [S,D,G] fp64
    (box3d_iou, utils/box_util.py:116-141, on fp32 corners cast to fp64)."""
    C.require_cuda(dets, gts)
    S, D, G = dets.shape[0], dets.shape[1], gts.shape[1]
    d = dets.detach().to(torch.float32).contiguous()
    g = gts.detach().to(torch.float32).contiguous()
    out = torch.empty((S, D, G), dtype=torch.float64, device=d.device)
    out2 = torch.empty((S, D, G), dtype=torch.float64, device=d.device) if want_2d else None
    ndt = None if nd is None else torch.as_tensor(nd).to(device=d.device, dtype=torch.int32).contiguous()
    ngt = None if ng is None else torch.as_tensor(ng).to(device=d.device, dtype=torch.int32).contiguous()
    with torch.cuda.device(d.device):
        C.check(C.lib().ovdet_box3d_iou_f64(C.ptr(d), C.ptr(g), C.ptr(ndt), C.ptr(ngt), S, D, G, C.ptr(out), C.ptr(out2),
                                            C.stream(d.device)))
    return (out, out2) if want_2d else out


def box3d_iou(corners1, corners2):
    """utils/box_util.py:116-141: (8,3) x (8,3) -> (iou, iou_2d) python floats.
    One pair per call is a poor use of a GPU; the AP path uses box3d_iou_batch."""
    dev = torch.device("cuda")
    a = torch.as_tensor(np.asarray(corners1), dtype=torch.float32, device=dev).reshape(1, 1, 8, 3)
    b = torch.as_tensor(np.asarray(corners2), dtype=torch.float32, device=dev).reshape(1, 1, 8, 3)
    iou, iou2 = box3d_iou_batch(a, b, want_2d=True)
    return float(iou.item()), float(iou2.item())


def box_parametrization_to_corners(box_center_unnorm, box_size, box_angle):
    """datasets/sunrgbd.py:145-148 / scannet.py:138-141: (centre in the depth frame [...,3], size l/w/h [...,3],
    heading [...]) -> corners [...,8,3] in the upright-camera frame; one kernel instead of ~15 torch ops."""
    C.require_cuda(box_center_unnorm)
    dev = box_center_unnorm.device
    shape = box_angle.shape
    ctr = box_center_unnorm.detach().to(torch.float32).reshape(-1, 3).contiguous()
    sz = box_size.detach().to(device=dev, dtype=torch.float32).reshape(-1, 3).contiguous()
    ang = box_angle.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
    n = ang.numel()
    out = torch.empty((n, 8, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_box_corners_f32(C.ptr(ctr), C.ptr(sz), C.ptr(ang), n, C.ptr(out), C.stream(dev)))
    return out.reshape(*shape, 8, 3)


def get_3d_box_batch_tensor(box_size, angle, center):
    """utils/box_util.py:313-352: centre already in the upright-camera frame."""
    depth_center = torch.stack([center[..., 0], center[..., 2], -center[..., 1]], -1)   # undo flip_axis_to_camera
    return box_parametrization_to_corners(depth_center, box_size, angle)


def generalized_box3d_iou_from_params(center1, size1, angle1, corners2, nums_k2, rotated_boxes=True,
                                      return_inter_vols_only=False, needs_grad=False, *, mode=None, prefilter=True,
                                      k2_cap=None, enclosing="aabb", return_corners=False):
    """generalized_box3d_iou with the query boxes still in (centre, size, heading) form -- the decode of
    models/model_3detr.py:279 is fused into the GIoU kernel's load stage."""
    C.require_cuda(center1, corners2)
    if mode is None:
        mode = "tensor" if needs_grad else "cython"
    if k2_cap is None:
        k2_cap = DEFAULT_K2_CAP if mode == "cython" else 0
    flags = giou_flags(rotated_boxes, return_inter_vols_only, mode, prefilter, enclosing)
    dev = center1.device
    B, K1 = center1.shape[0], center1.shape[1]
    K2 = corners2.shape[1]
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    ctr, sz, ang, c2 = f32(center1), f32(size1), f32(angle1), f32(corners2)
    nk = None if nums_k2 is None else torch.as_tensor(nums_k2).detach().to(device=dev, dtype=torch.int64).contiguous()
    out = torch.empty((B, K1, K2), dtype=torch.float32, device=dev)
    c1o = torch.empty((B, K1, 8, 3), dtype=torch.float32, device=dev) if return_corners else None
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_giou3d_decode_f32(C.ptr(ctr), C.ptr(sz), C.ptr(ang), C.ptr(c2), C.ptr(nk), B, K1, K2,
                                                int(k2_cap), flags, C.ptr(out), C.ptr(c1o), C.stream(dev)))
    return (out, c1o) if return_corners else out
