"""Mirror of the box-projection part of the reference's utils/image_util.py (the RegionCLIP crop branch,
criterion.py:380-391): ``SUNRGBD_Calibration_cuda`` (:247-298), ``project_box_3d_cuda`` (:117-134), ``rotz_cuda`` (:136-146)
and a batched form that also clips to the image.  One kernel per call (csrc/project.cu) instead of ~15 torch ops."""
import torch

from .. import _capi as C


def rotz_cuda(t):
    """utils/image_util.py:136-146 (plain torch; kept for interface parity)."""
    c, s = torch.cos(t), torch.sin(t)
    z, o = torch.zeros_like(t), torch.ones_like(t)
    return torch.stack([torch.stack([c, -s, z], -1), torch.stack([s, c, z], -1), torch.stack([z, z, o], -1)], -2)


class SUNRGBD_Calibration_cuda(object):
    """Holds Rtilt [3,3] and K [3,3] like the reference (:276-282); the projections run inside the kernel."""

    def __init__(self, Rtilt, K):
        self.Rtilt = Rtilt.float()
        self.K = K.float()
        self.f_u, self.f_v = self.K[0, 0], self.K[1, 1]
        self.c_u, self.c_v = self.K[0, 2], self.K[1, 2]


def project_boxes_3d(Rtilt, K, center, size, heading_angle, image_wh=None):
    """Batched: Rtilt/K [B,3,3], center/size [B,Q,3], heading [B,Q], image_wh [B,2] (width, height) or None ->
    boxes [B,Q,4] in the reference's (x1, y1, x2, y2) order, clipped to the image when image_wh is given
    (criterion.py:387-391)."""
    C.require_cuda(center, size, heading_angle, Rtilt, K)
    dev = center.device
    f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    ctr, sz, ang, R, Km = f(center), f(size), f(heading_angle), f(Rtilt), f(K)
    B, Q = ctr.shape[0], ctr.shape[1]
    assert R.shape == (B, 3, 3) and Km.shape == (B, 3, 3) and sz.shape == (B, Q, 3) and ang.shape == (B, Q)
    wh = None if image_wh is None else f(torch.as_tensor(image_wh))
    out = torch.empty((B, Q, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_project_box3d_f32(C.ptr(ctr), C.ptr(sz), C.ptr(ang), C.ptr(R), C.ptr(Km), C.ptr(wh), B, Q,
                                                C.ptr(out), C.stream(dev)))
    return out


def project_box_3d_cuda(calib, center, size, heading_angle):
    """utils/image_util.py:117-134: center/size [..., Q, 3], heading [..., Q] of ONE scene -> [..., Q, 4]."""
    lead = center.shape[:-1]
    ctr = center.reshape(1, -1, 3)
    out = project_boxes_3d(calib.Rtilt[None], calib.K[None], ctr, size.reshape(1, -1, 3), heading_angle.reshape(1, -1))
    return out.reshape(*lead, 4)
