"""Seeded synthetic inputs in the shapes of BASELINE.json's configs.

Everything is generated with ``torch.Generator().manual_seed(seed)`` on the CPU
in fp32 and returned as CPU tensors; callers move them.  Boxes are sampled in
the upright-depth frame and converted with the reference's corner convention
(``box_parametrization_to_corners``: datasets/sunrgbd.py:145-148 ->
utils/box_util.py:288-352): x = +-l/2 pattern (+,+,-,-)x2, y = +h/2 x4 then
-h/2 x4, z = (+,-,-,+)x2 * w/2, rotation about Y, centre (x,y,z)->(x,-z,y).
"""
import math

import torch


def params_to_corners(center, size, angle):
    """(centre [..,3] depth frame, size [..,3] = l,w,h, heading [..]) -> corners [..,8,3]
    in the upright-camera frame, the ordering every kernel relies on
    (utils/box_util.py:313-352 via datasets/sunrgbd.py:145-148)."""
    cx, cy, cz = center[..., 0], -center[..., 2], center[..., 1]
    l, w, h = size[..., 0:1] / 2, size[..., 1:2] / 2, size[..., 2:3] / 2
    sx = torch.tensor([1, 1, -1, -1, 1, 1, -1, -1], dtype=size.dtype)
    sy = torch.tensor([1, 1, 1, 1, -1, -1, -1, -1], dtype=size.dtype)
    sz = torch.tensor([1, -1, -1, 1, 1, -1, -1, 1], dtype=size.dtype)
    x, y, z = l * sx, h * sy, w * sz
    c, s = torch.cos(angle)[..., None], torch.sin(angle)[..., None]
    X = c * x + s * z + cx[..., None]
    Y = y + cy[..., None]
    Z = -s * x + c * z + cz[..., None]
    return torch.stack([X, Y, Z], -1).contiguous()


def _u(g, shape, lo, hi):
    return torch.rand(shape, generator=g, device=g.device) * (hi - lo) + lo


def sample_boxes(g, shape, room="sunrgbd", heading=math.pi):
    """centre/size/heading per SURVEY 8d.  room 'sunrgbd': x,y~U(-3,3), z~U(-1,1);
    'scannet': x,y~U(0,8), z~U(0,2.5).  size ~U(0.3,1.8); heading ~U(-h,h)."""
    if room == "sunrgbd":
        xy = _u(g, (*shape, 2), -3.0, 3.0)
        z = _u(g, (*shape, 1), -1.0, 1.0)
    else:
        xy = _u(g, (*shape, 2), 0.0, 8.0)
        z = _u(g, (*shape, 1), 0.0, 2.5)
    center = torch.cat([xy, z], -1)
    size = _u(g, (*shape, 3), 0.3, 1.8)
    ang = _u(g, shape, -heading, heading) if heading > 0 else torch.zeros(shape, device=g.device)
    return center, size, ang


def detection_batch(B, Q, G=64, C=20, seed=0, room="sunrgbd", heading=math.pi, max_gt=64):
    """One batch in the reference's `outputs`/`targets` layout.

    GT: nactual ~ U{1..max_gt}, padded to G with zero boxes.  Predictions: the
    first nactual queries are jittered copies of the GT (centre + N(0,.15),
    size x clip(N(1,.1),.5,1.5), heading + N(0,.1)), the rest fresh samples.
    Class probs softmax(N(0,1)[Q,C+1]); labels ~ U{0..C-1}."""
    g = torch.Generator().manual_seed(seed)
    nactual = torch.randint(1, max_gt + 1, (B,), generator=g)
    gc, gs, ga = sample_boxes(g, (B, G), room, heading)
    present = (torch.arange(G)[None, :] < nactual[:, None])
    pc, ps, pa = sample_boxes(g, (B, Q), room, heading)
    n = min(Q, G)
    jc = gc[:, :n] + torch.randn((B, n, 3), generator=g) * 0.15
    js = gs[:, :n] * (torch.randn((B, n, 3), generator=g) * 0.1 + 1.0).clamp(0.5, 1.5)
    ja = ga[:, :n] + (torch.randn((B, n), generator=g) * 0.1 if heading > 0 else 0.0)
    m = present[:, :n]
    pc[:, :n] = torch.where(m[..., None], jc, pc[:, :n])
    ps[:, :n] = torch.where(m[..., None], js, ps[:, :n])
    pa[:, :n] = torch.where(m, ja, pa[:, :n])
    gt_corners = params_to_corners(gc, gs, ga) * present[..., None, None]
    box_corners = params_to_corners(pc, ps, pa)
    logits = torch.randn((B, Q, C + 1), generator=g)
    prob = torch.softmax(logits, -1)
    labels = torch.randint(0, C, (B, G), generator=g) * present
    lo = torch.tensor([-3.5, -3.5, -1.5]) if room == "sunrgbd" else torch.tensor([-0.5, -0.5, -0.5])
    hi = torch.tensor([3.5, 3.5, 1.5]) if room == "sunrgbd" else torch.tensor([8.5, 8.5, 3.0])
    outputs = {
        "box_corners": box_corners,
        "sem_cls_prob": prob[..., :-1].contiguous(),
        "objectness_prob": (1 - prob[..., -1]).contiguous(),
        "center_unnormalized": pc,
        "size_unnormalized": ps,
        "angle_continuous": pa,
        "center_normalized": ((pc - lo) / (hi - lo)).contiguous(),
    }
    targets = {
        "gt_box_corners": gt_corners.contiguous(),
        "gt_box_sem_cls_label": labels.long(),
        "gt_box_present": present.float(),
        "gt_box_angles": ga * present,
        "gt_box_centers": gc * present[..., None],
        "gt_box_sizes": gs * present[..., None],
        "gt_box_centers_normalized": (((gc - lo) / (hi - lo)) * present[..., None]).contiguous(),
        "nactual_gt": nactual.long(),
    }
    return outputs, targets


def clip_logits_inputs(M=8192, K=640, N=1203, seed=0):
    """C4: x ~ N(0,1)[M,K]; T = normalize(N(0,1)[N,K]) (unit-norm rows, as
    3DOVDet_tools/extract_class_features.py:28-30 produces); both bf16."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((M, K), generator=g)
    t = torch.nn.functional.normalize(torch.randn((N, K), generator=g), dim=-1)
    return x.to(torch.bfloat16), t.to(torch.bfloat16)


def pseudo_label_scenes(S, P=256, pool=512, C=18, seed=0, device="cpu"):
    """C5: per scene P axis-aligned proposals [x1,y1,z1,x2,y2,z2,score,label] and a
    pool of `pool` AABBs (half of them jittered copies of proposals so that the
    IoU>=0.3 match of lift_boxes.py:151-158 fires).  `device`: where the generator runs (the stream of a CUDA
    generator differs from the CPU one; a (seed, device type) pair is reproducible)."""
    g = torch.Generator(device=device).manual_seed(seed)
    dev = g.device
    c, s, _ = sample_boxes(g, (S, P), "scannet", 0.0)
    score = torch.rand((S, P), generator=g, device=dev)
    label = torch.randint(0, C, (S, P), generator=g, device=dev).float()
    h = P // 2  # second half: near-duplicates of the first half (same label) so that NMS@0.7 suppresses
    c[:, h:2 * h] = c[:, :h] + torch.randn((S, h, 3), generator=g, device=dev) * 0.03
    s[:, h:2 * h] = s[:, :h] * (torch.randn((S, h, 3), generator=g, device=dev) * 0.03 + 1.0).clamp(0.8, 1.2)
    label[:, h:2 * h] = label[:, :h]
    boxes = torch.cat([c - s / 2, c + s / 2, score[..., None], label[..., None]], -1)
    pc, ps, _ = sample_boxes(g, (S, pool), "scannet", 0.0)
    n = min(pool // 2, P)
    pc[:, :n] = c[:, :n] + torch.randn((S, n, 3), generator=g, device=dev) * 0.08
    ps[:, :n] = s[:, :n] * (torch.randn((S, n, 3), generator=g, device=dev) * 0.08 + 1.0).clamp(0.6, 1.4)
    poolb = torch.cat([pc - ps / 2, pc + ps / 2], -1)
    return boxes.double().contiguous(), poolb.double().contiguous()


def scene_points(corners_cam, n_points=4000, seed=0, frac_inside=0.5):
    """A synthetic depth-frame point cloud [B,N,3] for remove_empty_box: `frac_inside` of the points are
    sampled inside randomly chosen predicted boxes (so some boxes hold >= 5 points, many hold none),
    the rest uniformly in the room."""
    g = torch.Generator().manual_seed(seed)
    B, K = corners_cam.shape[0], corners_cam.shape[1]
    depth = torch.stack([corners_cam[..., 0], corners_cam[..., 2], -corners_cam[..., 1]], -1)   # flip_axis_to_depth
    n_in = int(n_points * frac_inside)
    pick = torch.randint(0, max(K // 3, 1), (B, n_in), generator=g)          # only the first third of the boxes get points
    o = torch.gather(depth[:, :, 0], 1, pick[..., None].expand(B, n_in, 3))
    ea = torch.gather(depth[:, :, 1] - depth[:, :, 0], 1, pick[..., None].expand(B, n_in, 3))
    eb = torch.gather(depth[:, :, 3] - depth[:, :, 0], 1, pick[..., None].expand(B, n_in, 3))
    ec = torch.gather(depth[:, :, 4] - depth[:, :, 0], 1, pick[..., None].expand(B, n_in, 3))
    t = torch.rand((B, n_in, 3), generator=g) * 0.9 + 0.05
    inside = o + t[..., 0:1] * ea + t[..., 1:2] * eb + t[..., 2:3] * ec
    lo = depth.reshape(B, -1, 3).min(1).values[:, None]
    hi = depth.reshape(B, -1, 3).max(1).values[:, None]
    noise = lo + torch.rand((B, n_points - n_in, 3), generator=g) * (hi - lo)
    return torch.cat([inside, noise], 1).float().contiguous()
