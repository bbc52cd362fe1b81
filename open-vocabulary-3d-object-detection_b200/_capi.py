"""ctypes binding of lib/libovdet_b200.so (include/ovdet_b200.h).

There is no CPU fallback: if the library has not been built, or a compute entry
point is called without a CUDA device, an exception is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libovdet_b200.so")

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_u = ctypes.c_uint
c_f = ctypes.c_float
c_d = ctypes.c_double
c_i64 = ctypes.c_int64
c_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/ovdet_b200.h one to one
SIGNATURES = {
    "ovdet_version": (c_i, []),
    "ovdet_last_error": (ctypes.c_char_p, []),
    "ovdet_device_count": (c_i, []),
    "ovdet_stream_synchronize": (c_i, [c_p]),
    "ovdet_giou3d_f32": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_u, c_p, c_p]),
    "ovdet_box_corners_f32": (c_i, [c_p, c_p, c_p, c_i64, c_p, c_p]),
    "ovdet_giou3d_decode_f32": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_u, c_p, c_p, c_p]),
    "ovdet_project_box3d_f32": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p]),
    "ovdet_giou3d_backward_f32": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_u, c_p, c_p]),
    "ovdet_box_intersection_f32": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "ovdet_box_intersection_host_f32": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i]),
    "ovdet_giou3d_host_f32": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_u, c_p]),
    "ovdet_box3d_iou_f64": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p]),
    "ovdet_matcher_cost_f32": (c_i, [c_p] * 10 + [c_i] * 4 + [c_f] * 4 + [c_u, c_i, c_p, c_p, c_p]),
    "ovdet_matcher_step_f32": (c_i, [c_p] * 8 + [c_i] * 4 + [c_f] * 4 + [c_u, c_i] + [c_p] * 6),
    "ovdet_lsap_f32": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "ovdet_nms_f64": (c_i, [c_p, c_p, c_i, c_i, c_i, c_d, c_d, c_u, c_p, c_p, c_p, c_p]),
    "ovdet_parse_predictions_f32": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_d, c_f, c_u, c_p, c_p, c_p, c_p, c_p]),
    "ovdet_ap_match": (c_i, [c_p] * 8 + [c_i] * 4 + [c_p, c_i] + [c_p] * 5),
    "ovdet_aabb_iou_f64": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "ovdet_ap_match_iou": (c_i, [c_p] * 7 + [c_i] * 4 + [c_p, c_i] + [c_p] * 4),
    "ovdet_ap_reduce_ws_bytes": (c_sz, [c_i, c_i64]),
    "ovdet_ap_reduce": (c_i, [c_p, c_p, c_p, c_i, c_i64, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "ovdet_apc_collect": (c_i, [c_p, c_p, c_i, c_i64, c_i, c_p, c_p, c_p, c_p, c_p]),
    "ovdet_apc_sort": (c_i, [c_p, c_p, c_i, c_i, c_p]),
    "ovdet_apc_hist": (c_i, [c_p, c_i, c_i64, c_p, c_i, c_p, c_p]),
    "ovdet_apc_final": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "ovdet_ap_front_f32": (c_i, [c_p] * 7 + [c_i] * 4 + [c_d, c_f, c_u, c_p, c_i] + [c_p] * 7 + [c_i, c_p, c_p]),
    "ovdet_symm_alloc": (c_i, [c_sz, c_p]),
    "ovdet_symm_free": (c_i, [c_p]),
    "ovdet_symm_export": (c_i, [c_p, c_p]),
    "ovdet_symm_open": (c_i, [c_p, c_p]),
    "ovdet_symm_close": (c_i, [c_p]),
    "ovdet_apx_local_bytes": (c_sz, [c_i, c_i]),
    "ovdet_apx_symm_bytes": (c_sz, [c_i, c_i, c_i]),
    "ovdet_apx_reduce": (c_i, [c_p, c_p, c_i, c_i] + [c_p] * 4 + [c_i, c_i, c_i, c_u, c_i, c_i] + [c_p] * 5),
    "ovdet_points_in_boxes_count": (c_i, [c_p, c_i, c_i, c_i, c_p, c_i, c_p, c_p]),
    "ovdet_box_label_mode": (c_i, [c_p, c_p, c_i, c_p, c_i, c_i, c_d, c_p, c_p, c_p]),
    "ovdet_clip_logits_bf16": (c_i, [c_p, c_p, c_i, c_i, c_i, c_u, c_f, c_p, c_i, c_p, c_i, c_p, c_p]),
    "ovdet_pseudo_filter_f64": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_d, c_d, c_d, c_p, c_p, c_p, c_p, c_p]),
}

# flag words (include/ovdet_b200.h)
GIOU_ROTATED, GIOU_PREFILTER, GIOU_INTER_ONLY, GIOU_CLIP_F64, GIOU_ENCL_HULL = 1, 2, 4, 8, 16
NMS_2D, NMS_SAMECLS, NMS_OLD_TYPE, NMS_LHS, PARSE_NO_NMS = 1, 2, 4, 8, 0x100
LOGITS_L2NORM = 1
FRONT_PER_CLASS, FRONT_CLS_CONF, FRONT_GT_PRESENT_F32, FRONT_RESET = 0x1000, 0x2000, 0x4000, 0x8000
APX_FORCE_EXCHANGE, APX_USE_07_METRIC = 1, 2
APX_STAGE_PUSH, APX_STAGE_MERGE_HIST, APX_STAGE_FINAL = 0x10, 0x20, 0x40
SYMM_HANDLE_BYTES = 64

_LIB = None


class OvdetError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise OvdetError(
                "libovdet_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python open-vocabulary-3d-object-detection_b200/build.py`. There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        msg = lib().ovdet_last_error().decode("utf-8", "replace")
        raise OvdetError("libovdet_b200 error %d: %s" % (rc, msg))


def ptr(t):
    """Device (or host) address of a contiguous tensor / numpy array, None -> NULL."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous(), "ovdet_b200 kernels need contiguous tensors"
        return t.data_ptr()
    return t.ctypes.data


def stream(device=None):
    """Raw handle (int) of torch's current CUDA stream on `device` (a torch.device with an index, or None = current)."""
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    return torch._C._cuda_getCurrentRawStream(idx)


def stream_synchronize(device=None):
    """Block until the current stream of `device` has drained (one cudaStreamSynchronize through the C ABI)."""
    check(lib().ovdet_stream_synchronize(stream(device)))


def as_input(t, dtype, device):
    """`t` as a contiguous, detached tensor of `dtype` on `device` -- the tensor itself when it already is one (the
    usual case on the hot path: no torch op, no allocation)."""
    if t.dtype is dtype and t.device == device and t.is_contiguous() and not t.requires_grad:
        return t
    return t.detach().to(device=device, dtype=dtype).contiguous()


class on_device(object):
    """``with on_device(dev):`` = torch.cuda.device(dev), skipped when `dev` is already current (saves ~10 us per call)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index is None or torch.cuda.current_device() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            return self.ctx.__exit__(*a)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise OvdetError("expected a CUDA tensor (the product path has no CPU implementation)")
