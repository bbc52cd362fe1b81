"""Drop-in for the reference's Cython extension ``utils.box_intersection``
(utils/box_intersection.pyx:166-198): same name, same positional signature, same
in-place contract on numpy buffers -- the hot loop runs on the GPU through the
host-buffer C entry point ``ovdet_box_intersection_host_f32``."""
import numpy as np

from .. import _capi as C

K2_LOOP_AS_SHIPPED = 4  # `K2 = rect2.shape[2]` (pyx:180)


def box_intersection(rect1, rect2, non_rot_inter_areas, nums_k2, inter_areas, approximate, k2_loop=None):
    """rect1 float32[B,K1,4,2], rect2 float32[B,K2,4,2], non_rot_inter_areas
    float32[B,K1,K2], nums_k2 int32[B], inter_areas float32[B,K1,K2] (written in
    place), approximate bool.  ``k2_loop=None`` reproduces the shipped loop bound
    ``rect2.shape[2]`` (=4); pass ``rect2.shape[1]`` for all GT columns."""
    for a, dt in ((rect1, np.float32), (rect2, np.float32), (non_rot_inter_areas, np.float32),
                  (nums_k2, np.int32), (inter_areas, np.float32)):
        # typed-memoryview contract of the Cython signature (pyx:166-171)
        if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous):
            raise ValueError("Buffer dtype mismatch or non-contiguous buffer")
    B, K1, K2 = rect1.shape[0], rect1.shape[1], rect2.shape[1]
    if k2_loop is None:
        k2_loop = rect2.shape[2]
    C.check(C.lib().ovdet_box_intersection_host_f32(C.ptr(rect1), C.ptr(rect2), C.ptr(non_rot_inter_areas),
                                                    C.ptr(nums_k2), C.ptr(inter_areas), int(bool(approximate)),
                                                    B, K1, K2, int(k2_loop)))
