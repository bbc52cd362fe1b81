// giou_hull.cu -- placeholder until the convex-hull enclosing volume lands.
#include "common.cuh"
namespace ovdet {
int giou3d_hull_impl(const float *, const float *, const int64_t *, int, int, int, int, unsigned, float *, void *)
{
    set_error("OVDET_GIOU_ENCL_HULL not implemented yet");
    return OVDET_ERR_UNSUPPORTED;
}
}  // namespace ovdet
