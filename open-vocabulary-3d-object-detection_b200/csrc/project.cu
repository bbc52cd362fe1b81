// project.cu -- 3D box -> 2D image box for the RegionCLIP crop branch (second half).
//
// Replaces project_box_3d_cuda (utils/image_util.py:117-134) + the clip to the image (criterion.py:387-391):
//   corners (upright depth) = rotz(-heading) * (+-l, +-w, +-h) + centre      (size used as half extents, as the reference does)
//   upright depth -> depth  : Rtilt^T p                                      (image_util.py:284-290)
//   depth -> camera         : (x, y, z) -> (x, -z, y)                        (image_util.py:238-245)
//   camera -> image         : uv = K p, u /= w, v /= w                       (:292-298)
//   box = [min v, min u, max v, max u] over the 8 corners -- the reference unpacks `y1, x1 = min(corners_2d)` with
//   corners_2d = (u, v), so its "x" is the image row coordinate; kept as is (:130-133).
// One thread per box, the scene's Rtilt / K (18 floats) read through the read-only path; ~15 small torch kernels and
// an [..,3,8] intermediate per decoder layer become one launch.
#include "common.cuh"

namespace ovdet {

__global__ void __launch_bounds__(256) project_box3d_kernel(const float *__restrict__ center, const float *__restrict__ size,
                                                            const float *__restrict__ angle, const float *__restrict__ rtilt,
                                                            const float *__restrict__ kmat, const float *__restrict__ clip_wh,
                                                            int B, int Q, float *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * Q) return;
    const int b = (int)(i / Q);
    const float *R = rtilt + (size_t)b * 9, *K = kmat + (size_t)b * 9;
    const float cx = center[3 * i], cy = center[3 * i + 1], cz = center[3 * i + 2];
    const float l = size[3 * i], w = size[3 * i + 1], h = size[3 * i + 2];
    const float t = -angle[i];
    const float c = cosf(t), s = sinf(t);
    const float sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {1, 1, -1, -1, 1, 1, -1, -1}, sz[8] = {1, 1, 1, 1, -1, -1, -1, -1};
    float umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x = sx[k] * l, y = sy[k] * w, z = sz[k] * h;
        // rotz(t) = [[c,-s,0],[s,c,0],[0,0,1]]
        const float px = (c * x - s * y) + cx, py = (s * x + c * y) + cy, pz = z + cz;
        // Rtilt^T p
        const float dx = R[0] * px + R[3] * py + R[6] * pz;
        const float dy = R[1] * px + R[4] * py + R[7] * pz;
        const float dz = R[2] * px + R[5] * py + R[8] * pz;
        const float qx = dx, qy = -dz, qz = dy;   // depth -> camera
        const float u0 = K[0] * qx + K[1] * qy + K[2] * qz;
        const float v0 = K[3] * qx + K[4] * qy + K[5] * qz;
        const float w0 = K[6] * qx + K[7] * qy + K[8] * qz;
        const float u = u0 / w0, v = v0 / w0;
        umin = fminf(umin, u); umax = fmaxf(umax, u);
        vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
    }
    float bx[4] = {vmin, umin, vmax, umax};
    if (clip_wh) {   // criterion.py:387-391: clamp_min(0), then minimum with (w, h, w, h)
        const float W = clip_wh[2 * b], H = clip_wh[2 * b + 1];
        bx[0] = fminf(fmaxf(bx[0], 0.f), W); bx[1] = fminf(fmaxf(bx[1], 0.f), H);
        bx[2] = fminf(fmaxf(bx[2], 0.f), W); bx[3] = fminf(fmaxf(bx[3], 0.f), H);
    }
    *reinterpret_cast<float4 *>(out + 4 * i) = make_float4(bx[0], bx[1], bx[2], bx[3]);
}

}  // namespace ovdet

using namespace ovdet;

extern "C" int ovdet_project_box3d_f32(const float *center, const float *size, const float *angle, const float *rtilt,
                                       const float *kmat, const float *clip_wh, int B, int Q, float *boxes2d, void *stream)
{
    OVDET_REQUIRE(B >= 0 && Q >= 0, "negative size");
    if (B == 0 || Q == 0) return OVDET_OK;
    OVDET_REQUIRE(center && size && angle && rtilt && kmat && boxes2d, "null pointer");
    OVDET_REQUIRE((reinterpret_cast<uintptr_t>(boxes2d) & 15) == 0, "boxes2d must be 16-byte aligned");
    const long long n = (long long)B * Q;
    project_box3d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        center, size, angle, rtilt, kmat, clip_wh, B, Q, boxes2d);
    return launch_ok("project_box3d_kernel");
}
