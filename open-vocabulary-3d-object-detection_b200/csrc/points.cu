// points.cu -- point-in-box passes of the evaluation / pseudo-label paths.
//
//   ovdet_points_in_boxes_count   remove_empty_box of parse_predictions (utils/ap_calculator.py:70-84):
//       number of scene points inside each predicted box.  The reference builds a Delaunay
//       triangulation of the 8 corners per box (utils/box_util.py:22-31) and calls find_simplex on
//       20-40k points; a box from box_parametrization_to_corners is a rectangular cuboid, so the hull
//       test is three slab tests in the box frame (points exactly on a face: inclusive, as Qhull's
//       find_simplex >= 0 with its tolerance).
//   ovdet_box_label_mode          LabelFormatter.gen_pseudo (utils/label_formatter.py:150-159): for each
//       (centre,size) box the mode of the labels of the points inside its axis-aligned extent
//       (crop_pc :183-188, inclusive bounds), ignoring IGNORE_LABEL; scipy.stats.mode returns the
//       smallest label among ties.
// Both stream the point cloud once per CTA (HBM/L2-bound), boxes of the scene staged in shared memory.
#include "common.cuh"

namespace ovdet {

constexpr int PT_NT = 256;
constexpr int PT_BOXES = 32;   // boxes per CTA

struct ObbFrame { float o[3], a[3], b[3], c[3], la, lb, lc; };

__global__ void __launch_bounds__(PT_NT) points_in_boxes_kernel(const float *__restrict__ pc, int N, int pstride,
                                                                const float *__restrict__ corners, int K, int32_t *counts)
{
    __shared__ ObbFrame fr[PT_BOXES];
    __shared__ int cnt[PT_BOXES];
    const int s = blockIdx.y, k0 = blockIdx.x * PT_BOXES;
    const int nb = min(PT_BOXES, K - k0);
    if (threadIdx.x < nb) {
        // corners are in the upright-camera frame; the point cloud is in the depth frame:
        // flip_axis_to_depth (ap_calculator.py:22-26): depth (x,y,z) = cam (x, z, -y)
        const float *c = corners + ((size_t)s * K + k0 + threadIdx.x) * 24;
        float p0[3], p1[3], p3[3], p4[3];
        auto to_depth = [&](int i, float *o) { o[0] = c[3 * i]; o[1] = c[3 * i + 2]; o[2] = -c[3 * i + 1]; };
        to_depth(0, p0); to_depth(1, p1); to_depth(3, p3); to_depth(4, p4);
        ObbFrame f;
        f.la = f.lb = f.lc = 0.f;
        for (int a = 0; a < 3; ++a) {
            f.o[a] = p0[a];
            f.a[a] = p1[a] - p0[a]; f.b[a] = p3[a] - p0[a]; f.c[a] = p4[a] - p0[a];
            f.la += f.a[a] * f.a[a]; f.lb += f.b[a] * f.b[a]; f.lc += f.c[a] * f.c[a];
        }
        fr[threadIdx.x] = f;
        cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    const float *pts = pc + (size_t)s * N * pstride;
    for (int i = threadIdx.x; i < N; i += PT_NT) {
        const float x = __ldg(pts + (size_t)i * pstride), y = __ldg(pts + (size_t)i * pstride + 1), z = __ldg(pts + (size_t)i * pstride + 2);
        for (int k = 0; k < nb; ++k) {
            const ObbFrame &f = fr[k];
            const float dx = x - f.o[0], dy = y - f.o[1], dz = z - f.o[2];
            const float ta = dx * f.a[0] + dy * f.a[1] + dz * f.a[2];
            const float tb = dx * f.b[0] + dy * f.b[1] + dz * f.b[2];
            const float tc = dx * f.c[0] + dy * f.c[1] + dz * f.c[2];
            if (ta >= 0.f && ta <= f.la && tb >= 0.f && tb <= f.lb && tc >= 0.f && tc <= f.lc) atomicAdd(&cnt[k], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < nb) counts[(size_t)s * K + k0 + threadIdx.x] = cnt[threadIdx.x];
}

constexpr int LM_MAXL = 64;   // label values 0..63

__global__ void __launch_bounds__(PT_NT) box_label_mode_kernel(const double *__restrict__ pts, const double *__restrict__ labels, int N,
                                                               const double *__restrict__ boxes, int bstride, int M, double ignore_label,
                                                               int32_t *mode_out, int32_t *count_out)
{
    __shared__ double lo[PT_BOXES][3], hi[PT_BOXES][3];
    __shared__ int hist[PT_BOXES][LM_MAXL];
    const int k0 = blockIdx.x * PT_BOXES;
    const int nb = min(PT_BOXES, M - k0);
    for (int i = threadIdx.x; i < PT_BOXES * LM_MAXL; i += PT_NT) (&hist[0][0])[i] = 0;
    if (threadIdx.x < nb) {
        const double *b = boxes + (size_t)(k0 + threadIdx.x) * bstride;   // centre(3), size(3), ...
        for (int a = 0; a < 3; ++a) {
            lo[threadIdx.x][a] = __dsub_rn(b[a], __ddiv_rn(b[3 + a], 2.0));   // crop_pc: box[0:3] -/+ box[3:6] / 2
            hi[threadIdx.x][a] = __dadd_rn(b[a], __ddiv_rn(b[3 + a], 2.0));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += PT_NT) {
        const double l = labels[i];
        if (l == ignore_label) continue;
        const int li = (int)l;
        if (li < 0 || li >= LM_MAXL || (double)li != l) continue;
        const double x = pts[(size_t)i * 3], y = pts[(size_t)i * 3 + 1], z = pts[(size_t)i * 3 + 2];
        for (int k = 0; k < nb; ++k)
            if (x >= lo[k][0] && x <= hi[k][0] && y >= lo[k][1] && y <= hi[k][1] && z >= lo[k][2] && z <= hi[k][2])
                atomicAdd(&hist[k][li], 1);
    }
    __syncthreads();
    if (threadIdx.x < nb) {
        int best = -1, bc = 0, tot = 0;
        for (int l = 0; l < LM_MAXL; ++l) { const int h = hist[threadIdx.x][l]; tot += h; if (h > bc) { bc = h; best = l; } }
        mode_out[k0 + threadIdx.x] = best;    // -1 when no labelled point falls inside
        count_out[k0 + threadIdx.x] = tot;
    }
}

}  // namespace ovdet

using namespace ovdet;

extern "C" int ovdet_points_in_boxes_count(const float *point_cloud, int S, int N, int point_stride,
                                           const float *corners, int K, int32_t *counts, void *stream)
{
    OVDET_REQUIRE(S >= 0 && N >= 0 && K >= 0 && point_stride >= 3, "bad size");
    if (S == 0 || K == 0) return OVDET_OK;
    OVDET_REQUIRE(corners && counts && (point_cloud || N == 0), "null pointer");
    points_in_boxes_kernel<<<dim3((K + PT_BOXES - 1) / PT_BOXES, S), PT_NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        point_cloud, N, point_stride, corners, K, counts);
    return launch_ok("points_in_boxes_kernel");
}

extern "C" int ovdet_box_label_mode(const double *points, const double *labels, int N, const double *boxes, int box_stride, int M,
                                    double ignore_label, int32_t *mode_out, int32_t *count_out, void *stream)
{
    OVDET_REQUIRE(N >= 0 && M >= 0 && box_stride >= 6, "bad size");
    if (M == 0) return OVDET_OK;
    OVDET_REQUIRE(boxes && mode_out && count_out && ((points && labels) || N == 0), "null pointer");
    box_label_mode_kernel<<<(M + PT_BOXES - 1) / PT_BOXES, PT_NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        points, labels, N, boxes, box_stride, M, ignore_label, mode_out, count_out);
    return launch_ok("box_label_mode_kernel");
}
