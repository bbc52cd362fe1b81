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

constexpr int PT_PER_THREAD = 4;                       // points held in registers per thread
constexpr int PT_CHUNK = PT_NT * PT_PER_THREAD * 2;    // points per CTA (two register rounds)

struct ObbFrame { float o[3], a[3], b[3], c[3], la, lb, lc, pad; };   // 16 floats: four LDS.128 per box

// grid (box chunks, point chunks, scenes): each CTA tests PT_CHUNK points against PT_BOXES boxes; per box the warp's
// hits are counted with one ballot per point register, one shared atomic per warp, one global atomic per CTA.
__global__ void __launch_bounds__(PT_NT) points_in_boxes_kernel(const float *__restrict__ pc, int N, int pstride,
                                                                const float *__restrict__ corners, int K, int32_t *counts)
{
    __shared__ __align__(16) ObbFrame fr[PT_BOXES];
    __shared__ int cnt[PT_BOXES];
    const int s = blockIdx.z, k0 = blockIdx.x * PT_BOXES;
    const int nb = min(PT_BOXES, K - k0);
    if (threadIdx.x < nb) {
        // corners are in the upright-camera frame; the point cloud is in the depth frame:
        // flip_axis_to_depth (ap_calculator.py:22-26): depth (x,y,z) = cam (x, z, -y)
        const float *c = corners + ((size_t)s * K + k0 + threadIdx.x) * 24;
        float p0[3], p1[3], p3[3], p4[3];
        auto to_depth = [&](int i, float *o) { o[0] = c[3 * i]; o[1] = c[3 * i + 2]; o[2] = -c[3 * i + 1]; };
        to_depth(0, p0); to_depth(1, p1); to_depth(3, p3); to_depth(4, p4);
        ObbFrame f;
        f.la = f.lb = f.lc = 0.f; f.pad = 0.f;
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
    const int lane = threadIdx.x & 31;
    const int base = blockIdx.y * PT_CHUNK;
    for (int i0 = base; i0 < min(base + PT_CHUNK, N); i0 += PT_NT * PT_PER_THREAD) {
        float x[PT_PER_THREAD], y[PT_PER_THREAD], z[PT_PER_THREAD];
        bool ok[PT_PER_THREAD];
#pragma unroll
        for (int j = 0; j < PT_PER_THREAD; ++j) {
            const int i = i0 + j * PT_NT + threadIdx.x;
            ok[j] = i < N;
            const size_t o = (size_t)(ok[j] ? i : 0) * pstride;
            x[j] = __ldg(pts + o); y[j] = __ldg(pts + o + 1); z[j] = __ldg(pts + o + 2);
        }
        for (int k = 0; k < nb; ++k) {
            const float4 f0 = reinterpret_cast<const float4 *>(&fr[k])[0], f1 = reinterpret_cast<const float4 *>(&fr[k])[1],
                         f2 = reinterpret_cast<const float4 *>(&fr[k])[2], f3 = reinterpret_cast<const float4 *>(&fr[k])[3];
            // o = f0.xyz, a = (f0.w, f1.x, f1.y), b = (f1.z, f1.w, f2.x), c = (f2.y, f2.z, f2.w), la/lb/lc = f3.xyz
            int hits = 0;
#pragma unroll
            for (int j = 0; j < PT_PER_THREAD; ++j) {
                const float dx = x[j] - f0.x, dy = y[j] - f0.y, dz = z[j] - f0.z;
                const float ta = dx * f0.w + dy * f1.x + dz * f1.y;
                const float tb = dx * f1.z + dy * f1.w + dz * f2.x;
                const float tc = dx * f2.y + dy * f2.z + dz * f2.w;
                const bool in = ok[j] && ta >= 0.f && ta <= f3.x && tb >= 0.f && tb <= f3.y && tc >= 0.f && tc <= f3.z;
                hits += __popc(__ballot_sync(0xffffffffu, in));
            }
            if (hits && lane == 0) atomicAdd(&cnt[k], hits);
        }
    }
    __syncthreads();
    if (threadIdx.x < nb && cnt[threadIdx.x]) atomicAdd(counts + (size_t)s * K + k0 + threadIdx.x, cnt[threadIdx.x]);
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
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    OVDET_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)S * K, st));
    if (N == 0) return OVDET_OK;
    OVDET_REQUIRE(S <= 65535 && (N + PT_CHUNK - 1) / PT_CHUNK <= 65535, "grid too large");
    points_in_boxes_kernel<<<dim3((K + PT_BOXES - 1) / PT_BOXES, (N + PT_CHUNK - 1) / PT_CHUNK, S), PT_NT, 0, st>>>(
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
