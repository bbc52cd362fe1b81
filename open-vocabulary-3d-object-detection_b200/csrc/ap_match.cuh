// ap_match.cuh -- feature records and exact-IoU helpers shared by the AP matching kernels (eval.cu, ap_front.cu).
#pragma once
#include <math.h>

#include "common.cuh"

namespace ovdet {

// ---------------------------------------------------------------- AP matching
// One 128-thread CTA per scene (small CTAs: the per-scene work is a chain of short latency-bound phases, so the SM
// is kept busy by having ~6 scenes in flight, not by wide CTAs):
//   stage   kept detections (ordered compaction) and present GT -> 64-byte feature records in shared memory
//           (BEV quad, y extent, BEV bounding rectangle, fp64 volume)
//   step 1  every det x GT pair: the cheap exact rejects (no height overlap / disjoint BEV rectangles => IoU 0),
//           survivors on a warp-aggregated queue (processed in slabs of AM_QCAP pairs)
//   step 2  fp64 Sutherland-Hodgman clip of the survivors: 8 lanes per pair when few (short chain), else one per lane
//   records score of every (class, slot), tp = 0
//   pass 1  thread per det: for each GT that is the det's first-max IoU within its class and beats a threshold:
//           claim = atomicMax((score_bits << 32) | ~det) per (GT, threshold)
//   pass 2  same walk: TP iff this det holds the claim
// Candidate pairs and their IoUs live in shared memory; the caller's workspace is only touched in the dense fallback.
constexpr int AM_NT = 128;
constexpr int AM_CLIP = 16;        // lanes of the dense (serial) clip path
constexpr int AM_QCAP = 512;       // candidate pairs held in shared memory (pair index + IoU)

struct __align__(16) AmBox { float qx[4], qz[4]; float ytop, ybot, lox, hix, loz, hiz; double vol; };   // 64 B

__device__ __forceinline__ void am_features(const float *c, AmBox &f)
{
    using A = Ar<double>;
#pragma unroll
    for (int i = 0; i < 4; ++i) { f.qx[i] = c[3 * (3 - i)]; f.qz[i] = c[3 * (3 - i) + 2]; }   // rect order of box3d_iou (box_util.py:127-128)
    f.ytop = c[1]; f.ybot = c[13];
    f.lox = fminf(fminf(c[0], c[3]), fminf(c[6], c[9])); f.hix = fmaxf(fmaxf(c[0], c[3]), fmaxf(c[6], c[9]));
    f.loz = fminf(fminf(c[2], c[5]), fminf(c[8], c[11])); f.hiz = fmaxf(fmaxf(c[2], c[5]), fmaxf(c[8], c[11]));
    const int pa[3] = {0, 1, 0}, pb[3] = {1, 2, 4};
    double e[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const double dx = A::sub((double)c[3 * pa[t]], (double)c[3 * pb[t]]);
        const double dy = A::sub((double)c[3 * pa[t] + 1], (double)c[3 * pb[t] + 1]);
        const double dz = A::sub((double)c[3 * pa[t] + 2], (double)c[3 * pb[t] + 2]);
        e[t] = A::sqrt(A::add(A::add(A::mul(dx, dx), A::mul(dy, dy)), A::mul(dz, dz)));
    }
    f.vol = A::mul(A::mul(e[0], e[1]), e[2]);
}

__device__ __forceinline__ void am_load_box(const float *g, float *c)
{
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) { const float4 v = __ldg(reinterpret_cast<const float4 *>(g) + i); c[4 * i] = v.x; c[4 * i + 1] = v.y; c[4 * i + 2] = v.z; c[4 * i + 3] = v.w; }
    } else {
#pragma unroll
        for (int i = 0; i < 24; ++i) c[i] = __ldg(g + i);
    }
}

__device__ __forceinline__ double am_finish_iou(double ia, const AmBox &a, const AmBox &b)
{
    using A = Ar<double>;
    const double h = A::max(0.0, A::sub(A::min((double)a.ytop, (double)b.ytop), A::max((double)a.ybot, (double)b.ybot)));
    const double iv = A::mul(ia, h);
    return A::div(iv, A::sub(A::add(a.vol, b.vol), iv));
}

__device__ __forceinline__ double coop_area_f64(double vx, double vy, int n, int gl)
{   // area_f64 on a register-resident polygon: products in parallel, sums in the reference order
    const unsigned full = 0xffffffffu;
    const int src = max((gl == 0 ? n : gl) - 1, 0);
    const double px = __shfl_sync(full, vx, src, 8), py = __shfl_sync(full, vy, src, 8);
    const double t1 = __dmul_rn(vx, py), t2 = __dmul_rn(vy, px);
    double d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int i = 0; i < SH_MAXV; ++i) {
        const double u = __shfl_sync(full, t1, i, 8), w = __shfl_sync(full, t2, i, 8);
        if (i < n) { d1 = __dadd_rn(d1, u); d2 = __dadd_rn(d2, w); }
    }
    return n < 3 ? 0.0 : __dmul_rn(0.5, fabs(__dsub_rn(d1, d2)));
}

__device__ __forceinline__ uint32_t score_key(float s)
{   // ascending key order == descending score; -inf (absent) sorts last
    const uint32_t b = __float_as_uint(s);
    const uint32_t ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~ord;
}

}  // namespace ovdet
