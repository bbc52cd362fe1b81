// giou3d_bwd.cu -- backward of the GIoU w.r.t. the query corners.
//
// The reference trains loss_giou (criterion.py:274-296) through autograd of generalized_box3d_iou_tensor
// (utils/box_util.py:517-618), which forces the 3e3 pairs/s TorchScript path whenever needs_grad=True.  The upstream
// gradient is non-zero only at the matched (query, GT) pairs, so the backward is sparse: one thread per pair, pairs
// with dL/dgiou == 0 or beyond nums_k2 exit at once.  An active pair differentiates exactly what the forward computes:
//   height overlap, box volume (edge lengths with the 1e-6 / 1e-8 clamps), AABB enclosing volume (gradient to the
//   arg-min / arg-max corner, first index on ties like torch.min/max(dim)), and the BEV intersection area --
//   axis-aligned: the product of the two clamped extents; rotated: FORWARD-MODE differentiation of the
//   Sutherland-Hodgman clip with 8 tangents (x,z of the 4 BEV vertices of the query box) carried through every
//   intersection point and the shoelace sum.
// Gradients w.r.t. corners2 (ground truth) are not produced.  fp32; atomicAdd into grad_corners1 [B,K1,8,3].
#include <math.h>

#include "common.cuh"

namespace ovdet {

struct DVert { float x, z, dx[8], dz[8]; };   // BEV vertex with tangents w.r.t. (r1[0].x, r1[0].z, r1[1].x, ... r1[3].z)

__device__ __forceinline__ bool bw_inside(float c1x, float c1z, float c2x, float c2z, float px, float pz)
{
    return (c2x - c1x) * (pz - c1z) > (c2z - c1z) * (px - c1x);
}

__device__ void bw_isect(float c1x, float c1z, float c2x, float c2z, const DVert &s, const DVert &e, DVert &o)
{
    const float dcx = c1x - c2x, dcz = c1z - c2z;
    const float dpx = s.x - e.x, dpz = s.z - e.z;
    const float n1 = c1x * c2z - c1z * c2x;
    const float n2 = s.x * e.z - s.z * e.x;
    const float den = dcx * dpz - dcz * dpx;
    const float n3 = 1.0f / den;
    const float ax = n1 * dpx - n2 * dcx, az = n1 * dpz - n2 * dcz;
    o.x = ax * n3; o.z = az * n3;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float ddpx = s.dx[k] - e.dx[k], ddpz = s.dz[k] - e.dz[k];
        const float dn2 = s.dx[k] * e.z + s.x * e.dz[k] - s.dz[k] * e.x - s.z * e.dx[k];
        const float dden = dcx * ddpz - dcz * ddpx;
        const float dn3 = -dden * n3 * n3;
        o.dx[k] = (n1 * ddpx - dn2 * dcx) * n3 + ax * dn3;
        o.dz[k] = (n1 * ddpz - dn2 * dcz) * n3 + az * dn3;
    }
}

// area of clip(r1, r2) and its gradient w.r.t. the 8 coordinates of r1
__device__ float bw_clip_area(const float *r1, const float *r2, float *darea)
{
    DVert A[SH_MAXV + 1], Bv[SH_MAXV + 1];
    DVert *cur = A, *nxt = Bv;
    int n = 4;
    for (int i = 0; i < 4; ++i) {
        cur[i].x = r1[2 * i]; cur[i].z = r1[2 * i + 1];
        for (int k = 0; k < 8; ++k) { cur[i].dx[k] = (k == 2 * i) ? 1.f : 0.f; cur[i].dz[k] = (k == 2 * i + 1) ? 1.f : 0.f; }
    }
    float c1x = r2[6], c1z = r2[7];
    for (int ci = 0; ci < 4 && n > 0; ++ci) {
        const float c2x = r2[2 * ci], c2z = r2[2 * ci + 1];
        int m = 0;
        DVert s = cur[n - 1];
        bool s_in = bw_inside(c1x, c1z, c2x, c2z, s.x, s.z);
        for (int i = 0; i < n; ++i) {
            const DVert e = cur[i];
            const bool e_in = bw_inside(c1x, c1z, c2x, c2z, e.x, e.z);
            if (e_in != s_in && m < SH_MAXV) bw_isect(c1x, c1z, c2x, c2z, s, e, nxt[m++]);
            if (e_in && m < SH_MAXV) nxt[m++] = e;
            s = e; s_in = e_in;
        }
        c1x = c2x; c1z = c2z;
        DVert *t = cur; cur = nxt; nxt = t;
        n = m;
    }
    for (int k = 0; k < 8; ++k) darea[k] = 0.f;
    if (n <= 0) return 0.f;
    float S = 0.f, dS[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const DVert &v = cur[i], &pv = cur[(i + n - 1) % n];
        S += v.x * pv.z - v.z * pv.x;
        for (int k = 0; k < 8; ++k) dS[k] += v.dx[k] * pv.z + v.x * pv.dz[k] - v.dz[k] * pv.x - v.z * pv.dx[k];
    }
    const float sg = S > 0.f ? 1.f : (S < 0.f ? -1.f : 0.f);
    for (int k = 0; k < 8; ++k) darea[k] = 0.5f * sg * dS[k];
    return 0.5f * fabsf(S);
}

struct BwdParams { const float *c1, *c2; const int64_t *nums_k2; const float *gout; float *gc1; int B, K1, K2; unsigned flags; long long total; };

__global__ void __launch_bounds__(128) giou3d_backward_kernel(BwdParams p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.total) return;
    const float go = p.gout[idx];
    if (go == 0.f) return;
    const int k2 = (int)(idx % p.K2);
    const long long bq = idx / p.K2;
    const int b = (int)(bq / p.K1);
    if (p.nums_k2 && k2 >= p.nums_k2[b]) return;   // gious *= mask (box_util.py:611-616)
    float c1[24], c2[24], g[24];
    const float *a = p.c1 + bq * 24, *bb = p.c2 + ((long long)b * p.K2 + k2) * 24;
    for (int i = 0; i < 24; ++i) { c1[i] = a[i]; c2[i] = bb[i]; g[i] = 0.f; }
    const bool rotated = p.flags & OVDET_GIOU_ROTATED, prefilter = p.flags & OVDET_GIOU_PREFILTER;
    // ---- forward pieces
    const float ymax = fminf(c1[1], c2[1]), ymin = fmaxf(c1[13], c2[13]);
    const float hraw = ymax - ymin, h = fmaxf(hraw, 0.f);
    float r1[8], r2[8];
    for (int i = 0; i < 4; ++i) { r1[2 * i] = c1[3 * (3 - i)]; r1[2 * i + 1] = c1[3 * (3 - i) + 2]; r2[2 * i] = c2[3 * (3 - i)]; r2[2 * i + 1] = c2[3 * (3 - i) + 2]; }
    const float w0r = fminf(r1[6], r2[6]) - fmaxf(r1[2], r2[2]), w1r = fminf(r1[7], r2[7]) - fmaxf(r1[3], r2[3]);
    const float w0 = fmaxf(w0r, 0.f), w1 = fmaxf(w1r, 0.f);
    const float nonrot = w0 * w1;
    // volumes
    float ev[3], dsq[3], dvec[3][3];
    const int pa[3] = {0, 1, 0}, pb[3] = {1, 2, 4};
    for (int t = 0; t < 3; ++t) {
        float s = 0.f;
        for (int ax = 0; ax < 3; ++ax) { dvec[t][ax] = c1[3 * pa[t] + ax] - c1[3 * pb[t] + ax]; s += dvec[t][ax] * dvec[t][ax]; }
        dsq[t] = s;
        ev[t] = sqrtf(fmaxf(s, 1e-6f));
    }
    const float v1raw = ev[0] * ev[1] * ev[2], v1 = fmaxf(v1raw, 1e-8f);
    float e2[3];
    for (int t = 0; t < 3; ++t) {
        float s = 0.f;
        for (int ax = 0; ax < 3; ++ax) { const float d = c2[3 * pa[t] + ax] - c2[3 * pb[t] + ax]; s += d * d; }
        e2[t] = sqrtf(fmaxf(s, 1e-6f));
    }
    const float v2 = fmaxf(e2[0] * e2[1] * e2[2], 1e-8f);
    const float sumv = v1 + v2;
    // enclosing AABB
    float ext[3]; int imin[3], imax[3]; bool min1[3], max1[3]; float sgn[3];
    for (int ax = 0; ax < 3; ++ax) {
        float mn1 = c1[ax], mx1 = c1[ax], mn2 = c2[ax], mx2 = c2[ax];
        imin[ax] = 0; imax[ax] = 0;
        for (int i = 1; i < 8; ++i) {
            const float v = c1[3 * i + ax];
            if (v < mn1) { mn1 = v; imin[ax] = i; }
            if (v > mx1) { mx1 = v; imax[ax] = i; }
            mn2 = fminf(mn2, c2[3 * i + ax]); mx2 = fmaxf(mx2, c2[3 * i + ax]);
        }
        min1[ax] = mn1 < mn2; max1[ax] = mx1 > mx2;
        const float d = fmaxf(mx1, mx2) - fminf(mn1, mn2);
        sgn[ax] = d >= 0.f ? 1.f : -1.f;
        ext[ax] = fabsf(d);
    }
    const float encl = ext[0] * ext[1] * ext[2];
    if (!(encl > 2e-8f && sumv > 4e-8f)) return;   // gious *= good_boxes
    // intersection area (+ tangents)
    float area = 0.f, darea[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (rotated) {
        if (!(prefilter && nonrot == 0.f)) area = bw_clip_area(r1, r2, darea);
    } else {
        area = nonrot;
        // r1[3] = corner 0, r1[1] = corner 2 (box_util.py:557-560); tangent slots: r1[i].x -> 2i, r1[i].z -> 2i+1
        if (w0r > 0.f) { if (r1[6] < r2[6]) darea[6] += w1; if (r1[2] > r2[2]) darea[2] -= w1; }
        if (w1r > 0.f) { if (r1[7] < r2[7]) darea[7] += w0; if (r1[3] > r2[3]) darea[3] -= w0; }
    }
    const float inter = area * h;
    const float uraw = sumv - inter;
    const float uni = fmaxf(uraw, 1e-8f), uc = uraw > 1e-8f ? 1.f : 0.f;
    const float q = inter / (uni * uni) - 1.f / encl;        // -(d giou / d uni)
    const float k_inter = go * (1.f / uni + uc * q);
    const float k_v1 = go * uc * (-q) * (v1raw > 1e-8f ? 1.f : 0.f);
    const float k_encl = go * (-uni / (encl * encl));
    // ---- scatter
    // area -> BEV x,z of corners 0..3 (rect vertex i = corner 3-i)
    const float ka = k_inter * h;
    for (int i = 0; i < 4; ++i) { g[3 * (3 - i)] += ka * darea[2 * i]; g[3 * (3 - i) + 2] += ka * darea[2 * i + 1]; }
    // height -> y of corners 0 and 4
    if (hraw > 0.f) {
        const float kh = k_inter * area;
        if (c1[1] < c2[1]) g[1] += kh; else if (c1[1] == c2[1]) g[1] += 0.5f * kh;
        if (c1[13] > c2[13]) g[13] -= kh; else if (c1[13] == c2[13]) g[13] -= 0.5f * kh;
    }
    // volume -> corners 0,1,2,4
    for (int t = 0; t < 3; ++t) {
        if (dsq[t] > 1e-6f) {
            const float others = v1raw / ev[t];
            for (int ax = 0; ax < 3; ++ax) {
                const float d = k_v1 * others * dvec[t][ax] / ev[t];
                g[3 * pa[t] + ax] += d; g[3 * pb[t] + ax] -= d;
            }
        }
    }
    // enclosing volume -> arg-max / arg-min corners
    for (int ax = 0; ax < 3; ++ax) {
        const float k = k_encl * (encl / fmaxf(ext[ax], 1e-30f)) * sgn[ax];
        if (max1[ax]) g[3 * imax[ax] + ax] += k;
        if (min1[ax]) g[3 * imin[ax] + ax] -= k;
    }
    float *out = p.gc1 + bq * 24;
    for (int i = 0; i < 24; ++i) if (g[i] != 0.f) atomicAdd(out + i, g[i]);
}

}  // namespace ovdet

using namespace ovdet;

extern "C" int ovdet_giou3d_backward_f32(const float *corners1, const float *corners2, const int64_t *nums_k2, const float *grad_out,
                                         int B, int K1, int K2, unsigned flags, float *grad_corners1, void *stream)
{
    OVDET_REQUIRE(B >= 0 && K1 >= 0 && K2 >= 0, "negative size");
    if (B == 0 || K1 == 0) return OVDET_OK;
    OVDET_REQUIRE(grad_corners1, "null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    OVDET_CUDA_TRY(cudaMemsetAsync(grad_corners1, 0, sizeof(float) * (size_t)B * K1 * 24, st));
    if (K2 == 0) return OVDET_OK;
    OVDET_REQUIRE(corners1 && corners2 && grad_out, "null pointer");
    OVDET_REQUIRE(!(flags & (OVDET_GIOU_INTER_ONLY | OVDET_GIOU_ENCL_HULL | OVDET_GIOU_CLIP_F64)), "backward exists for the fp32 torch-path GIoU only");
    BwdParams p{corners1, corners2, nums_k2, grad_out, grad_corners1, B, K1, K2, flags, (long long)B * K1 * K2};
    giou3d_backward_kernel<<<(unsigned)((p.total + 127) / 128), 128, 0, st>>>(p);
    return launch_ok("giou3d_backward_kernel");
}
