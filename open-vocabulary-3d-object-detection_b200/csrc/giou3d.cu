// giou3d.cu -- rotated / axis-aligned 3D GIoU for sm_100a.
//
// Replaces generalized_box3d_iou (utils/box_util.py:717-737; bodies :517-618 and
// :624-714) and the Cython hot loop box_intersection (utils/box_intersection.pyx:166-198).
//
// Layout.  One CTA owns a tile of TQ query boxes of one batch element and walks
// the GT boxes in chunks of TG.  Corners ([.,8,3] fp32 = 96 B per box) are staged
// through shared memory with coalesced float4 loads, reduced once per box to 17
// features (BEV quad, y-extent, volume, AABB) kept SoA in shared memory.
//   phase A  every thread takes pairs (row, col) of the tile: height overlap,
//            axis-aligned prefilter area, AABB enclosing volume, validity; pairs
//            that need no polygon clip are finished here; the others are pushed
//            (warp-aggregated) on a CTA-local queue.
//   phase B  the queue is drained.  Short queue (the shipped semantics): 8 lanes per
//            pair, the polygon in registers, one vertex per lane -- the dependent chain
//            of a clip pass is one vertex long.  Long queue (every pair clipped): one
//            Sutherland-Hodgman clip per lane, vertex lists in a conflict-free per-thread
//            shared scratch.  Both give bit-identical polygons; the shoelace is summed in
//            the reference's term order.
//   split    when at most 64 pairs of the tile are clip-eligible (k2_cap = 4 columns),
//            the last two warps find and clip them WHILE the other six run phase A.
// Results go to a shared [TQ][TG] tile and leave with coalesced row stores.
// The pair work is ~60 (skip) to ~1200 (clip) instructions against 4 B written,
// i.e. issue-bound, not HBM-bound: see DESIGN.md for the roofline used.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ovdet {

constexpr int TQ_MAX = 32;  // query rows per CTA (template parameter TQ: 32, 16 or 8)
constexpr int TG = 64;    // GT columns per chunk
constexpr int NT = 256;   // threads per CTA
constexpr int NF = 17;    // features per box
// feature indices
enum { F_RX = 0, F_RZ = 4, F_YTOP = 8, F_YBOT = 9, F_VOL = 10, F_MN = 11, F_MX = 14 };

// optional fused Hungarian-matcher cost epilogue (criterion.py:40-63)
struct MatcherEpi {
    const float *prob = nullptr;         // [B,K1,C]
    const float *obj = nullptr;          // [B,K1]
    const float *center_dist = nullptr;  // [B,K1,K2] or null -> L1 of cq/cg
    const float *cq = nullptr, *cg = nullptr;  // [B,K1,3], [B,K2,3]
    const int64_t *labels = nullptr;     // [B,K2]
    float *cost = nullptr;               // [B,K1,K2]
    int C = 0;
    float wc = 0, wo = 0, wce = 0, wg = 0;
};

// optional fused box decode: query boxes given as (centre in the depth frame, size l/w/h, heading)
// instead of 8 corners -- 28 B/box read instead of 96 B, and no separate decode kernels
struct BoxDecode { const float *center = nullptr, *size = nullptr, *angle = nullptr; float *corners_out = nullptr; };

// box_parametrization_to_corners (datasets/sunrgbd.py:145-148): flip_axis_to_camera (utils/box_util.py:288-295)
// then get_3d_box_batch_tensor (:313-352): corners = local @ roty(angle)^T + centre
__device__ __forceinline__ void decode_box(const float *ctr, const float *sz, float ang, float *c)
{
    const float cx = ctr[0], cy = -ctr[2], cz = ctr[1];
    const float l = sz[0] * 0.5f, w = sz[1] * 0.5f, h = sz[2] * 0.5f;
    const float co = cosf(ang), si = sinf(ang);
    const float sx[8] = {1, 1, -1, -1, 1, 1, -1, -1}, sy[8] = {1, 1, 1, 1, -1, -1, -1, -1}, szn[8] = {1, -1, -1, 1, 1, -1, -1, 1};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float x = sx[i] * l, y = sy[i] * h, z = szn[i] * w;
        c[3 * i] = __fadd_rn(__fadd_rn(__fmul_rn(x, co), __fmul_rn(z, si)), cx);
        c[3 * i + 1] = __fadd_rn(y, cy);
        c[3 * i + 2] = __fadd_rn(__fadd_rn(__fmul_rn(x, -si), __fmul_rn(z, co)), cz);
    }
}

struct GiouParams {
    const float *c1, *c2;
    const int64_t *nums_k2;
    float *out;   // may be null when only the matcher cost is wanted
    int B, K1, K2, k2_cap;
    unsigned flags;
    int tiles_per_b;
    MatcherEpi epi;
    BoxDecode dec1;
    unsigned long long *dbg;   // optional [grid][8] globaltimer stamps of thread 0 (OVDET_GIOU_DBG_PTR; null in production)
};

__device__ __forceinline__ unsigned long long gtimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define GSTAMP(i) do { if (p.dbg && threadIdx.x == 0) p.dbg[(size_t)blockIdx.x * 8 + (i)] = gtimer_ns(); } while (0)

// features of one box from its 24 staged floats (utils/box_util.py:550-555 rect,
// :544-546 y extent, :443-463 volume with the 1e-8 clamp of :568-569, AABB for :466-514)
__device__ __forceinline__ void box_features(const float *c, float *f, int stride)
{
    using A = Ar<float>;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[(F_RX + i) * stride] = c[3 * (3 - i)];
        f[(F_RZ + i) * stride] = c[3 * (3 - i) + 2];
    }
    f[F_YTOP * stride] = c[1];
    f[F_YBOT * stride] = c[13];
    float e[3];
    const int pa[3] = {0, 1, 0}, pb[3] = {1, 2, 4};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const float dx = A::sub(c[3 * pa[t]], c[3 * pb[t]]);
        const float dy = A::sub(c[3 * pa[t] + 1], c[3 * pb[t] + 1]);
        const float dz = A::sub(c[3 * pa[t] + 2], c[3 * pb[t] + 2]);
        const float s = A::add(A::add(A::mul(dx, dx), A::mul(dy, dy)), A::mul(dz, dz));
        e[t] = A::sqrt(fmaxf(s, 1e-6f));
    }
    f[F_VOL * stride] = fmaxf(A::mul(A::mul(e[0], e[1]), e[2]), 1e-8f);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float mn = c[a], mx = c[a];
#pragma unroll
        for (int i = 1; i < 8; ++i) { mn = fminf(mn, c[3 * i + a]); mx = fmaxf(mx, c[3 * i + a]); }
        f[(F_MN + a) * stride] = mn;
        f[(F_MX + a) * stride] = mx;
    }
}

// cooperative coalesced stage of `n` boxes (n*24 floats) into shared memory
__device__ __forceinline__ void stage_boxes(const float *__restrict__ g, int n, float *raw, bool vec)
{
    const int nfl = n * 24;
    if (vec) {
        for (int i = threadIdx.x; i < nfl / 4; i += NT) reinterpret_cast<float4 *>(raw)[i] = ldg4(g + 4 * i);
    } else {
        for (int i = threadIdx.x; i < nfl; i += NT) raw[i] = __ldg(g + i);
    }
}

struct PairTerms { float h, nonrot, encl, sumv; };

// the 13 per-box features the pair maths needs, as registers
struct BoxF { float ytop, ybot, x1, x3, z1, z3, vol, mn[3], mx[3]; };

__device__ __forceinline__ BoxF load_boxf(const float *f, int i, int stride)
{
    BoxF b;
    b.ytop = f[F_YTOP * stride + i]; b.ybot = f[F_YBOT * stride + i];
    b.x1 = f[(F_RX + 1) * stride + i]; b.x3 = f[(F_RX + 3) * stride + i];
    b.z1 = f[(F_RZ + 1) * stride + i]; b.z3 = f[(F_RZ + 3) * stride + i];
    b.vol = f[F_VOL * stride + i];
#pragma unroll
    for (int a = 0; a < 3; ++a) { b.mn[a] = f[(F_MN + a) * stride + i]; b.mx[a] = f[(F_MX + a) * stride + i]; }
    return b;
}

__device__ __forceinline__ PairTerms pair_terms(const BoxF &q, const BoxF &g)
{
    using A = Ar<float>;
    PairTerms t;
    t.h = fmaxf(A::sub(fminf(q.ytop, g.ytop), fmaxf(q.ybot, g.ybot)), 0.f);
    // prefilter: rect points 1 and 3 as lt / rb (box_util.py:557-560)
    const float w0 = fmaxf(A::sub(fminf(q.x3, g.x3), fmaxf(q.x1, g.x1)), 0.f);
    const float w1 = fmaxf(A::sub(fminf(q.z3, g.z3), fmaxf(q.z1, g.z1)), 0.f);
    t.nonrot = A::mul(w0, w1);
    float d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] = fabsf(A::sub(fmaxf(q.mx[a], g.mx[a]), fminf(q.mn[a], g.mn[a])));
    t.encl = A::mul(A::mul(d[0], d[1]), d[2]);
    t.sumv = A::add(q.vol, g.vol);
    return t;
}

// box_util.py:600-618 / :700-714
__device__ __forceinline__ float finish_pair(const PairTerms &t, float area, bool valid, bool has_nums, bool inter_only)
{
    using A = Ar<float>;
    const float inter = A::mul(area, t.h);
    if (inter_only) return inter;
    const float uni = fmaxf(A::sub(t.sumv, inter), 1e-8f);
    // inter == 0 (most pairs) -> 0 / uni == +0 exactly; skipping it also skips the IEEE divide's slow path, which a
    // zero numerator always takes (ncu: the CALL was executed by every warp)
    float iou = 0.f;
    if (inter != 0.f) iou = A::div(inter, uni);
    const float second = -A::sub(1.f, A::div(uni, t.encl));
    float g = A::add(iou, second);
    const float good = (t.encl > 2e-8f && t.sumv > 4e-8f) ? 1.f : 0.f;
    g = A::mul(g, good);
    if (has_nums) g = A::mul(g, valid ? 1.f : 0.f);
    return g;
}

// ---------------------------------------------------------------------------
// Convex-hull enclosing volume of two upright boxes (utils/box_ops3d.py:458-473: scipy ConvexHull of the 16
// corners), in closed form.  Both boxes are vertical prisms A x [a0,a1], B x [b0,b1] over BEV quads A, B.  The hull's
// horizontal cross-section at height y is the Minkowski combination  alpha*A + beta*B + gamma*H,  H = conv(A u B),
// with (alpha, beta, gamma) = (1-l2, l1, l2-l1) and [l1,l2] the feasible blend interval at y -- piecewise linear in
// y between the breakpoints {a0,a1,b0,b1}.  Its area is the quadratic form
//   alpha^2|A| + beta^2|B| + gamma^2|H| + 2ab V(A,B) + 2ag V(A,H) + 2bg V(B,H)
// (V = mixed area = 1/2 sum over CCW edges e of Q of max_{v in P} v x e), so Simpson's rule per piece is exact.
// fp64; ~700 flops, only evaluated for pairs whose intersection volume is > 0 (box_ops3d.py:467,:514-519).
// ---------------------------------------------------------------------------
struct P2d { double x, z; };

__device__ __forceinline__ double cross2(const P2d &o, const P2d &a, const P2d &b) { return (a.x - o.x) * (b.z - o.z) - (a.z - o.z) * (b.x - o.x); }

__device__ double poly_area_ccw(P2d *p, int n)
{   // signed shoelace; reverses p in place when clockwise so that callers always see CCW polygons
    double s = 0.0;
    for (int i = 0; i < n; ++i) { const P2d &a = p[i], &b = p[(i + 1) % n]; s += a.x * b.z - a.z * b.x; }
    if (s < 0.0) { for (int i = 0; i < n / 2; ++i) { const P2d t = p[i]; p[i] = p[n - 1 - i]; p[n - 1 - i] = t; } s = -s; }
    return 0.5 * s;
}

__device__ double mixed_area(const P2d *P, int np, const P2d *Q, int nq)
{
    double v = 0.0;
    for (int e = 0; e < nq; ++e) {
        const double dx = Q[(e + 1) % nq].x - Q[e].x, dz = Q[(e + 1) % nq].z - Q[e].z;
        double h = -1e300;
        for (int i = 0; i < np; ++i) h = fmax(h, P[i].x * dz - P[i].z * dx);
        v += h;
    }
    return 0.5 * v;
}

__device__ __noinline__ double hull_enclosing_volume(const float *ax, const float *az, const float *bx, const float *bz,
                                                     double ay0, double ay1, double by0, double by1)
{
    P2d A[4], Bq[4], pts[8], H[16];
    for (int i = 0; i < 4; ++i) { A[i].x = ax[i]; A[i].z = az[i]; Bq[i].x = bx[i]; Bq[i].z = bz[i]; pts[i] = A[i]; pts[4 + i] = Bq[i]; }
    const double a0 = fmin(ay0, ay1), a1 = fmax(ay0, ay1), b0 = fmin(by0, by1), b1 = fmax(by0, by1);
    // Andrew's monotone chain on the 8 BEV points
    for (int i = 1; i < 8; ++i) {
        const P2d t = pts[i]; int j = i - 1;
        while (j >= 0 && (pts[j].x > t.x || (pts[j].x == t.x && pts[j].z > t.z))) { pts[j + 1] = pts[j]; --j; }
        pts[j + 1] = t;
    }
    int k = 0;
    for (int i = 0; i < 8; ++i) { while (k >= 2 && cross2(H[k - 2], H[k - 1], pts[i]) <= 0.0) --k; H[k++] = pts[i]; }
    for (int i = 6, lo = k + 1; i >= 0; --i) { while (k >= lo && cross2(H[k - 2], H[k - 1], pts[i]) <= 0.0) --k; H[k++] = pts[i]; }
    const int nh = k - 1;
    if (nh < 3) return 0.0;
    const double aA = poly_area_ccw(A, 4), aB = poly_area_ccw(Bq, 4), aH = poly_area_ccw(H, nh);
    const double vAB = mixed_area(A, 4, Bq, 4), vAH = mixed_area(A, 4, H, nh), vBH = mixed_area(Bq, 4, H, nh);
    auto area_at = [&](double y) {
        double l1 = 0.0, l2 = 1.0;
        const double d0 = b0 - a0, d1 = b1 - a1;           // lo(l) = a0 + l*d0 <= y <= a1 + l*d1 = hi(l)
        if (d0 > 0.0) l2 = fmin(l2, (y - a0) / d0); else if (d0 < 0.0) l1 = fmax(l1, (y - a0) / d0);
        if (d1 > 0.0) l1 = fmax(l1, (y - a1) / d1); else if (d1 < 0.0) l2 = fmin(l2, (y - a1) / d1);
        l1 = fmin(fmax(l1, 0.0), 1.0); l2 = fmin(fmax(l2, l1), 1.0);
        const double al = 1.0 - l2, be = l1, ga = l2 - l1;
        return al * al * aA + be * be * aB + ga * ga * aH + 2.0 * (al * be * vAB + al * ga * vAH + be * ga * vBH);
    };
    double ys[4] = {a0, a1, b0, b1};
    for (int i = 1; i < 4; ++i) { const double t = ys[i]; int j = i - 1; while (j >= 0 && ys[j] > t) { ys[j + 1] = ys[j]; --j; } ys[j + 1] = t; }
    double vol = 0.0;
    for (int i = 0; i < 3; ++i) {
        const double y0 = ys[i], y1 = ys[i + 1];
        if (y1 > y0) vol += (y1 - y0) / 6.0 * (area_at(y0) + 4.0 * area_at(0.5 * (y0 + y1)) + area_at(y1));
    }
    return vol;
}

// ClipT = float: all-fp32 clip (torch path); double: Cython arithmetic.  PB = threads that drain the clip queue
// (their Sutherland-Hodgman scratch is the largest shared-memory consumer: 128 B (fp32) / 256 B (fp64) per thread).
template <typename ClipT, int PB, int TQ, bool HULL>
__global__ void __launch_bounds__(NT, TQ == 32 ? 2 : 4) giou3d_kernel(GiouParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // carve: f1[NF*TQ] f2[NF*TG] tile[TQ*TG] raw[(TQ+TG)*24] queue[TQ*TG u16] scratch[2*SH_MAXV*PB V2<ClipT>]
    float *f1 = reinterpret_cast<float *>(smem_raw);
    float *f2 = f1 + NF * TQ;
    float *tile = f2 + NF * TG;
    float *raw1 = tile + TQ * TG;
    float *raw2 = raw1 + TQ * 24;
    unsigned short *queue = reinterpret_cast<unsigned short *>(raw2 + TG * 24);
    V2<ClipT> *scratch = reinterpret_cast<V2<ClipT> *>(queue + TQ * TG);
    __shared__ int qcount;
    // Programmatic dependent launch: the grid may be scheduled while the previous kernel of the stream drains; nothing
    // in global memory is touched before this wait (full completion + visibility of the prerequisite grids), and the
    // next launch is allowed to start its own scheduling right away.  Plain stream order when the neighbours are not PDL.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int b = blockIdx.x / p.tiles_per_b;
    const int q0 = (blockIdx.x - b * p.tiles_per_b) * TQ;
    const int nq = min(TQ, p.K1 - q0);
    const bool rotated = p.flags & OVDET_GIOU_ROTATED;
    const bool prefilter = p.flags & OVDET_GIOU_PREFILTER;
    const bool inter_only = p.flags & OVDET_GIOU_INTER_ONLY;
    const bool hull = HULL && rotated;   // compile-time: the default instantiation carries no call / stack frame
    const bool has_nums = p.nums_k2 != nullptr;
    int nk = p.K2;
    if (has_nums) { const long long v = p.nums_k2[b]; nk = v < 0 ? 0 : (v > p.K2 ? p.K2 : (int)v); }
    const int clip_lim = p.k2_cap > 0 ? min(nk, p.k2_cap) : nk;
    const bool vec1 = ((reinterpret_cast<uintptr_t>(p.c1) & 15) == 0);
    const bool vec2 = ((reinterpret_cast<uintptr_t>(p.c2) & 15) == 0);
    const bool vec_out = p.out && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (p.K2 % 4 == 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GSTAMP(0);

    for (int g0 = 0; g0 < p.K2; g0 += TG) {
        const int ng = min(TG, p.K2 - g0);
        // ---- box features.  Corner tensors: thread t loads ITS box (six 16-byte loads, 96 contiguous bytes) straight into
        // registers and reduces it to the SoA features -- no staging buffer, one barrier.  Fused decode: the query box is
        // built from (centre, size, heading) in the staging buffer first (it is also written out when asked for).
        auto features_from_global = [&](const float *g, bool vec, float *f, int stride) {
            float c[24];
            if (vec) {
#pragma unroll
                for (int i = 0; i < 6; ++i) { const float4 v = ldg4(g + 4 * i); c[4 * i] = v.x; c[4 * i + 1] = v.y; c[4 * i + 2] = v.z; c[4 * i + 3] = v.w; }
            } else {
#pragma unroll
                for (int i = 0; i < 24; ++i) c[i] = __ldg(g + i);
            }
            box_features(c, f, stride);
        };
        if (threadIdx.x == 0) qcount = 0;
        if (threadIdx.x < TG) {
            if (threadIdx.x < ng) features_from_global(p.c2 + ((size_t)b * p.K2 + g0 + threadIdx.x) * 24, vec2, f2 + threadIdx.x, TG);
        } else if (g0 == 0 && threadIdx.x - TG < nq) {
            const int r = threadIdx.x - TG;
            if (p.dec1.center) {
                const size_t q = (size_t)b * p.K1 + q0 + r;
                decode_box(p.dec1.center + 3 * q, p.dec1.size + 3 * q, __ldg(p.dec1.angle + q), raw1 + r * 24);
                if (p.dec1.corners_out)
                    for (int i = 0; i < 24; ++i) p.dec1.corners_out[q * 24 + i] = raw1[r * 24 + i];
                box_features(raw1 + r * 24, f1 + r, TQ);
            } else features_from_global(p.c1 + ((size_t)b * p.K1 + q0 + r) * 24, vec1, f1 + r, TQ);
        }
        GSTAMP(1);
        __syncthreads();
        GSTAMP(2);

        // ---- phase A: warp w owns rows w*4 .. w*4+3, lane l owns columns l and l+32 (features in registers)
        BoxF gf[2];
        bool gvalid[2], gin[2], gclip[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            gin[h] = c < ng;
            gf[h] = load_boxf(f2, gin[h] ? c : 0, TG);
            gvalid[h] = (g0 + c) < nk;
            gclip[h] = (g0 + c) < clip_lim;
        }
        // one pair of the clip queue, 8 lanes per pair (shared by phase B and by the clip warps of the split mode)
        auto clip_pair_coop = [&](int r, int c, bool act, int gl, int gshift, V2<ClipT> *gbuf) {
            ClipT cl[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cl[2 * i] = (ClipT)f2[(F_RX + i) * TG + c];
                cl[2 * i + 1] = (ClipT)f2[(F_RZ + i) * TG + c];
            }
            ClipT vx = (ClipT)f1[(F_RX + (gl & 3)) * TQ + r], vy = (ClipT)f1[(F_RZ + (gl & 3)) * TQ + r];
            const int n = coop_clip_quads<ClipT>(cl, vx, vy, act ? 4 : 0, gl, gshift, gbuf);
            float area;
            if constexpr (sizeof(ClipT) == 8) area = coop_area_cython(vx, vy, n, gl);
            else area = coop_area_f32(vx, vy, n, gl);
            if (act && gl == 0) {
                PairTerms t = pair_terms(load_boxf(f1, r, TQ), load_boxf(f2, c, TG));
                if (hull && __fmul_rn(area, t.h) > 0.f) {   // box_ops3d.py:467: hull only where the intersection volume is > 0
                    float axv[4], azv[4], bxv[4], bzv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        axv[i] = f1[(F_RX + i) * TQ + r]; azv[i] = f1[(F_RZ + i) * TQ + r];
                        bxv[i] = f2[(F_RX + i) * TG + c]; bzv[i] = f2[(F_RZ + i) * TG + c];
                    }
                    t.encl = (float)hull_enclosing_volume(axv, azv, bxv, bzv, f1[F_YTOP * TQ + r], f1[F_YBOT * TQ + r],
                                                          f2[F_YTOP * TG + c], f2[F_YBOT * TG + c]);
                }
                tile[r * TG + c] = finish_pair(t, area, true, has_nums, inter_only);
            }
        };
        // Split mode (the shipped semantics: only GT columns < k2_cap are ever clipped, <= 64 eligible pairs per tile):
        // the last two warps find and clip those pairs WHILE the other six run phase A, instead of after it -- the fp64
        // clip chain (~2.3 us) and phase A (~2 us) overlap; every CTA of the one-wave grid finishes together.
        constexpr int NCW = 2, NWARP = NT / 32;
        const int clw = min(clip_lim - g0, ng);
        const int ncand = (rotated && clw > 0) ? nq * clw : 0;
        const bool split = ncand > 0 && ncand <= 32 * NCW;
        auto phase_a_row = [&](int r) {
            const BoxF qf = load_boxf(f1, r, TQ);   // broadcast loads
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                bool need_clip = false;
                if (gin[h]) {
                    PairTerms t = pair_terms(qf, gf[h]);
                    if (!gvalid[h]) t.nonrot = 0.f;   // box_util.py:562-564
                    float area = 0.f;
                    if (rotated) need_clip = gclip[h] && !(prefilter && t.nonrot == 0.f);
                    else area = t.nonrot;
                    if (!need_clip) tile[r * TG + lane + 32 * h] = finish_pair(t, area, gvalid[h], has_nums, inter_only);
                }
                if (split) continue;   // the clip warps find these pairs themselves
                const unsigned m = __ballot_sync(0xffffffffu, need_clip);   // warp-aggregated push
                if (m) {
                    int basepos = 0;
                    if (lane == 0) basepos = atomicAdd(&qcount, __popc(m));
                    basepos = __shfl_sync(0xffffffffu, basepos, 0);
                    if (need_clip) queue[basepos + __popc(m & ((1u << lane) - 1))] = (unsigned short)(r * TG + lane + 32 * h);
                }
            }
        };
        if (split) {
            if (warp >= NWARP - NCW) {
                const int cw = warp - (NWARP - NCW);
                const int idx = cw + NCW * lane;   // candidates interleaved over the clip warps
                bool need = false;
                if (idx < ncand) {
                    const int r = idx / clw, c = idx - r * clw;
                    const PairTerms t = pair_terms(load_boxf(f1, r, TQ), load_boxf(f2, c, TG));
                    need = !(prefilter && t.nonrot == 0.f);   // columns < clip_lim <= nk are valid
                }
                const unsigned m = __ballot_sync(0xffffffffu, need);
                const int n = __popc(m);
                {   // one round (four pairs) per clip warp under phase A ...
                    const int k = lane >> 3;
                    const bool act = k < n;
                    const int src = act ? (int)__fns(m, 0, k + 1) : 0;
                    const int sidx = cw + NCW * src;
                    const int r = sidx / clw, c = sidx - r * clw;
                    clip_pair_coop(r, c, act, lane & 7, lane & 24, scratch + (warp * 4 + (lane >> 3)) * 8);
                }
                if (n > 4) {   // ... the rest goes on the CTA queue and is drained by all 32 groups in phase B (clip-heavy tiles)
                    const int rank = __popc(m & ((1u << lane) - 1));
                    int basepos = 0;
                    if (lane == 0) basepos = atomicAdd(&qcount, n - 4);
                    basepos = __shfl_sync(0xffffffffu, basepos, 0);
                    if (need && rank >= 4) {
                        const int r = idx / clw, c = idx - r * clw;
                        queue[basepos + rank - 4] = (unsigned short)(r * TG + c);
                    }
                }
            } else {
                for (int r = warp; r < nq; r += NWARP - NCW) phase_a_row(r);
            }
        } else {
#pragma unroll
            for (int rr = 0; rr < TQ / NWARP; ++rr) {
                const int r = warp * (TQ / NWARP) + rr;
                if (r >= nq) break;   // warp-uniform
                phase_a_row(r);
            }
        }
        __syncthreads();
        GSTAMP(3);

        // ---- phase B: drain the clip queue, one Sutherland-Hodgman clip per lane
        const int nclip = qcount;
        if (p.dbg && threadIdx.x == 0) p.dbg[(size_t)blockIdx.x * 8 + 6] = nclip;
        constexpr int NGROUP = NT / 8;
        if (nclip <= 2 * NGROUP) {   // sparse (the shipped semantics: a few pairs per tile): 8 lanes per pair, short chain
            const int gl = lane & 7, gshift = lane & 24, group = threadIdx.x >> 3;
            V2<ClipT> *gbuf = scratch + group * 8;
            for (int base = 0; base < nclip; base += NGROUP) {
                if (base + warp * 4 >= nclip) break;   // warp-uniform
                const int qi = base + group;
                const bool act = qi < nclip;
                int r = 0, c = 0;
                if (act) { const int code = queue[qi]; r = code / TG; c = code - r * TG; }
                clip_pair_coop(r, c, act, gl, gshift, gbuf);
            }
        } else if (threadIdx.x < PB) {   // dense: one serial clip per lane, PB lanes
            V2<ClipT> *bufA = scratch + threadIdx.x;
            V2<ClipT> *bufB = scratch + SH_MAXV * PB + threadIdx.x;
            for (int qi = threadIdx.x; qi < nclip; qi += PB) {
                const int code = queue[qi];
                const int r = code / TG, c = code - r * TG;
                ClipT s[8], cl[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    s[2 * i] = (ClipT)f1[(F_RX + i) * TQ + r];
                    s[2 * i + 1] = (ClipT)f1[(F_RZ + i) * TQ + r];
                    cl[2 * i] = (ClipT)f2[(F_RX + i) * TG + c];
                    cl[2 * i + 1] = (ClipT)f2[(F_RZ + i) * TG + c];
                }
                float area;
                if constexpr (sizeof(ClipT) == 8) {
                    const int n = sh_clip_quads<double, PB>(s, cl, bufA, bufB);
                    area = area_cython<PB>(bufB, n);
                } else {
                    const int n = sh_clip_quads<float, PB>(s, cl, bufA, bufB);
                    area = area_f32<PB>(bufB, n);
                }
                PairTerms t = pair_terms(load_boxf(f1, r, TQ), load_boxf(f2, c, TG));
                if (hull && __fmul_rn(area, t.h) > 0.f) {   // box_ops3d.py:467: hull only where the intersection volume is > 0
                    float axv[4], azv[4], bxv[4], bzv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        axv[i] = f1[(F_RX + i) * TQ + r]; azv[i] = f1[(F_RZ + i) * TQ + r];
                        bxv[i] = f2[(F_RX + i) * TG + c]; bzv[i] = f2[(F_RZ + i) * TG + c];
                    }
                    t.encl = (float)hull_enclosing_volume(axv, azv, bxv, bzv, f1[F_YTOP * TQ + r], f1[F_YBOT * TQ + r],
                                                          f2[F_YTOP * TG + c], f2[F_YBOT * TG + c]);
                }
                tile[r * TG + c] = finish_pair(t, area, true, has_nums, inter_only);
            }
        }
        __syncthreads();
        GSTAMP(4);

        // ---- fused matcher cost: ((wc*-P[label] + wo*-obj) + wce*center) + wg*-giou
        if (p.epi.cost) {
            const MatcherEpi &e = p.epi;
            using A = Ar<float>;
            for (int r = warp; r < nq; r += NT / 32) {
                const size_t q = (size_t)b * p.K1 + q0 + r;
                const float om = -__ldg(e.obj + q);
                for (int c = lane; c < ng; c += 32) {
                    const size_t g = (size_t)b * p.K2 + g0 + c;
                    const size_t oidx = q * p.K2 + g0 + c;
                    long long lab = e.labels[g];
                    lab = lab < 0 ? 0 : (lab >= e.C ? e.C - 1 : lab);
                    const float cm = -__ldg(e.prob + q * e.C + lab);
                    float cen;
                    if (e.center_dist) cen = __ldg(e.center_dist + oidx);
                    else {
                        const float d0 = fabsf(A::sub(__ldg(e.cq + 3 * q), __ldg(e.cg + 3 * g)));
                        const float d1 = fabsf(A::sub(__ldg(e.cq + 3 * q + 1), __ldg(e.cg + 3 * g + 1)));
                        const float d2 = fabsf(A::sub(__ldg(e.cq + 3 * q + 2), __ldg(e.cg + 3 * g + 2)));
                        cen = A::add(A::add(d0, d1), d2);
                    }
                    const float gm = -tile[r * TG + c];
                    e.cost[oidx] = A::add(A::add(A::add(A::mul(e.wc, cm), A::mul(e.wo, om)), A::mul(e.wce, cen)), A::mul(e.wg, gm));
                }
            }
        }
        // ---- coalesced store of the [nq][ng] tile
        if (p.out) {
            float *o = p.out + ((size_t)b * p.K1 + q0) * p.K2 + g0;
            if (vec_out && ng == TG) {
                for (int i = threadIdx.x; i < nq * (TG / 4); i += NT) {
                    const int r = i / (TG / 4), c4 = i % (TG / 4);
                    *reinterpret_cast<float4 *>(o + (size_t)r * p.K2 + 4 * c4) = *reinterpret_cast<const float4 *>(tile + r * TG + 4 * c4);
                }
            } else {
                for (int r = warp; r < nq; r += NT / 32)
                    for (int c = lane; c < ng; c += 32) o[(size_t)r * p.K2 + c] = tile[r * TG + c];
            }
        }
        __syncthreads();
        GSTAMP(5);
    }
}

template <typename ClipT, int PB, int TQ> static size_t giou_smem_bytes()
{
    return sizeof(float) * (NF * TQ + NF * TG + TQ * TG + (TQ + TG) * 24) + sizeof(unsigned short) * TQ * TG +
           sizeof(V2<ClipT>) * 2 * SH_MAXV * PB;
}

template <typename ClipT, int PB, int TQ, bool HULL> static int launch_giou_tq2(GiouParams p, cudaStream_t st)
{
    p.tiles_per_b = (p.K1 + TQ - 1) / TQ;
    { const char *e = getenv("OVDET_GIOU_DBG_PTR"); p.dbg = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0)) : nullptr; }
    const size_t smem = giou_smem_bytes<ClipT, PB, TQ>();
    static bool attr_set_dev[OVDET_MAX_DEVICES] = {};   // the attribute is per device: set it once on each
    int dev_ = 0;
    OVDET_CUDA_TRY(cudaGetDevice(&dev_));
    bool untracked = false;   // a device index beyond the table: set the attribute on every call
    bool &attr_set = (dev_ >= 0 && dev_ < OVDET_MAX_DEVICES) ? attr_set_dev[dev_] : untracked;
    if (!attr_set) {
        OVDET_CUDA_TRY(cudaFuncSetAttribute(giou3d_kernel<ClipT, PB, TQ, HULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long grid = (long long)p.B * p.tiles_per_b;
    static int pdl = -1;   // OVDET_GIOU_PDL=0 switches programmatic dependent launch off (A/B measurements)
    if (pdl < 0) { const char *e = getenv("OVDET_GIOU_PDL"); pdl = e ? atoi(e) : 1; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    OVDET_CUDA_TRY(cudaLaunchKernelEx(&cfg, giou3d_kernel<ClipT, PB, TQ, HULL>, p));
    return launch_ok("giou3d_kernel");
}


template <typename ClipT, int PB, int TQ> static int launch_giou_tq(const GiouParams &p, cudaStream_t st)
{
    if (p.flags & OVDET_GIOU_ENCL_HULL) return launch_giou_tq2<ClipT, PB, TQ, true>(p, st);
    return launch_giou_tq2<ClipT, PB, TQ, false>(p, st);
}

// Tile height.  The workload is tiny (config 1 = 64 x 128 x 64 pairs): latency, not throughput, decides.
template <typename ClipT, int PB> static int launch_giou(const GiouParams &p, cudaStream_t st)
{
    OVDET_REQUIRE((long long)p.B * ((p.K1 + 7) / 8) < 2147483647LL, "grid too large");
    static int force = -1;   // OVDET_GIOU_TQ=8|16|32 overrides the heuristic (profiling experiments)
    if (force < 0) { const char *e = getenv("OVDET_GIOU_TQ"); force = e ? atoi(e) : 0; }
    if (force == 32) return launch_giou_tq<ClipT, PB, 32>(p, st);
    if (force == 16) return launch_giou_tq<ClipT, PB, 16>(p, st);
    if (force == 8) return launch_giou_tq<ClipT, PB, 8>(p, st);
    // 16-row tiles everywhere (measured on B200, profiles/r1_notes.md): <= 64 registers, 4 CTAs/SM, config 1 is one wave of
    // 512 CTAs; with the split mode and the cooperative clip they also beat 32-row tiles in the many-wave regime
    // (4096 box sets: 115 vs 98 Gpairs/s as shipped, 95 vs 92 torch path, 33.5 vs 29.3 every pair clipped) and 8-row
    // tiles everywhere (65 / 26 Gpairs/s).  OVDET_GIOU_TQ keeps the other heights reachable for experiments.
    return launch_giou_tq<ClipT, PB, 16>(p, st);
}

// ---------------------------------------------------------------------------
// Cython ABI kernel: box_intersection(rect1, rect2, non_rot, nums_k2, inter_areas, approximate)
// one CTA per 2048 consecutive (b,k1,k2) pairs, same queue + drain structure.
// ---------------------------------------------------------------------------
struct BiParams {
    const float *rect1, *rect2, *nonrot;
    const int32_t *nums_k2;
    float *inter;
    int B, K1, K2, k2_loop, approximate;
    long long total;
};
constexpr int BI_PAIRS = 2048;

__global__ void __launch_bounds__(NT) box_intersection_kernel(BiParams p)
{
    __shared__ unsigned short queue[BI_PAIRS];
    __shared__ int qcount;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V2<double> *scratch = reinterpret_cast<V2<double> *>(smem_raw);
    const long long base0 = (long long)blockIdx.x * BI_PAIRS;
    if (threadIdx.x == 0) qcount = 0;
    __syncthreads();
    for (int off = 0; off < BI_PAIRS; off += NT) {
        const long long idx = base0 + off + threadIdx.x;
        bool need = false;
        if (idx < p.total) {
            const int k2 = (int)(idx % p.K2);
            const long long bk1 = idx / p.K2;
            const int b = (int)(bk1 / p.K1);
            need = k2 < p.k2_loop && k2 < p.nums_k2[b] && !(p.approximate && p.nonrot[idx] == 0.f);
        }
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (m) {
            const int lane = threadIdx.x & 31;
            int bp = 0;
            if (lane == 0) bp = atomicAdd(&qcount, __popc(m));
            bp = __shfl_sync(0xffffffffu, bp, 0);
            if (need) queue[bp + __popc(m & ((1u << lane) - 1))] = (unsigned short)(off + threadIdx.x);
        }
    }
    __syncthreads();
    const int n = qcount;
    V2<double> *bufA = scratch + threadIdx.x, *bufB = scratch + SH_MAXV * NT + threadIdx.x;
    for (int qi = threadIdx.x; qi < n; qi += NT) {
        const long long idx = base0 + queue[qi];
        const int k2 = (int)(idx % p.K2);
        const long long bk1 = idx / p.K2;
        const int b = (int)(bk1 / p.K1);
        const float *r1 = p.rect1 + bk1 * 8;
        const float *r2 = p.rect2 + ((long long)b * p.K2 + k2) * 8;
        double s[8], c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] = (double)__ldg(r1 + i); c[i] = (double)__ldg(r2 + i); }
        const int n = sh_clip_quads<double, NT>(s, c, bufA, bufB);
        if (n > 0) p.inter[idx] = area_cython<NT>(bufB, n);  // pyx:195-198: untouched when the clip is empty
    }
}

__global__ void __launch_bounds__(256) box_corners_kernel(const float *__restrict__ center, const float *__restrict__ size,
                                                          const float *__restrict__ angle, long long n, float *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float c[24];
    decode_box(center + 3 * i, size + 3 * i, __ldg(angle + i), c);
#pragma unroll
    for (int k = 0; k < 24; ++k) out[i * 24 + k] = c[k];
}

__global__ void __launch_bounds__(256) matcher_cost_kernel(MatcherEpi e, const float *__restrict__ gious, int Q, int G, long long total)
{
    using A = Ar<float>;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long q = idx / G;           // b*Q + q
        const int gcol = (int)(idx - q * G);
        const long long b = q / Q;
        const long long g = b * G + gcol;
        long long lab = e.labels[g];
        lab = lab < 0 ? 0 : (lab >= e.C ? e.C - 1 : lab);
        const float cm = -__ldg(e.prob + q * e.C + lab);
        const float om = -__ldg(e.obj + q);
        float cen;
        if (e.center_dist) cen = __ldg(e.center_dist + idx);
        else {
            const float d0 = fabsf(A::sub(__ldg(e.cq + 3 * q), __ldg(e.cg + 3 * g)));
            const float d1 = fabsf(A::sub(__ldg(e.cq + 3 * q + 1), __ldg(e.cg + 3 * g + 1)));
            const float d2 = fabsf(A::sub(__ldg(e.cq + 3 * q + 2), __ldg(e.cg + 3 * g + 2)));
            cen = A::add(A::add(d0, d1), d2);
        }
        const float gm = -__ldg(gious + idx);
        e.cost[idx] = A::add(A::add(A::add(A::mul(e.wc, cm), A::mul(e.wo, om)), A::mul(e.wce, cen)), A::mul(e.wg, gm));
    }
}

static int matcher_cost_elementwise(const MatcherEpi &e, const float *gious, int B, int Q, int G, cudaStream_t st)
{
    const long long total = (long long)B * Q * G;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    matcher_cost_kernel<<<(unsigned)blocks, 256, 0, st>>>(e, gious, Q, G, total);
    return launch_ok("matcher_cost_kernel");
}

int giou3d_launch(const GiouParams &p, cudaStream_t st)
{
    if ((p.flags & OVDET_GIOU_ROTATED) && (p.flags & OVDET_GIOU_CLIP_F64)) return launch_giou<double, 128>(p, st);
    return launch_giou<float, 256>(p, st);
}

}  // namespace ovdet

using namespace ovdet;

extern "C" int ovdet_giou3d_f32(const float *corners1, const float *corners2, const int64_t *nums_k2,
                                int B, int K1, int K2, int k2_cap, unsigned flags, float *out, void *stream)
{
    OVDET_REQUIRE(B >= 0 && K1 >= 0 && K2 >= 0, "negative size");
    if (B == 0 || K1 == 0 || K2 == 0) return OVDET_OK;
    OVDET_REQUIRE(corners1 && corners2 && out, "null pointer");
    GiouParams p;
    p.c1 = corners1; p.c2 = corners2; p.nums_k2 = nums_k2; p.out = out;
    p.B = B; p.K1 = K1; p.K2 = K2; p.k2_cap = k2_cap; p.flags = flags;
    p.tiles_per_b = 0;
    return giou3d_launch(p, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ovdet_box_corners_f32(const float *center, const float *size, const float *angle, int64_t n, float *corners, void *stream)
{
    OVDET_REQUIRE(n >= 0, "negative size");
    if (n == 0) return OVDET_OK;
    OVDET_REQUIRE(center && size && angle && corners, "null pointer");
    box_corners_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(center, size, angle, n, corners);
    return launch_ok("box_corners_kernel");
}

extern "C" int ovdet_giou3d_decode_f32(const float *center1, const float *size1, const float *angle1, const float *corners2,
                                       const int64_t *nums_k2, int B, int K1, int K2, int k2_cap, unsigned flags,
                                       float *out, float *corners1_out, void *stream)
{
    OVDET_REQUIRE(B >= 0 && K1 >= 0 && K2 >= 0, "negative size");
    if (B == 0 || K1 == 0 || K2 == 0) return OVDET_OK;
    OVDET_REQUIRE(center1 && size1 && angle1 && corners2 && out, "null pointer");
    GiouParams p;
    p.c1 = nullptr; p.c2 = corners2; p.nums_k2 = nums_k2; p.out = out;
    p.B = B; p.K1 = K1; p.K2 = K2; p.k2_cap = k2_cap; p.flags = flags; p.tiles_per_b = 0;
    p.dec1.center = center1; p.dec1.size = size1; p.dec1.angle = angle1; p.dec1.corners_out = corners1_out;
    return giou3d_launch(p, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ovdet_matcher_cost_f32(const float *sem_cls_prob, const float *objectness, const float *center_dist,
                                      const float *center_q, const float *center_g, const float *gious,
                                      const float *corners1, const float *corners2, const int64_t *gt_labels,
                                      const int64_t *nactual_gt, int B, int Q, int G, int C,
                                      float w_class, float w_obj, float w_center, float w_giou,
                                      unsigned giou_flags, int k2_cap, float *gious_out, float *cost, void *stream)
{
    OVDET_REQUIRE(B >= 0 && Q >= 0 && G >= 0 && C > 0, "bad size");
    if (B == 0 || Q == 0 || G == 0) return OVDET_OK;
    OVDET_REQUIRE(sem_cls_prob && objectness && gt_labels && cost, "null pointer");
    OVDET_REQUIRE(center_dist || (center_q && center_g), "need center_dist or center_q/center_g");
    OVDET_REQUIRE(gious || (corners1 && corners2), "need gious or corners");
    MatcherEpi e;
    e.prob = sem_cls_prob; e.obj = objectness; e.center_dist = center_dist; e.cq = center_q; e.cg = center_g;
    e.labels = gt_labels; e.cost = cost; e.C = C; e.wc = w_class; e.wo = w_obj; e.wce = w_center; e.wg = w_giou;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (gious) return matcher_cost_elementwise(e, gious, B, Q, G, st);
    GiouParams p;
    p.c1 = corners1; p.c2 = corners2; p.nums_k2 = nactual_gt; p.out = gious_out;
    p.B = B; p.K1 = Q; p.K2 = G; p.k2_cap = k2_cap; p.flags = giou_flags & ~OVDET_GIOU_INTER_ONLY;
    p.tiles_per_b = 0;
    p.epi = e;
    return giou3d_launch(p, st);
}

extern "C" int ovdet_box_intersection_f32(const float *rect1, const float *rect2, const float *non_rot_inter_areas,
                                          const int32_t *nums_k2, float *inter_areas, int approximate,
                                          int B, int K1, int K2, int k2_loop, void *stream)
{
    OVDET_REQUIRE(B >= 0 && K1 >= 0 && K2 >= 0, "negative size");
    if (B == 0 || K1 == 0 || K2 == 0) return OVDET_OK;
    OVDET_REQUIRE(rect1 && rect2 && nums_k2 && inter_areas, "null pointer");
    OVDET_REQUIRE(!approximate || non_rot_inter_areas, "approximate needs non_rot_inter_areas");
    BiParams p;
    p.rect1 = rect1; p.rect2 = rect2; p.nonrot = non_rot_inter_areas; p.nums_k2 = nums_k2; p.inter = inter_areas;
    p.B = B; p.K1 = K1; p.K2 = K2; p.k2_loop = k2_loop; p.approximate = approximate;
    p.total = (long long)B * K1 * K2;
    const size_t smem = sizeof(V2<double>) * 2 * SH_MAXV * NT;
    static bool attr_set_dev[OVDET_MAX_DEVICES] = {};   // the attribute is per device: set it once on each
    int dev_ = 0;
    OVDET_CUDA_TRY(cudaGetDevice(&dev_));
    bool untracked = false;   // a device index beyond the table: set the attribute on every call
    bool &attr_set = (dev_ >= 0 && dev_ < OVDET_MAX_DEVICES) ? attr_set_dev[dev_] : untracked;
    if (!attr_set) {
        OVDET_CUDA_TRY(cudaFuncSetAttribute(box_intersection_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long grid = (p.total + BI_PAIRS - 1) / BI_PAIRS;
    OVDET_REQUIRE(grid < 2147483647LL, "grid too large");
    box_intersection_kernel<<<(unsigned)grid, NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return launch_ok("box_intersection_kernel");
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int ovdet_box_intersection_host_f32(const float *rect1, const float *rect2, const float *non_rot_inter_areas,
                                               const int32_t *nums_k2, float *inter_areas, int approximate,
                                               int B, int K1, int K2, int k2_loop)
{
    OVDET_REQUIRE(B >= 0 && K1 >= 0 && K2 >= 0, "negative size");
    if (B == 0 || K1 == 0 || K2 == 0) return OVDET_OK;
    OVDET_REQUIRE(rect1 && rect2 && nums_k2 && inter_areas, "null pointer");
    // Only the first kc = min(k2_loop, K2) GT columns are ever touched (pyx:180 loops k2 < k2_loop).  The caller's numpy
    // buffers are pageable, and every pageable copy costs ~15-25 us whatever its size, so the call is packed: rect1, the kc
    // used rows of rect2, the kc used columns of non_rot_inter_areas / inter_areas and nums_k2 go into ONE pinned buffer
    // and ONE upload; the kernel runs on the compact [B,K1,kc] problem; ONE download brings the kc columns back and they
    // are scattered into inter_areas (in-place contract: entries the loop does not reach keep their values).
    // Measured on B200 at one decoder layer (8x128x64, k2_loop 4): 150 us with six pageable copies -> 52 us packed; the
    // reference's compiled Cython loop takes 320-590 us for the same call on the same host.
    const int kc = k2_loop < K2 ? (k2_loop < 0 ? 0 : k2_loop) : K2;
    if (kc == 0) return OVDET_OK;
    const size_t n1 = (size_t)B * K1 * 32, n2 = (size_t)B * kc * 32, np = (size_t)B * K1 * kc * 4, nn = (size_t)B * 4;
    HostStaging &hs = host_staging();
    const size_t o1 = 0, o2 = o1 + align256(n1), o3 = o2 + align256(n2), o4 = o3 + align256(np), o5 = o4 + align256(np);
    const size_t total = o5 + align256(nn);
    int rc = hs.ensure(total);
    if (rc) return rc;
    rc = hs.ensure_pinned(total);
    if (rc) return rc;
    char *d = static_cast<char *>(hs.dev);
    char *h = static_cast<char *>(hs.pinned);
    memcpy(h + o1, rect1, n1);
    for (int b = 0; b < B; ++b) memcpy(h + o2 + (size_t)b * kc * 32, rect2 + (size_t)b * K2 * 8, (size_t)kc * 32);
    {
        float *hn = reinterpret_cast<float *>(h + o3), *hi = reinterpret_cast<float *>(h + o4);
        const size_t rows = (size_t)B * K1;
        for (size_t r = 0; r < rows; ++r) {
            if (non_rot_inter_areas) memcpy(hn + r * kc, non_rot_inter_areas + r * K2, (size_t)kc * 4);
            memcpy(hi + r * kc, inter_areas + r * K2, (size_t)kc * 4);
        }
    }
    memcpy(h + o5, nums_k2, nn);
    OVDET_CUDA_TRY(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, hs.stream));
    rc = ovdet_box_intersection_f32((const float *)(d + o1), (const float *)(d + o2),
                                    non_rot_inter_areas ? (const float *)(d + o3) : nullptr, (const int32_t *)(d + o5),
                                    (float *)(d + o4), approximate, B, K1, kc, kc, hs.stream);
    if (rc) return rc;
    OVDET_CUDA_TRY(cudaMemcpyAsync(h + o4, d + o4, np, cudaMemcpyDeviceToHost, hs.stream));
    OVDET_CUDA_TRY(cudaStreamSynchronize(hs.stream));
    {
        const float *hi = reinterpret_cast<const float *>(h + o4);
        const size_t rows = (size_t)B * K1;
        for (size_t r = 0; r < rows; ++r) memcpy(inter_areas + r * K2, hi + r * kc, (size_t)kc * 4);
    }
    return OVDET_OK;
}

extern "C" int ovdet_giou3d_host_f32(const float *corners1, const float *corners2, const int64_t *nums_k2,
                                     int B, int K1, int K2, int k2_cap, unsigned flags, float *out)
{
    OVDET_REQUIRE(B >= 0 && K1 >= 0 && K2 >= 0, "negative size");
    if (B == 0 || K1 == 0 || K2 == 0) return OVDET_OK;
    OVDET_REQUIRE(corners1 && corners2 && out, "null pointer");
    const size_t n1 = (size_t)B * K1 * 96, n2 = (size_t)B * K2 * 96, no = (size_t)B * K1 * K2 * 4, nn = (size_t)B * 8;
    HostStaging &hs = host_staging();
    const size_t o1 = 0, o2 = o1 + align256(n1), o3 = o2 + align256(n2), o4 = o3 + align256(no);
    int rc = hs.ensure(o4 + align256(nn));
    if (rc) return rc;
    char *d = static_cast<char *>(hs.dev);
    // The call is bound by the PCIe transfers and their fixed latencies (config 1: 1.2 MB up, 2.1 MB down, ~7 us of
    // kernel).  Measured on B200 (profiles/r1_notes.md): a pinned `out` is written by the kernel itself through its
    // mapped address -- the read-back then overlaps the compute and needs no copy call (128 -> 89 us); reading the
    // inputs the same way is slower than a copy (SM loads over PCIe are latency-bound).  Splitting the batch into chunks
    // on two streams (upload of chunk i+1 under the write-back of chunk i) is kept behind OVDET_HOST_CHUNKS for
    // experiments: at this size the extra calls cost more than the overlap wins (98 -> 114 us with 2 chunks).
    float *dout = reinterpret_cast<float *>(d + o3);
    bool mapped = false;
    static int zc_env = -1;   // OVDET_HOST_ZEROCOPY=0 forces the copy-back path (A/B measurements)
    if (zc_env < 0) { const char *e = getenv("OVDET_HOST_ZEROCOPY"); zc_env = e ? atoi(e) : 1; }
    if (zc_env) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            dout = static_cast<float *>(at.devicePointer); mapped = true;
        } else cudaGetLastError();
    }
    static int nchunk_env = -1;
    if (nchunk_env < 0) { const char *e = getenv("OVDET_HOST_CHUNKS"); nchunk_env = e ? atoi(e) : 0; }
    int nchunk = nchunk_env > 0 ? nchunk_env : 1;   // measured: 2+ chunks lose more to per-call latencies than the overlap wins
    nchunk = nchunk > 8 ? 8 : (nchunk > B ? B : nchunk);
    cudaStream_t sts[2] = {hs.stream, nchunk > 1 ? hs.stream_d2h : hs.stream};
    if (nums_k2) {
        OVDET_CUDA_TRY(cudaMemcpyAsync(d + o4, nums_k2, nn, cudaMemcpyHostToDevice, hs.stream));
        if (nchunk > 1) {
            OVDET_CUDA_TRY(cudaEventRecord(hs.ev[0], hs.stream));
            OVDET_CUDA_TRY(cudaStreamWaitEvent(sts[1], hs.ev[0], 0));
        }
    }
    for (int c = 0; c < nchunk; ++c) {
        const int b0 = (int)((long long)B * c / nchunk), b1 = (int)((long long)B * (c + 1) / nchunk), nb = b1 - b0;
        if (nb == 0) continue;
        cudaStream_t st = sts[c & 1];
        const size_t s1 = (size_t)b0 * K1 * 96, s2 = (size_t)b0 * K2 * 96, so = (size_t)b0 * K1 * K2;
        OVDET_CUDA_TRY(cudaMemcpyAsync(d + o1 + s1, (const char *)corners1 + s1, (size_t)nb * K1 * 96, cudaMemcpyHostToDevice, st));
        OVDET_CUDA_TRY(cudaMemcpyAsync(d + o2 + s2, (const char *)corners2 + s2, (size_t)nb * K2 * 96, cudaMemcpyHostToDevice, st));
        rc = ovdet_giou3d_f32((const float *)(d + o1 + s1), (const float *)(d + o2 + s2),
                              nums_k2 ? (const int64_t *)(d + o4) + b0 : nullptr, nb, K1, K2, k2_cap, flags, dout + so, st);
        if (rc) return rc;
        if (!mapped) OVDET_CUDA_TRY(cudaMemcpyAsync(out + so, dout + so, (size_t)nb * K1 * K2 * 4, cudaMemcpyDeviceToHost, st));
    }
    if (nchunk > 1) OVDET_CUDA_TRY(cudaStreamSynchronize(sts[1]));
    OVDET_CUDA_TRY(cudaStreamSynchronize(hs.stream));
    return OVDET_OK;
}
