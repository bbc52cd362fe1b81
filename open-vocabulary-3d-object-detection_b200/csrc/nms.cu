// nms.cu -- greedy NMS family for sm_100a: one CTA per scene.
//
// Replaces nms_2d_faster / nms_3d_faster / nms_3d_faster_samecls (utils/nms.py:43-162),
// the tools variant (3DOVDet_tools/utils/box_3d_utils.py:60-120), the NMS branches
// of parse_predictions (utils/ap_calculator.py:86-189) and the per-scene
// NMS -> pool match -> size-scored NMS of 3DOVDet_tools/scannet/lift_boxes.py:139-166.
//
// Algorithm (same picks as the reference's while-loop on tie-free scores):
//   1. bitonic sort of (score desc) indices in shared memory;
//   2. boxes gathered in sorted order (SoA, conflict-free);
//   3. suppression bitmask: warp w owns rows i = w, w+nw, ...; its lanes test 32
//      columns j > i at a time in fp64 (no FMA contraction, same op order as
//      numpy) and one __ballot_sync packs the 32 verdicts into a mask word;
//   4. one warp scans the rows in score order keeping the `removed` bitset
//      distributed one 32-bit word per lane (K <= 1024): bit test by shuffle,
//      OR of the picked row's mask words.
//   3'/4' (class-wise NMS with thr >= 0 -- parse_predictions' default, the tools' class_wise=True): boxes of
//      different classes never suppress each other, so the scene splits into independent per-class problems.
//      Positions are grouped by class (stable counting sort with __match_any_sync ranks), then one warp per class
//      runs the greedy loop directly -- sum_c n_c^2/2 pair tests instead of K^2/2 (20x fewer at 20 classes), no
//      mask matrix, no global ordered scan.  Same arithmetic per tested pair, same picks.
// All compares are IEEE fp64 like the reference (np.zeros default dtype,
// ap_calculator.py:157), so keep-indices are bit-exact except on score ties
// (numpy's default argsort is not stable: documented).
#include <math.h>
#include <stdlib.h>

#include "nms_core.cuh"

namespace ovdet {

// ------------------------------------------------------------------ plain NMS
struct ArraySrc {
    const double *b; int ncols, dims, n; bool has_cls;
    __device__ bool alive(int k) const { return k < n; }
    __device__ double score(int k) const { return b[(size_t)k * ncols + 2 * dims]; }
    __device__ double cls_of(int k) const { return has_cls ? b[(size_t)k * ncols + 2 * dims + 1] : 0.0; }
    __device__ void box(int k, double *lo, double *hi, double &cl) const
    {
        const double *r = b + (size_t)k * ncols;
        for (int a = 0; a < dims; ++a) { lo[a] = r[a]; hi[a] = r[dims + a]; }
        cl = has_cls ? r[2 * dims + 1] : 0.0;
    }
};

struct NmsParams {
    const double *boxes; const int32_t *counts;
    int S, K, ncols; double thr, eps; unsigned flags;
    uint8_t *keep; int32_t *pick_order; int32_t *npick;
};

__global__ void __launch_bounds__(NMS_NT) nms_kernel(NmsParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int s = blockIdx.x;
    NmsSmem sh = nms_carve(sm, p.K);
    const int dims = (p.flags & OVDET_NMS_2D) ? 2 : 3;
    const bool samecls = p.flags & OVDET_NMS_SAMECLS;
    int n = p.counts ? p.counts[s] : p.K;
    n = n < 0 ? 0 : (n > p.K ? p.K : n);
    ArraySrc src{p.boxes + (size_t)s * p.K * p.ncols, p.ncols, dims, n, samecls};
    for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
        p.keep[(size_t)s * p.K + k] = 0;
        if (p.pick_order) p.pick_order[(size_t)s * p.K + k] = -1;
    }
    __syncthreads();
    nms_core(src, p.K, dims, samecls, p.flags & OVDET_NMS_OLD_TYPE, p.thr, p.eps, sh,
             p.pick_order ? p.pick_order + (size_t)s * p.K : nullptr, nullptr, (p.flags & OVDET_NMS_LHS) != 0);
    const int na = sh.misc[0];
    for (int pos = threadIdx.x; pos < na; pos += blockDim.x)
        if (sh.picked[pos]) p.keep[(size_t)s * p.K + sh.sidx[pos]] = 1;
    if (threadIdx.x == 0 && p.npick) p.npick[s] = sh.misc[1];
}


struct ParseParams {
    const float *corners, *probs, *obj; const uint8_t *nonempty;
    int S, K, C; double nms_iou; float conf; unsigned flags;
    uint8_t *pred_mask, *keep; int32_t *pred_cls; float *pred_cls_prob;
    unsigned long long *dbg;   // optional [S][16] globaltimer stamps of thread 0 (OVDET_PARSE_DBG_PTR; null in production)
};

__global__ void __launch_bounds__(NMS_NT) parse_predictions_kernel(ParseParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int s = blockIdx.x;
    NmsSmem sh = nms_carve(sm, p.K);
    int *cls_sm = reinterpret_cast<int *>(sm + nms_smem_bytes(p.K));
    const float *probs = p.probs + (size_t)s * p.K * p.C;
    const float *obj = p.obj + (size_t)s * p.K;
    const uint8_t *ne = p.nonempty ? p.nonempty + (size_t)s * p.K : nullptr;
    NSTAMP(p.dbg, 0);
    // argmax / max class prob (ap_calculator.py:59-61; np.argmax = first maximum).  The scene's [K, C] tile is staged
    // through shared memory (aliasing the NMS tables, not yet in use) with independent 16-byte loads, so that no thread
    // walks C dependent global loads; row pitch C+1 keeps the per-thread row walks conflict-free.
    {
        float *ptile = reinterpret_cast<float *>(sm);
        const size_t kc = (size_t)p.K * p.C;
        const bool staged = (size_t)p.K * (p.C + 1) * sizeof(float) <= nms_smem_bytes(p.K);
        if (staged) {
            const float inv_c = 1.f / (float)p.C;
            auto rowcol = [&](int e, int &r, int &c) {   // e / C without an integer divide (e < 2^21: the float estimate is off by at most one)
                r = (int)(((float)e + 0.5f) * inv_c);
                c = e - r * p.C;
                if (c < 0) { --r; c += p.C; } else if (c >= p.C) { ++r; c -= p.C; }
            };
            if ((reinterpret_cast<uintptr_t>(probs) & 15) == 0 && (kc & 3) == 0) {
                for (int i = threadIdx.x; i < (int)(kc >> 2); i += blockDim.x) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(probs) + i);
                    const float vv[4] = {v.x, v.y, v.z, v.w};
                    int r, c;
                    rowcol(4 * i, r, c);
#pragma unroll
                    for (int q = 0; q < 4; ++q) { ptile[r * (p.C + 1) + c] = vv[q]; if (++c == p.C) { c = 0; ++r; } }
                }
            } else {
                for (int e = threadIdx.x; e < (int)kc; e += blockDim.x) { int r, c; rowcol(e, r, c); ptile[r * (p.C + 1) + c] = __ldg(probs + e); }
            }
            __syncthreads();
        }
        int my_cls[NMS_MAXK / 128];   // a thread owns at most K / blockDim.x <= 8 boxes
        int cnt = 0;
        for (int k = threadIdx.x; k < p.K; k += blockDim.x, ++cnt) {
            const float *row = staged ? ptile + (size_t)k * (p.C + 1) : nullptr;
            float best = staged ? row[0] : __ldg(probs + (size_t)k * p.C);
            int bi = 0;
            for (int c = 1; c < p.C; ++c) { const float v = staged ? row[c] : __ldg(probs + (size_t)k * p.C + c); if (v > best) { best = v; bi = c; } }
            my_cls[cnt] = bi;
            p.pred_cls[(size_t)s * p.K + k] = bi;
            p.pred_cls_prob[(size_t)s * p.K + k] = best;
            p.pred_mask[(size_t)s * p.K + k] = 0;
        }
        __syncthreads();   // the tile aliases the NMS tables and cls_sm's neighbours: everyone is done reading it
        cnt = 0;
        for (int k = threadIdx.x; k < p.K; k += blockDim.x, ++cnt) cls_sm[k] = my_cls[cnt];
    }
    __syncthreads();
    if (p.flags & OVDET_PARSE_NO_NMS) {
        for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
            const uint8_t m = ne ? (ne[k] != 0) : 1;
            p.pred_mask[(size_t)s * p.K + k] = m;
            p.keep[(size_t)s * p.K + k] = m && (obj[k] > p.conf);
        }
        return;
    }
    const bool d2 = p.flags & OVDET_NMS_2D;
    const bool samecls = p.flags & OVDET_NMS_SAMECLS;
    CornerSrc src{p.corners + (size_t)s * p.K * 24, obj, ne, cls_sm, d2 ? 1 : 0};
    NSTAMP(p.dbg, 1);
    nms_core(src, p.K, d2 ? 2 : 3, samecls, p.flags & OVDET_NMS_OLD_TYPE, p.nms_iou, 0.0, sh, nullptr, p.dbg);
    NSTAMP(p.dbg, 6);
    const int na = sh.misc[0];
    for (int pos = threadIdx.x; pos < na; pos += blockDim.x)
        if (sh.picked[pos]) p.pred_mask[(size_t)s * p.K + sh.sidx[pos]] = 1;
    __syncthreads();
    for (int k = threadIdx.x; k < p.K; k += blockDim.x)
        p.keep[(size_t)s * p.K + k] = p.pred_mask[(size_t)s * p.K + k] && (obj[k] > p.conf);
    NSTAMP(p.dbg, 7);
}

// ------------------------------------------------------- pseudo-label filter
struct PoolSrc {
    const double *pool; const double *label; const double *tmp_score; int n;
    __device__ bool alive(int k) const { return k < n && label[k] != -100.0; }
    __device__ double cls_of(int k) const { return label[k]; }
    __device__ double score(int k) const
    {   // use_size_score: score *= size, size = prod(scale) (lift_boxes.py:159-160, box_3d_utils.py:76-79)
        const double *r = pool + (size_t)k * 6;
        const double v = __dmul_rn(__dmul_rn(__dsub_rn(r[3], r[0]), __dsub_rn(r[4], r[1])), __dsub_rn(r[5], r[2]));
        return __dmul_rn(tmp_score[k], v);
    }
    __device__ void box(int k, double *lo, double *hi, double &cl) const
    {
        const double *r = pool + (size_t)k * 6;
        for (int a = 0; a < 3; ++a) { lo[a] = r[a]; hi[a] = r[3 + a]; }
        cl = label[k];
    }
};

struct PseudoParams {
    const double *boxes, *pool; const int32_t *nboxes, *npool;
    int S, P, M, Kmax; double nms_thr, match_thr, size_thr;
    uint8_t *nms1_keep; double *out_label, *out_score; uint8_t *out_keep;
    unsigned long long *dbg;   // optional [S][16] globaltimer stamps (OVDET_PSEUDO_DBG_PTR; null in production)
};

__global__ void __launch_bounds__(NMS_NT) pseudo_filter_kernel(PseudoParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    using A = Ar<double>;
    const int s = blockIdx.x;
    NmsSmem sh = nms_carve(sm, p.Kmax);
    // [M] earliest claiming pick and [P] pick order of NMS #1 live in the sort-key array: it is dead once NMS #1 has sorted
    // (the pick order is written after that) and NMS #2 only rewrites it after both are consumed.  (M + P) * 4 <= Kp * 8.
    // Without these 3 KB a third CTA fits per SM at the config-5 shape.
    unsigned int *best = reinterpret_cast<unsigned int *>(sh.skey);
    int *order = reinterpret_cast<int *>(best + p.M);
    int nb = p.nboxes ? p.nboxes[s] : p.P;
    nb = nb < 0 ? 0 : (nb > p.P ? p.P : nb);
    int np_ = p.npool ? p.npool[s] : p.M;
    np_ = np_ < 0 ? 0 : (np_ > p.M ? p.M : np_);
    const double *boxes = p.boxes + (size_t)s * p.P * 8;
    const double *pool = p.pool + (size_t)s * p.M * 6;
    double *olab = p.out_label + (size_t)s * p.M, *osc = p.out_score + (size_t)s * p.M;
    for (int k = threadIdx.x; k < p.P; k += blockDim.x) p.nms1_keep[(size_t)s * p.P + k] = 0;
    for (int j = threadIdx.x; j < p.M; j += blockDim.x) { olab[j] = -100.0; osc[j] = 0.0; p.out_keep[(size_t)s * p.M + j] = 0; }
    __syncthreads();
    NSTAMP(p.dbg, 0);
    // 1. class-wise NMS (lift_boxes.py:140; volume + 1e-8, box_3d_utils.py:74)
    ArraySrc src{boxes, 8, 3, nb, true};
    nms_core(src, p.P, 3, true, false, p.nms_thr, 1e-8, sh, order);
    const int npick = sh.misc[1];
    {
        const int na = sh.misc[0];
        for (int pos = threadIdx.x; pos < na; pos += blockDim.x)
            if (sh.picked[pos]) p.nms1_keep[(size_t)s * p.P + sh.sidx[pos]] = 1;
    }
    for (int j = threadIdx.x; j < p.M; j += blockDim.x) best[j] = 0xFFFFFFFFu;
    __syncthreads();
    NSTAMP(p.dbg, 1);
    // 2. argmax-IoU match to the pool (lift_boxes.py:151-158): one warp per surviving box.  The pool boxes and their
    // volumes are staged once in shared memory (the suppression-word region, idle between the two NMS passes); a pair
    // with an empty overlap on some axis has IoU = +0 / positive = 0 exactly, so its fp64 divide is skipped.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wm = (p.Kmax + 31) / 32;
    const bool pool_staged = (size_t)np_ * 7 * sizeof(double) <= sizeof(uint32_t) * (size_t)p.Kmax * Wm;
    double *pl = reinterpret_cast<double *>(sh.mask);   // [np_][6] then kv[np_]
    double *pkv = pl + (size_t)np_ * 6;
    if (pool_staged) {
        for (int i = threadIdx.x; i < np_ * 6; i += blockDim.x) pl[i] = pool[i];
        __syncthreads();
        for (int j = threadIdx.x; j < np_; j += blockDim.x) {
            const double *r = pl + (size_t)j * 6;
            pkv[j] = A::mul(A::mul(A::sub(r[3], r[0]), A::sub(r[4], r[1])), A::sub(r[5], r[2]));
        }
        __syncthreads();
    }
    // With a positive match threshold only OVERLAPPING pairs can match, and they are ~1 % of the P x M pairs: a
    // conservative fp32 test (pick bounds rounded outwards against pool bounds rounded outwards; "certainly disjoint on
    // some axis" implies the exact overlap is empty) runs at full fp32 rate, the few survivors go on a per-warp queue in
    // pool order and only they get the fp64 IoU.  A pick with no survivor has IoU 0 everywhere and cannot match.
    constexpr int NSLAB = 16;
    const int W16 = (np_ + 31) >> 5;
    const bool prune = pool_staged && p.match_thr > 0.0 &&
                       (size_t)np_ * 32 + sizeof(uint32_t) * 2 * NSLAB * (W16 + 1) + 64 <=
                           (size_t)(reinterpret_cast<unsigned char *>(sh.skey) - sm);
    float4 *plo = reinterpret_cast<float4 *>(sm);            // [np_] lower bounds rounded down (the NMS tables are idle here)
    float4 *phi = plo + np_;                                 // [np_] upper bounds rounded up
    // Slab masks: the pool's x and y extents are cut into 16 slabs each; xs[s] / ys[s] = bitmask of the pool boxes that
    // reach into slab s.  A pick only looks at (OR of its x slabs) & (OR of its y slabs) -- ~5 % of the pool -- instead
    // of testing all 512 boxes; the conservative fp32 test and the fp64 IoU then run on those few.
    uint32_t *xs = reinterpret_cast<uint32_t *>(phi + np_);
    uint32_t *ys = xs + NSLAB * W16;
    float *ext = reinterpret_cast<float *>(ys + NSLAB * W16);   // xmin, inv_x, ymin, inv_y
    if (prune) {
        for (int j = threadIdx.x; j < np_; j += blockDim.x) {
            const double *r = pl + (size_t)j * 6;
            plo[j] = make_float4(__double2float_rd(r[0]), __double2float_rd(r[1]), __double2float_rd(r[2]), __double2float_rd(pkv[j]) * (1.f - 1e-6f));   // .w: lower bound of the volume
            phi[j] = make_float4(__double2float_ru(r[3]), __double2float_ru(r[4]), __double2float_ru(r[5]), 0.f);
        }
        for (int i = threadIdx.x; i < 2 * NSLAB * W16; i += blockDim.x) xs[i] = 0u;
        __syncthreads();
        if (warp == 0) {
            float x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;
            for (int j = lane; j < np_; j += 32) { x0 = fminf(x0, plo[j].x); x1 = fmaxf(x1, phi[j].x); y0 = fminf(y0, plo[j].y); y1 = fmaxf(y1, phi[j].y); }
            for (int off = 16; off > 0; off >>= 1) {
                x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, off)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, off));
                y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, off)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, off));
            }
            if (lane == 0) { ext[0] = x0; ext[1] = x1 > x0 ? NSLAB / (x1 - x0) : 0.f; ext[2] = y0; ext[3] = y1 > y0 ? NSLAB / (y1 - y0) : 0.f; }
        }
        __syncthreads();
    }
    // slab of a coordinate: one monotone function for box and pick bounds alike, so overlapping intervals share a slab
    auto slab = [&](float v, float mn, float inv) { const float t = (v - mn) * inv; return t >= (float)(NSLAB - 1) ? NSLAB - 1 : (t > 0.f ? (int)t : 0); };
    if (prune) {
        const float xmn = ext[0], xin = ext[1], ymn = ext[2], yin = ext[3];
        for (int j = threadIdx.x; j < np_; j += blockDim.x) {
            const float4 a = plo[j], c = phi[j];
            const uint32_t bit = 1u << (j & 31);
            for (int sl = slab(a.x, xmn, xin); sl <= slab(c.x, xmn, xin); ++sl) atomicOr(&xs[sl * W16 + (j >> 5)], bit);
            for (int sl = slab(a.y, ymn, yin); sl <= slab(c.y, ymn, yin); ++sl) atomicOr(&ys[sl * W16 + (j >> 5)], bit);
        }
        __syncthreads();
    }
    NSTAMP(p.dbg, 2);
    // A group of GL lanes per surviving box: 8 lanes (two mask words each, four boxes per warp) when the pool fits 16 mask
    // words, else the whole warp.  The boxes' coordinates are staged in shared memory in pick order first, so that a
    // group does not start every box with a dependent global load.
    const int GL = (prune && W16 <= 16) ? 8 : 32;
    double *pk = reinterpret_cast<double *>(ext + 4);          // [npick][6]
    const bool picks_staged = prune && (size_t)(reinterpret_cast<unsigned char *>(pk + (size_t)npick * 6) - sm) <= (size_t)(reinterpret_cast<unsigned char *>(sh.skey) - sm);
    if (picks_staged) {
        for (int i = threadIdx.x; i < npick * 6; i += blockDim.x) { const int t = i / 6; pk[i] = boxes[(size_t)order[t] * 8 + (i - 6 * t)]; }
        __syncthreads();
    }
    const int ngroups = blockDim.x / GL, group = threadIdx.x / GL, gl = threadIdx.x & (GL - 1);
    for (int t0 = 0; t0 < npick; t0 += ngroups) {
        const int t = t0 + group;
        const bool act = t < npick;
        const double *bx = boxes + (size_t)order[act ? t : 0] * 8;
        const double *bc = picks_staged ? pk + (size_t)(act ? t : 0) * 6 : bx;
        const double b0 = bc[0], b1 = bc[1], b2 = bc[2], b3 = bc[3], b4 = bc[4], b5 = bc[5];
        const double qv = A::mul(A::mul(A::sub(b3, b0), A::sub(b4, b1)), A::sub(b5, b2));
        double bi = -INFINITY; int bj = 0x7fffffff;
        auto exact_pair = [&](int j) {
            const double *r = (pool_staged ? pl : pool) + (size_t)j * 6;
            const double kv = pool_staged ? pkv[j] : A::mul(A::mul(A::sub(r[3], r[0]), A::sub(r[4], r[1])), A::sub(r[5], r[2]));
            // compare-and-select min / max (the library's fmin / fmax spend ~10 instructions on NaN handling per call)
            auto dmin = [](double x, double y) { return x < y ? x : y; };
            auto dmax = [](double x, double y) { return x > y ? x : y; };
            const double e0 = dmax(A::sub(dmin(b3, r[3]), dmax(b0, r[0])), 0.0);
            const double e1 = dmax(A::sub(dmin(b4, r[4]), dmax(b1, r[1])), 0.0);
            const double e2 = dmax(A::sub(dmin(b5, r[5]), dmax(b2, r[2])), 0.0);
            double iou;
            if (e0 == 0.0 || e1 == 0.0 || e2 == 0.0) {
                const double den = A::add(A::add(qv, kv), 1e-5);   // inter = +0: (qv + kv - 0) + 1e-5
                iou = den > 0.0 ? 0.0 : A::div(0.0, den);
            } else {
                const double inter = A::mul(A::mul(e0, e1), e2);
                iou = A::div(inter, A::add(A::sub(A::add(qv, kv), inter), 1e-5));
            }
            if (iou > bi || (iou == bi && j < bj)) { bi = iou; bj = j; }  // np.argmax: first maximum (whatever order the candidates come in)
        };
        if (prune) {
            const float lx = __double2float_rd(b0), ly = __double2float_rd(b1), lz = __double2float_rd(b2);
            const float hx = __double2float_ru(b3), hy = __double2float_ru(b4), hz = __double2float_ru(b5);
            const float qv_lo = __double2float_rd(qv) * (1.f - 1e-6f), thr_lo = __double2float_rd(p.match_thr) * (1.f - 1e-6f);
            // candidate words from the slab masks (lane w of the group owns pool boxes 32w .. 32w+31), then the conservative fp32 test
            if (act) for (int w = gl; w < W16; w += GL) {   // (one word per lane unless the pool is longer than 32 * GL)
                uint32_t cx = 0u, cy = 0u;
                for (int sl = slab(lx, ext[0], ext[1]); sl <= slab(hx, ext[0], ext[1]); ++sl) cx |= xs[sl * W16 + w];
                for (int sl = slab(ly, ext[2], ext[3]); sl <= slab(hy, ext[2], ext[3]); ++sl) cy |= ys[sl * W16 + w];
                uint32_t m = cx & cy;
                while (m) {
                    const int bb = __ffs(m) - 1;
                    m &= m - 1;
                    const float4 a = plo[32 * w + bb], c = phi[32 * w + bb];
                    // fp32 UPPER bound of the IoU from outward-rounded bounds (overlap too large, volumes too small): a pair that
                    // cannot reach the match threshold can neither match nor be the arg-max of a box that does; the few
                    // survivors (spread over the group's lanes) go straight to the exact fp64 IoU
                    const float ex = fminf(hx, c.x) - fmaxf(lx, a.x), ey = fminf(hy, c.y) - fmaxf(ly, a.y), ez = fminf(hz, c.z) - fmaxf(lz, a.z);
                    if (ex > 0.f && ey > 0.f && ez > 0.f) {
                        const float ih = ex * ey * ez * (1.f + 1e-5f), den = qv_lo + a.w - ih;
                        if (!(den > 0.f) || ih * (1.f + 1e-5f) >= thr_lo * den) exact_pair(32 * w + bb);
                    }
                }
            }
        } else if (act) {
            for (int j = gl; j < np_; j += GL) exact_pair(j);
        }
        for (int off = GL / 2; off > 0; off >>= 1) {
            const double oi = __shfl_xor_sync(0xffffffffu, bi, off, GL);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, off, GL);
            if (oi > bi || (oi == bi && oj < bj)) { bi = oi; bj = oj; }
        }
        if (act && gl == 0 && bj != 0x7fffffff && !(bi < p.match_thr)) {
            const double sc = bx[6];
            // `box[-2] > tmp_score[index]` with tmp_score starting at 0: boxes arrive in pick order
            // (= descending score), so the earliest claimant with score > 0 wins and is never replaced.
            if (sc > 0.0) atomicMin(&best[bj], (unsigned int)t);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < np_; j += blockDim.x) {
        if (best[j] != 0xFFFFFFFFu) {
            const int t = (int)best[j];
            const double *bx = boxes + (size_t)order[t] * 8;
            olab[j] = bx[7];
            osc[j] = bx[6];
        }
    }
    __syncthreads();
    NSTAMP(p.dbg, 3);
    // 3. size-scored class-wise NMS over the labelled pool boxes (lift_boxes.py:165)
    PoolSrc psrc{pool, olab, osc, np_};
    nms_core(psrc, p.M, 3, true, false, p.size_thr, 1e-8, sh, nullptr);
    const int na = sh.misc[0];
    for (int pos = threadIdx.x; pos < na; pos += blockDim.x)
        if (sh.picked[pos]) p.out_keep[(size_t)s * p.M + sh.sidx[pos]] = 1;
    NSTAMP(p.dbg, 4);
    if (p.dbg && threadIdx.x == 0) p.dbg[(size_t)blockIdx.x * 16 + 5] = (unsigned long long)npick;
}

}  // namespace ovdet

using namespace ovdet;

extern "C" int ovdet_nms_f64(const double *boxes, const int32_t *counts, int S, int K, int ncols,
                             double thr, double vol_eps, unsigned flags,
                             uint8_t *keep, int32_t *pick_order, int32_t *npick, void *stream)
{
    OVDET_REQUIRE(S >= 0 && K >= 0, "negative size");
    if (S == 0 || K == 0) return OVDET_OK;
    OVDET_REQUIRE(boxes && keep, "null pointer");
    OVDET_REQUIRE(K <= NMS_MAXK, "K must be <= 1024");
    const int dims = (flags & OVDET_NMS_2D) ? 2 : 3;
    OVDET_REQUIRE(ncols >= 2 * dims + 1 + ((flags & OVDET_NMS_SAMECLS) ? 1 : 0), "ncols too small");
    NmsParams p{boxes, counts, S, K, ncols, thr, vol_eps, flags, keep, pick_order, npick};
    const size_t smem = nms_smem_bytes(K);
    OVDET_CUDA_TRY(ensure_dyn_smem(nms_kernel, smem));
    nms_kernel<<<S, K <= 128 ? 128 : NMS_NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return launch_ok("nms_kernel");
}

extern "C" int ovdet_parse_predictions_f32(const float *corners, const float *probs, const float *obj,
                                           const uint8_t *nonempty, int S, int K, int C,
                                           double nms_iou, float conf_thresh, unsigned flags,
                                           uint8_t *pred_mask, uint8_t *keep, int32_t *pred_cls, float *pred_cls_prob,
                                           void *stream)
{
    OVDET_REQUIRE(S >= 0 && K >= 0 && C > 0, "bad size");
    if (S == 0 || K == 0) return OVDET_OK;
    OVDET_REQUIRE(corners && probs && obj && pred_mask && keep && pred_cls && pred_cls_prob, "null pointer");
    OVDET_REQUIRE(K <= NMS_MAXK, "K must be <= 1024");
    ParseParams p{corners, probs, obj, nonempty, S, K, C, nms_iou, conf_thresh, flags, pred_mask, keep, pred_cls, pred_cls_prob, nullptr};
    { const char *e = getenv("OVDET_PARSE_DBG_PTR"); p.dbg = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0)) : nullptr; }
    const size_t smem = nms_smem_bytes(K) + sizeof(int) * (size_t)K;
    OVDET_CUDA_TRY(ensure_dyn_smem(parse_predictions_kernel, smem));
    parse_predictions_kernel<<<S, K <= 128 ? 128 : NMS_NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return launch_ok("parse_predictions_kernel");
}

extern "C" int ovdet_pseudo_filter_f64(const double *boxes, const double *pool, const int32_t *nboxes, const int32_t *npool,
                                       int S, int P, int M, double nms_thr, double match_thr, double size_nms_thr,
                                       uint8_t *nms1_keep, double *out_label, double *out_score, uint8_t *out_keep,
                                       void *stream)
{
    OVDET_REQUIRE(S >= 0 && P > 0 && M > 0, "bad size");
    if (S == 0) return OVDET_OK;
    OVDET_REQUIRE(boxes && pool && nms1_keep && out_label && out_score && out_keep, "null pointer");
    OVDET_REQUIRE(P <= NMS_MAXK && M <= NMS_MAXK, "P and M must be <= 1024");
    PseudoParams p{boxes, pool, nboxes, npool, S, P, M, P > M ? P : M, nms_thr, match_thr, size_nms_thr,
                   nms1_keep, out_label, out_score, out_keep, nullptr};
    { const char *e = getenv("OVDET_PSEUDO_DBG_PTR"); p.dbg = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0)) : nullptr; }
    const size_t smem = nms_smem_bytes(p.Kmax);
    OVDET_CUDA_TRY(ensure_dyn_smem(pseudo_filter_kernel, smem));
    pseudo_filter_kernel<<<S, NMS_NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return launch_ok("pseudo_filter_kernel");
}
