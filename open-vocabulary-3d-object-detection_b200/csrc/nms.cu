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
    const bool prune = pool_staged && p.match_thr > 0.0 &&
                       (size_t)np_ * 32 + (size_t)(blockDim.x >> 5) * np_ * 2 <= (size_t)(reinterpret_cast<unsigned char *>(sh.skey) - sm);
    float4 *plo = reinterpret_cast<float4 *>(sm);            // [np_] lower bounds rounded down (the NMS tables are idle here)
    float4 *phi = plo + np_;                                 // [np_] upper bounds rounded up
    unsigned short *wq = reinterpret_cast<unsigned short *>(phi + np_) + (size_t)warp * np_;   // this warp's candidate queue
    if (prune) {
        for (int j = threadIdx.x; j < np_; j += blockDim.x) {
            const double *r = pl + (size_t)j * 6;
            plo[j] = make_float4(__double2float_rd(r[0]), __double2float_rd(r[1]), __double2float_rd(r[2]), 0.f);
            phi[j] = make_float4(__double2float_ru(r[3]), __double2float_ru(r[4]), __double2float_ru(r[5]), 0.f);
        }
        __syncthreads();
    }
    for (int t = warp; t < npick; t += (blockDim.x >> 5)) {
        const double *bx = boxes + (size_t)order[t] * 8;
        const double b0 = bx[0], b1 = bx[1], b2 = bx[2], b3 = bx[3], b4 = bx[4], b5 = bx[5];
        const double qv = A::mul(A::mul(A::sub(b3, b0), A::sub(b4, b1)), A::sub(b5, b2));
        double bi = -INFINITY; int bj = 0x7fffffff;
        auto exact_pair = [&](int j) {
            const double *r = (pool_staged ? pl : pool) + (size_t)j * 6;
            const double kv = pool_staged ? pkv[j] : A::mul(A::mul(A::sub(r[3], r[0]), A::sub(r[4], r[1])), A::sub(r[5], r[2]));
            const double e0 = A::max(A::sub(A::min(b3, r[3]), A::max(b0, r[0])), 0.0);
            const double e1 = A::max(A::sub(A::min(b4, r[4]), A::max(b1, r[1])), 0.0);
            const double e2 = A::max(A::sub(A::min(b5, r[5]), A::max(b2, r[2])), 0.0);
            double iou;
            if (e0 == 0.0 || e1 == 0.0 || e2 == 0.0) {
                const double den = A::add(A::add(qv, kv), 1e-5);   // inter = +0: (qv + kv - 0) + 1e-5
                iou = den > 0.0 ? 0.0 : A::div(0.0, den);
            } else {
                const double inter = A::mul(A::mul(e0, e1), e2);
                iou = A::div(inter, A::add(A::sub(A::add(qv, kv), inter), 1e-5));
            }
            if (iou > bi) { bi = iou; bj = j; }  // np.argmax: first maximum
        };
        if (prune) {
            const float lx = __double2float_rd(b0), ly = __double2float_rd(b1), lz = __double2float_rd(b2);
            const float hx = __double2float_ru(b3), hy = __double2float_ru(b4), hz = __double2float_ru(b5);
            int nq = 0;
            for (int j0 = 0; j0 < np_; j0 += 32) {
                const int j = j0 + lane;
                bool maybe = false;
                if (j < np_) {
                    const float4 a = plo[j], c = phi[j];
                    maybe = !(hx <= a.x || c.x <= lx || hy <= a.y || c.y <= ly || hz <= a.z || c.z <= lz);
                }
                const unsigned m = __ballot_sync(0xffffffffu, maybe);
                if (maybe) wq[nq + __popc(m & ((1u << lane) - 1))] = (unsigned short)j;
                nq += __popc(m);
            }
            __syncwarp();
            for (int q = lane; q < nq; q += 32) exact_pair(wq[q]);   // ascending pool index within a lane: first maximum kept
            __syncwarp();
        } else {
            for (int j = lane; j < np_; j += 32) exact_pair(j);
        }
        for (int off = 16; off > 0; off >>= 1) {
            const double oi = __shfl_xor_sync(0xffffffffu, bi, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
            if (oi > bi || (oi == bi && oj < bj)) { bi = oi; bj = oj; }
        }
        if (lane == 0 && bj != 0x7fffffff && !(bi < p.match_thr)) {
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
    // 3. size-scored class-wise NMS over the labelled pool boxes (lift_boxes.py:165)
    PoolSrc psrc{pool, olab, osc, np_};
    nms_core(psrc, p.M, 3, true, false, p.size_thr, 1e-8, sh, nullptr);
    const int na = sh.misc[0];
    for (int pos = threadIdx.x; pos < na; pos += blockDim.x)
        if (sh.picked[pos]) p.out_keep[(size_t)s * p.M + sh.sidx[pos]] = 1;
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
    OVDET_CUDA_TRY(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    OVDET_CUDA_TRY(cudaFuncSetAttribute(parse_predictions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
                   nms1_keep, out_label, out_score, out_keep};
    const size_t smem = nms_smem_bytes(p.Kmax);
    OVDET_CUDA_TRY(cudaFuncSetAttribute(pseudo_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pseudo_filter_kernel<<<S, NMS_NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return launch_ok("pseudo_filter_kernel");
}
