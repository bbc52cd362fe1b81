// eval.cu -- AP evaluation path for sm_100a.
//
//   ovdet_box3d_iou_f64   exact pairwise IoU, box3d_iou (utils/box_util.py:116-141)
//   ovdet_ap_match        per-scene det x GT exact IoU, argmax and first-claim TP
//                         flags (utils/eval_det.py:117-140) for all classes and
//                         all thresholds at once; emits class-major records
//   ovdet_ap_reduce       per-class segmented LSD radix sort by descending score
//                         + scans: cumulative TP/FP, precision/recall, VOC AP
//                         (utils/eval_det.py:108-111, :143-153, voc_ap :23-54)
//
// The greedy loop of eval_det_cls only couples detections of one scene and one
// class, and a detection's best GT (jmax) does not depend on earlier claims, so
//   TP(d) <=> ovmax(d) > thr  and  d is the highest-scoring detection among
//             those with the same jmax and ovmax > thr
// which is evaluated in parallel with a 64-bit atomicMax per GT -- no sequential
// pass.  Equality with the reference needs tie-free scores (numpy's argsort of
// -confidence is not stable: documented).
#include <math.h>
#include <stdlib.h>

#include "ap_match.cuh"
#include "nms_core.cuh"

namespace ovdet {

// ---------------------------------------------------------------- exact IoU
// box3d_iou on fp32 corners cast to fp64 (eval_det.py:120,122).  The hull area of
// box_util.py:96 (Qhull) is the shoelace area of the clipped polygon (convex).
template <int STRIDE>
__device__ __forceinline__ double exact_iou(const float *c1, const float *c2, V2<double> *bufA, V2<double> *bufB,
                                            bool want2d, double *iou2d)
{
    using A = Ar<double>;
    const double ymax = A::min((double)c1[1], (double)c2[1]);
    const double ymin = A::max((double)c1[13], (double)c2[13]);
    const double h = A::max(0.0, A::sub(ymax, ymin));
    double vol[2];
    const float *cc[2] = {c1, c2};
    const int pa[3] = {0, 1, 0}, pb[3] = {1, 2, 4};
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        double e[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const double dx = A::sub((double)cc[w][3 * pa[t]], (double)cc[w][3 * pb[t]]);
            const double dy = A::sub((double)cc[w][3 * pa[t] + 1], (double)cc[w][3 * pb[t] + 1]);
            const double dz = A::sub((double)cc[w][3 * pa[t] + 2], (double)cc[w][3 * pb[t] + 2]);
            e[t] = A::sqrt(A::add(A::add(A::mul(dx, dx), A::mul(dy, dy)), A::mul(dz, dz)));
        }
        vol[w] = A::mul(A::mul(e[0], e[1]), e[2]);
    }
    double ia = 0.0;
    // disjoint BEV bounding rectangles -> the clip is empty -> area 0 exactly; skip it
    bool bev = true;
#pragma unroll
    for (int a = 0; a < 3; a += 2) {
        const float lo1 = fminf(fminf(c1[a], c1[3 + a]), fminf(c1[6 + a], c1[9 + a])), hi1 = fmaxf(fmaxf(c1[a], c1[3 + a]), fmaxf(c1[6 + a], c1[9 + a]));
        const float lo2 = fminf(fminf(c2[a], c2[3 + a]), fminf(c2[6 + a], c2[9 + a])), hi2 = fmaxf(fmaxf(c2[a], c2[3 + a]), fmaxf(c2[6 + a], c2[9 + a]));
        if (hi1 < lo2 || hi2 < lo1) bev = false;
    }
    if ((h > 0.0 && bev) || want2d) {
        double s[8], c[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s[2 * i] = (double)c1[3 * (3 - i)]; s[2 * i + 1] = (double)c1[3 * (3 - i) + 2];
            c[2 * i] = (double)c2[3 * (3 - i)]; c[2 * i + 1] = (double)c2[3 * (3 - i) + 2];
        }
        const int n = sh_clip_quads<double, STRIDE>(s, c, bufA, bufB);
        ia = area_f64<STRIDE>(bufB, n);
        if (want2d) {
            double a[2];
            const double *rr[2] = {s, c};
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                double d1 = 0.0, d2 = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = (i + 3) & 3;
                    d1 = A::add(d1, A::mul(rr[w][2 * i], rr[w][2 * j + 1]));
                    d2 = A::add(d2, A::mul(rr[w][2 * i + 1], rr[w][2 * j]));
                }
                a[w] = A::mul(0.5, fabs(A::sub(d1, d2)));
            }
            *iou2d = A::div(ia, A::sub(A::add(a[0], a[1]), ia));
        }
    }
    const double iv = A::mul(ia, h);
    return A::div(iv, A::sub(A::add(vol[0], vol[1]), iv));
}

constexpr int EV_NT = 256;

// Axis-aligned IoU of the tools' evaluation (3DOVDet_tools/utils/evaluation/box_util.py:287-309 calc_iou, clamped to
// [0, 1] by get_iou, eval_det.py:63-77): boxes are (centre, lengths), fp64, numpy's operation order.
struct AabbParams { const double *dets, *gts; const int32_t *nd, *ng; int S, D, G; double *out; long long total; };

__global__ void __launch_bounds__(EV_NT) aabb_iou_kernel(AabbParams p)
{
    using A = Ar<double>;
    for (long long idx = (long long)blockIdx.x * EV_NT + threadIdx.x; idx < p.total; idx += (long long)gridDim.x * EV_NT) {
        const int g = (int)(idx % p.G);
        const long long sd = idx / p.G;
        const int d = (int)(sd % p.D), s = (int)(sd / p.D);
        double r = 0.0;
        if ((!p.nd || d < p.nd[s]) && (!p.ng || g < p.ng[s])) {
            const double *a = p.dets + sd * 6, *b = p.gts + ((long long)s * p.G + g) * 6;
            double e[3];
            bool pos = true;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double ha = A::div(a[3 + k], 2.0), hb = A::div(b[3 + k], 2.0);
                const double min_max = A::min(A::add(a[k], ha), A::add(b[k], hb)), max_min = A::max(A::sub(a[k], ha), A::sub(b[k], hb));
                pos = pos && (min_max > max_min);
                e[k] = A::sub(min_max, max_min);
            }
            if (pos) {
                const double inter = A::mul(A::mul(e[0], e[1]), e[2]);
                const double va = A::mul(A::mul(a[3], a[4]), a[5]), vb = A::mul(A::mul(b[3], b[4]), b[5]);
                r = A::div(inter, A::sub(A::add(va, vb), inter));
                if (r < 0.0) r = 0.0;
                if (r > 1.0) r = 1.0;
            }
        }
        p.out[idx] = r;
    }
}

struct IouParams {
    const float *dets, *gts; const int32_t *nd, *ng;
    int S, D, G; double *out, *out2d; long long total;
};

__global__ void __launch_bounds__(EV_NT) box3d_iou_kernel(IouParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    V2<double> *scratch = reinterpret_cast<V2<double> *>(sm);
    V2<double> *bufA = scratch + threadIdx.x, *bufB = scratch + SH_MAXV * EV_NT + threadIdx.x;
    for (long long idx = (long long)blockIdx.x * EV_NT + threadIdx.x; idx < p.total; idx += (long long)gridDim.x * EV_NT) {
        const int g = (int)(idx % p.G);
        const long long sd = idx / p.G;
        const int d = (int)(sd % p.D);
        const int s = (int)(sd / p.D);
        double r = 0.0, r2 = 0.0;
        const bool live = (!p.nd || d < p.nd[s]) && (!p.ng || g < p.ng[s]);
        if (live) {
            float c1[24], c2[24];
            const float *a = p.dets + sd * 24, *b = p.gts + ((long long)s * p.G + g) * 24;
#pragma unroll
            for (int i = 0; i < 24; ++i) { c1[i] = __ldg(a + i); c2[i] = __ldg(b + i); }
            r = exact_iou<EV_NT>(c1, c2, bufA, bufB, p.out2d != nullptr, &r2);
        }
        p.out[idx] = r;
        if (p.out2d) p.out2d[idx] = r2;
    }
}

__host__ __device__ inline size_t am_lower_bytes(int K, int G)
{   // feature records, clip scratch, candidate IoUs: idle while the class-probability tile is in use
    return sizeof(AmBox) * ((size_t)K + G) + sizeof(V2<double>) * 2 * SH_MAXV * AM_CLIP + sizeof(double) * AM_QCAP;
}
__host__ __device__ inline size_t am_upper_bytes(int K, int G, int nthr)
{   // claim table, compaction tables, candidate queue: live from the first phase on
    return sizeof(unsigned long long) * (size_t)G * nthr + sizeof(int) * ((size_t)K + 2 * (size_t)G) + (sizeof(unsigned) + sizeof(float)) * AM_QCAP + 64;
}
__host__ __device__ inline size_t am_smem_bytes(int K, int G, int nthr) { return am_lower_bytes(K, G) + am_upper_bytes(K, G, nthr); }

struct MatchParams {
    const float *corners, *probs, *obj; const uint8_t *keep; const int32_t *det_cls;
    const float *gt_corners; const int64_t *gt_labels; const uint8_t *gt_present;
    const float *gt_present_f32;   // the reference's float mask (datasets/*.py) read directly when gt_present is null
    int S, K, G, C, nthr; double thr[8];
    double *iou_ws; float *rec_score; uint8_t *rec_tp; unsigned long long *npos;
    const double *iou_given;   // [S,K,G] precomputed IoU matrix (original det / GT indices): skip the geometry (ovdet_ap_match_iou)
    // optional TP list (csrc/ap_compact.cu): every record with a TP bit is appended as (score key, bits)
    uint32_t *tp_key; uint8_t *tp_bits; int *tp_cnt; int tp_cap;
    unsigned long long *dbg;   // optional [S][8] globaltimer stamps of thread 0 (OVDET_APMATCH_DBG_PTR; null in production)
};
#define AMSTAMP(i) do { if (p.dbg && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.dbg[(size_t)blockIdx.x * 8 + (i)] = t_; } } while (0)

// One scene's inputs as the matching body sees them (generic pointers: global memory in ap_match_kernel, shared memory
// for what the fused front end has just computed).
struct AmScene {
    const float *probs;        // [K,C] global, or null (single-class layouts)
    const float *ptile;        // the same tile already staged in shared memory (row pitch ptile_pitch), or null
    int ptile_pitch;
    const float *score1;       // [K] objectness (per-class layout: multiplied with the class probability) or the det's score
    const uint8_t *keep;       // [K]
    const int32_t *det_cls;    // [K] or null
};

// sm_lo: am_lower_bytes (aliases the tile the body stages itself; sc.ptile may overlap its tail beyond the feature records --
// the tile is only read in the record phase, before anything but the feature records is written), sm_hi: am_upper_bytes.
__device__ __forceinline__ void am_scene_body(const MatchParams &p, const AmScene &sc, unsigned char *sm_lo, unsigned char *sm_hi, const int s)
{
    AmBox *dbox = reinterpret_cast<AmBox *>(sm_lo);
    AmBox *gbox = dbox + p.K;
    V2<double> *scratch = reinterpret_cast<V2<double> *>(gbox + p.G);
    double *qiou = reinterpret_cast<double *>(scratch + 2 * SH_MAXV * AM_CLIP);   // IoU of the queued (surviving) pairs
    unsigned long long *best = reinterpret_cast<unsigned long long *>(sm_hi);
    int *kd = reinterpret_cast<int *>(best + (size_t)p.G * p.nthr);
    int *gl = kd + p.K;
    int *glab = gl + p.G;
    unsigned *queue = reinterpret_cast<unsigned *>(glab + p.G);
    float *qscore = reinterpret_cast<float *>(queue + AM_QCAP);   // score of (det, class of the GT) per candidate: loaded early, used by the claims
    __shared__ int nk_s, ng_s, qn_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t N = (size_t)p.S * p.K;
    const uint8_t *keep = sc.keep;
    const float *corners = p.corners + (size_t)s * p.K * 24;
    const float *gt_corners = p.gt_corners + (size_t)s * p.G * 24;
    AMSTAMP(0);

    // ordered compaction of kept detections (warp 0) / present GT (warp 1), ballot + popc
    if (warp == 0) {
        int n = 0;
        for (int base = 0; base < p.K; base += 32) {
            const int k = base + lane;
            const bool f = k < p.K && keep[k];
            const unsigned m = __ballot_sync(0xffffffffu, f);
            if (f) kd[n + __popc(m & ((1u << lane) - 1))] = k;
            n += __popc(m);
        }
        if (lane == 0) nk_s = n;
    } else if (warp == 1) {
        int n = 0;
        for (int base = 0; base < p.G; base += 32) {
            const int g = base + lane;
            const bool f = g < p.G && (p.gt_present ? p.gt_present[(size_t)s * p.G + g] != 0 : p.gt_present_f32[(size_t)s * p.G + g] != 0.f);
            const unsigned m = __ballot_sync(0xffffffffu, f);
            if (f) {
                const int pos = n + __popc(m & ((1u << lane) - 1));
                gl[pos] = g;
                const long long lab = p.gt_labels[(size_t)s * p.G + g];
                glab[pos] = (int)lab;
                if (lab >= 0 && lab < p.C) atomicAdd(&p.npos[lab], 1ull);
            }
            n += __popc(m);
        }
        if (lane == 0) ng_s = n;
    }
    for (int i = tid; i < p.G * p.nthr; i += AM_NT) best[i] = 0ull;
    __syncthreads();
    const int nk = nk_s, ng = ng_s;
    AMSTAMP(1);

    // records: score of every (class, slot), tp = 0.  The scene's [K, C] probability tile goes through shared memory
    // (staged here over the idle feature / clip / candidate areas unless the caller already holds it) with independent
    // 16-byte loads, so that no thread walks C dependent global loads; the stores are coalesced in k for each class.
    {
        const float *ptile = sc.ptile;
        int pitch = sc.ptile_pitch;
        const size_t kc = (size_t)p.K * p.C;
        const float *pg = sc.probs;
        if (pg && !ptile && kc * sizeof(float) <= am_lower_bytes(p.K, p.G)) {
            float *t = reinterpret_cast<float *>(sm_lo);
            if ((reinterpret_cast<uintptr_t>(pg) & 15) == 0 && (kc & 3) == 0) {
                for (int i = tid; i < (int)(kc >> 2); i += AM_NT) reinterpret_cast<float4 *>(t)[i] = __ldg(reinterpret_cast<const float4 *>(pg) + i);
            } else {
                for (int i = tid; i < (int)kc; i += AM_NT) t[i] = __ldg(pg + i);
            }
            __syncthreads();
            ptile = t; pitch = p.C;
        }
        for (int k = tid; k < p.K; k += AM_NT) {
            const size_t slot = (size_t)s * p.K + k;
            const bool kept = keep[k];
            const float ob = kept ? sc.score1[k] : 0.f;
            const int dc = (kept && sc.det_cls) ? sc.det_cls[k] : -1;
            for (int c = 0; c < p.C; ++c) {
                float v = -INFINITY;
                if (kept) {
                    if (sc.det_cls) { if (dc == c) v = ob; }
                    else v = __fmul_rn(ptile ? ptile[(size_t)k * pitch + c] : __ldg(pg + (size_t)k * p.C + c), ob);
                }
                p.rec_score[(size_t)c * N + slot] = v;
                if (p.rec_tp) p.rec_tp[(size_t)c * N + slot] = 0;
            }
        }
    }
    AMSTAMP(2);
    if (ng == 0 || nk == 0) return;   // no claims possible (uniform)
    __syncthreads();                  // a tile staged here aliases the feature records written next

    const bool given = p.iou_given != nullptr;
    // stage features
    if (!given) for (int i = tid; i < nk + ng; i += AM_NT) {
        float c[24];
        if (i < nk) { am_load_box(corners + (size_t)kd[i] * 24, c); am_features(c, dbox[i]); }
        else { am_load_box(gt_corners + (size_t)gl[i - nk] * 24, c); am_features(c, gbox[i - nk]); }
    }
    double thr_min = p.thr[0];
    for (int t = 1; t < p.nthr; ++t) thr_min = fmin(thr_min, p.thr[t]);
    const int npair = nk * ng;
    if (tid == 0) { qn_s = 0; }
    __syncthreads();
    AMSTAMP(3);

    // score of (det k, class c) exactly as the record loop wrote it
    auto det_score = [&](int k, int c) -> float {
        if (sc.det_cls) return sc.score1[k];
        return __fmul_rn(__ldg(sc.probs + (size_t)k * p.C + c), sc.score1[k]);   // global: a staged tile is dead by now
    };
    auto emit_tp = [&](int k, int c, unsigned char tp, float score) {
        const size_t slot = (size_t)s * p.K + k;
        if (p.rec_tp) p.rec_tp[(size_t)c * N + slot] = tp;
        if (p.tp_key) {
            const int at = atomicAdd(&p.tp_cnt[c], 1);
            if (at < p.tp_cap) { p.tp_key[(size_t)c * p.tp_cap + at] = score_key(score); p.tp_bits[(size_t)c * p.tp_cap + at] = tp; }
        }
    };

    // ---- sparse mode (thr_min >= 0): only pairs whose IoU could exceed the smallest threshold are clipped.
    // Exact rejects: no height overlap or disjoint BEV rectangles (IoU = 0); conservative reject: the IoU is at most
    // ub = I/(V1+V2-I) with I = (overlap of the BEV bounding rectangles) x (height overlap) >= the true intersection;
    // a pair with ub(1+1e-6) < thr_min can neither be a TP nor change which GT is a detection's arg-max among those
    // above the threshold, so its IoU is never needed.  Survivors (with their IoU) form a short list in shared memory.
    bool dense = given || !(thr_min >= 0.0);
    if (!dense) {
        for (int base = 0; base < npair; base += AM_NT) {
            const int pi = base + tid;
            bool need = false;
            if (pi < npair) {
                const int i = pi / ng, j = pi - i * ng;
                const AmBox &a = dbox[i], &b = gbox[j];
                const float hh = fminf(a.ytop, b.ytop) - fmaxf(a.ybot, b.ybot);
                const float ox = fminf(a.hix, b.hix) - fmaxf(a.lox, b.lox), oz = fminf(a.hiz, b.hiz) - fmaxf(a.loz, b.loz);
                if (hh > 0.f && ox >= 0.f && oz >= 0.f) {
                    const double I = (double)ox * (double)oz * (double)hh * (1.0 + 1e-6);   // fp32 differences: widen
                    const double den = a.vol + b.vol - I;
                    need = !(den > 0.0) || I * (1.0 + 1e-6) >= thr_min * den;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (m) {
                int bp = 0;
                if (lane == 0) bp = atomicAdd(&qn_s, __popc(m));
                bp = __shfl_sync(0xffffffffu, bp, 0);
                const int q = bp + __popc(m & ((1u << lane) - 1));
                if (need && q < AM_QCAP) {   // K, G <= 32767
                    const int i = pi / ng, j = pi - i * ng;
                    queue[q] = ((unsigned)i << 16) | (unsigned)j;
                    const int c = glab[j];
                    qscore[q] = (c >= 0 && c < p.C) ? det_score(kd[i], c) : 0.f;   // the loads fly while the clipper runs
                }
            }
        }
        __syncthreads();
        if (qn_s > AM_QCAP) dense = true;   // uniform: fall back to the dense matrix in the caller's workspace
    }
    if (!dense) {
        const int qn = qn_s;
        constexpr int NGROUP = AM_NT / 8;
        if (qn <= 4 * NGROUP) {
            const int gli = lane & 7, gshift = lane & 24, group = tid >> 3;
            V2<double> *gbuf = scratch + group * 8;
            for (int base = 0; base < qn; base += NGROUP) {
                if (base + warp * 4 >= qn) break;   // warp-uniform
                const int qi = base + group;
                const bool act = qi < qn;
                int i = 0, j = 0;
                if (act) { const unsigned e = queue[qi]; i = (int)(e >> 16); j = (int)(e & 0xffffu); }
                const AmBox &a = dbox[i], &b = gbox[j];
                double cl[8];
#pragma unroll
                for (int t = 0; t < 4; ++t) { cl[2 * t] = (double)b.qx[t]; cl[2 * t + 1] = (double)b.qz[t]; }
                double vx = (double)a.qx[gli & 3], vy = (double)a.qz[gli & 3];
                const int n = coop_clip_quads<double>(cl, vx, vy, act ? 4 : 0, gli, gshift, gbuf);
                const double ia = coop_area_f64(vx, vy, n, gli);
                if (act && gli == 0) qiou[qi] = am_finish_iou(ia, a, b);
            }
        } else if (tid < AM_CLIP) {
            V2<double> *bufA = scratch + tid, *bufB = scratch + SH_MAXV * AM_CLIP + tid;
            for (int qi = tid; qi < qn; qi += AM_CLIP) {
                const unsigned e = queue[qi];
                const int i = (int)(e >> 16), j = (int)(e & 0xffffu);
                const AmBox &a = dbox[i], &b = gbox[j];
                double sq[8], cq[8];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    sq[2 * t] = (double)a.qx[t]; sq[2 * t + 1] = (double)a.qz[t];
                    cq[2 * t] = (double)b.qx[t]; cq[2 * t + 1] = (double)b.qz[t];
                }
                const int n = sh_clip_quads<double, AM_CLIP>(sq, cq, bufA, bufB);
                qiou[qi] = am_finish_iou(area_f64<AM_CLIP>(bufB, n), a, b);
            }
        }
        __syncthreads();
        AMSTAMP(4);
        if (p.dbg && threadIdx.x == 0) p.dbg[(size_t)blockIdx.x * 8 + 6] = (unsigned long long)qn | ((unsigned long long)nk << 16) | ((unsigned long long)ng << 32);
        // pass 1 (claims) / pass 2 (TP iff the claim is held): thread per survivor
        for (int pass = 0; pass < 2; ++pass) {
            for (int q = tid; q < qn; q += AM_NT) {
                const double v = qiou[q];
                if (!(v > thr_min)) continue;
                const unsigned e = queue[q];
                const int i = (int)(e >> 16), j = (int)(e & 0xffffu);
                const int c = glab[j];
                const int k = kd[i];
                if (c < 0 || c >= p.C || (sc.det_cls && sc.det_cls[k] != c)) continue;
                bool first_max = true;   // jmax of (det, class c): first GT of the class attaining the maximum (eval_det.py:121-126)
                for (int q2 = 0; q2 < qn; ++q2) {
                    const unsigned e2 = queue[q2];
                    if ((int)(e2 >> 16) != i) continue;
                    const int j2 = (int)(e2 & 0xffffu);
                    if (glab[j2] != c) continue;
                    const double v2 = qiou[q2];
                    if (v2 > v || (v2 == v && j2 < j)) { first_max = false; break; }
                }
                if (!first_max) continue;
                if (pass == 0) {
                    // non-negative fp32 scores order like their bit patterns; lower det index wins ties
                    const unsigned long long key = ((unsigned long long)__float_as_uint(qscore[q]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
                    for (int t = 0; t < p.nthr; ++t)
                        if (v > p.thr[t]) atomicMax(&best[(size_t)j * p.nthr + t], key);
                } else {
                    unsigned char tp = 0;
                    for (int t = 0; t < p.nthr; ++t)
                        if (v > p.thr[t]) {
                            const unsigned long long w = best[(size_t)j * p.nthr + t];
                            if ((unsigned)(0xFFFFFFFFu - (unsigned)(w & 0xFFFFFFFFull)) == (unsigned)i) tp |= (unsigned char)(1u << t);
                        }
                    if (tp) emit_tp(k, c, tp, qscore[q]);
                }
            }
            __syncthreads();
        }
        AMSTAMP(5);
        return;
    }

    // ---- dense mode (a negative threshold, or more than AM_QCAP candidate pairs): the full IoU matrix in the caller's
    // workspace, exact rejects only, pairs processed in slabs of AM_QCAP
    double *iou = given ? nullptr : p.iou_ws + (size_t)s * p.K * p.G;
    const int ldi = p.G;
    for (int slab = 0; !given && slab < npair; slab += AM_QCAP) {
        __syncthreads();
        if (tid == 0) qn_s = 0;
        __syncthreads();
        const int slab_end = min(npair, slab + AM_QCAP);
        for (int base = slab; base < slab_end; base += AM_NT) {
            const int pi = base + tid;
            bool need = false;
            if (pi < slab_end) {
                const int i = pi / ng, j = pi - i * ng;
                const AmBox &a = dbox[i], &b = gbox[j];
                need = fminf(a.ytop, b.ytop) > fmaxf(a.ybot, b.ybot) &&
                       !(a.hix < b.lox || b.hix < a.lox) && !(a.hiz < b.loz || b.hiz < a.loz);
                if (!need) iou[(size_t)i * ldi + j] = 0.0;
            }
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (m) {
                int bp = 0;
                if (lane == 0) bp = atomicAdd(&qn_s, __popc(m));
                bp = __shfl_sync(0xffffffffu, bp, 0);
                if (need) queue[bp + __popc(m & ((1u << lane) - 1))] = (unsigned)pi;
            }
        }
        __syncthreads();
        const int qn = qn_s;
        if (tid < AM_CLIP) {
            V2<double> *bufA = scratch + tid, *bufB = scratch + SH_MAXV * AM_CLIP + tid;
            for (int qi = tid; qi < qn; qi += AM_CLIP) {
                const int pi = (int)queue[qi];
                const int i = pi / ng, j = pi - i * ng;
                const AmBox &a = dbox[i], &b = gbox[j];
                double sq[8], cq[8];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    sq[2 * t] = (double)a.qx[t]; sq[2 * t + 1] = (double)a.qz[t];
                    cq[2 * t] = (double)b.qx[t]; cq[2 * t + 1] = (double)b.qz[t];
                }
                const int n = sh_clip_quads<double, AM_CLIP>(sq, cq, bufA, bufB);
                iou[(size_t)i * ldi + j] = am_finish_iou(area_f64<AM_CLIP>(bufB, n), a, b);
            }
        }
    }
    __threadfence_block();
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = tid; i < nk; i += AM_NT) {
            const int k = kd[i];
            const int dc = sc.det_cls ? sc.det_cls[k] : -1;
            // a given matrix is indexed by the ORIGINAL detection / GT positions, the one computed here by the compacted ones
            const double *row = given ? p.iou_given + ((size_t)s * p.K + k) * p.G : iou + (size_t)i * ldi;
            auto col = [&](int jj) { return given ? gl[jj] : jj; };
            for (int j = 0; j < ng; ++j) {
                const double v = row[col(j)];
                if (!(v > thr_min)) continue;
                const int c = glab[j];
                if (c < 0 || c >= p.C || (sc.det_cls && dc != c)) continue;
                bool first_max = true;
                for (int j2 = 0; j2 < ng; ++j2)
                    if (glab[j2] == c) { const double v2 = row[col(j2)]; if (v2 > v || (v2 == v && j2 < j)) { first_max = false; break; } }
                if (!first_max) continue;
                if (pass == 0) {
                    const unsigned long long key = ((unsigned long long)__float_as_uint(det_score(k, c)) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
                    for (int t = 0; t < p.nthr; ++t)
                        if (v > p.thr[t]) atomicMax(&best[(size_t)j * p.nthr + t], key);
                } else {
                    unsigned char tp = 0;
                    for (int t = 0; t < p.nthr; ++t)
                        if (v > p.thr[t]) {
                            const unsigned long long w = best[(size_t)j * p.nthr + t];
                            if ((unsigned)(0xFFFFFFFFu - (unsigned)(w & 0xFFFFFFFFull)) == (unsigned)i) tp |= (unsigned char)(1u << t);
                        }
                    if (tp) emit_tp(k, c, tp, det_score(k, c));
                }
            }
        }
        __syncthreads();
    }
    AMSTAMP(5);
}

__global__ void __launch_bounds__(AM_NT, 8) ap_match_kernel(MatchParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int s = blockIdx.x;
    AmScene sc;
    sc.probs = p.probs ? p.probs + (size_t)s * p.K * p.C : nullptr;
    sc.ptile = nullptr; sc.ptile_pitch = 0;
    sc.score1 = p.obj + (size_t)s * p.K;
    sc.keep = p.keep + (size_t)s * p.K;
    sc.det_cls = p.det_cls ? p.det_cls + (size_t)s * p.K : nullptr;
    am_scene_body(p, sc, sm, sm + am_lower_bytes(p.K, p.G), s);
}

int front1_launch(const float *corners, const float *probs, const float *obj, const uint8_t *nonempty,
                  const float *gt_corners, const int64_t *gt_labels, const void *gt_present,
                  int S, int K, int G, int C, double nms_iou, float conf_thresh, unsigned flags,
                  const double *thr, int nthr, double *iou_ws, float *rec_score, uint8_t *rec_tp, int64_t *npos,
                  uint32_t *tp_key, uint8_t *tp_bits, int32_t *tp_cnt, int tp_cap, uint8_t *keep_out, void *stream);

// ---------------------------------------------------------------- fused AP front end
// parse_predictions (argmax, AABB, NMS variant, confidence gate; utils/ap_calculator.py:39-238) and the AP matching of
// the same scene (utils/eval_det.py:117-140) in ONE 128-thread CTA: corners and the class-probability tile are read from
// HBM once, the keep mask / predicted class / class confidence never leave shared memory, no rec_tp stream is written
// (true positives go straight on the per-class TP lists of the compact reducer).
//   shared memory: [keep u8 K | cls i32 K | clsp f32 K] [union: NMS tables + probability tile  |  match lower area]
//                  [match upper area]
struct FrontParams {
    MatchParams m;
    const uint8_t *nonempty; double nms_iou; float conf; unsigned flags;   // parse_predictions' knobs (OVDET_NMS_*, OVDET_PARSE_NO_NMS, OVDET_FRONT_*)
    uint8_t *keep_out;   // optional [S,K]
    unsigned long long *dbg;   // optional [S][16] globaltimer stamps (OVDET_APFRONT_DBG_PTR; null in production)
};

__host__ __device__ inline size_t front_tile_bytes(int K, int C) { return sizeof(float) * (size_t)K * (C + 1); }
__host__ __device__ inline size_t front_persist_bytes(int K) { return ((size_t)K * 9 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t front_union_bytes(int K, int G, int C, bool tile)
{
    const size_t a = nms_smem_bytes(K) + (tile ? front_tile_bytes(K, C) : 0), b = am_lower_bytes(K, G);
    return ((a > b ? a : b) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t front_smem_bytes(int K, int G, int C, int nthr, bool tile)
{
    return front_persist_bytes(K) + front_union_bytes(K, G, C, tile) + am_upper_bytes(K, G, nthr);
}

template <bool TILE>
__global__ void __launch_bounds__(AM_NT, 7) ap_front_kernel(FrontParams fp)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const MatchParams &p = fp.m;
    const int s = blockIdx.x, K = p.K, C = p.C, tid = threadIdx.x;
    int *cls_sm = reinterpret_cast<int *>(sm);
    float *clsp_sm = reinterpret_cast<float *>(cls_sm + K);
    uint8_t *keep_sm = reinterpret_cast<uint8_t *>(clsp_sm + K);
    unsigned char *un = sm + front_persist_bytes(K);
    unsigned char *hi = un + front_union_bytes(K, p.G, C, TILE);
    NmsSmem sh = nms_carve(un, K);
    float *ptile = TILE ? reinterpret_cast<float *>(un + nms_smem_bytes(K)) : nullptr;
    const int pitch = C + 1;
    const float *probs = p.probs + (size_t)s * K * C;
    const float *obj = p.obj + (size_t)s * K;
    const uint8_t *ne = fp.nonempty ? fp.nonempty + (size_t)s * K : nullptr;

    NSTAMP(fp.dbg, 0);
    // ---- argmax / max class probability (ap_calculator.py:59-61; np.argmax = first maximum)
    if (TILE) {
        const size_t kc = (size_t)K * C;
        const float inv_c = 1.f / (float)C;
        auto rowcol = [&](int e, int &r, int &c) {   // e / C without an integer divide (e < 2^21: the float estimate is off by at most one)
            r = (int)(((float)e + 0.5f) * inv_c);
            c = e - r * C;
            if (c < 0) { --r; c += C; } else if (c >= C) { ++r; c -= C; }
        };
        if ((reinterpret_cast<uintptr_t>(probs) & 15) == 0 && (kc & 3) == 0) {
            for (int i = tid; i < (int)(kc >> 2); i += AM_NT) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(probs) + i);
                const float vv[4] = {v.x, v.y, v.z, v.w};
                int r, c;
                rowcol(4 * i, r, c);
#pragma unroll
                for (int q = 0; q < 4; ++q) { ptile[r * pitch + c] = vv[q]; if (++c == C) { c = 0; ++r; } }
            }
        } else {
            for (int e = tid; e < (int)kc; e += AM_NT) { int r, c; rowcol(e, r, c); ptile[r * pitch + c] = __ldg(probs + e); }
        }
        __syncthreads();
    }
    NSTAMP(fp.dbg, 1);
    for (int k = tid; k < K; k += AM_NT) {
        const float *row = TILE ? ptile + (size_t)k * pitch : nullptr;
        float best = TILE ? row[0] : __ldg(probs + (size_t)k * C);
        int bi = 0;
        for (int c = 1; c < C; ++c) { const float v = TILE ? row[c] : __ldg(probs + (size_t)k * C + c); if (v > best) { best = v; bi = c; } }
        cls_sm[k] = bi;
        clsp_sm[k] = best;
        keep_sm[k] = 0;
    }
    __syncthreads();
    NSTAMP(fp.dbg, 10);
    // ---- NMS variant + confidence gate -> keep (ap_calculator.py:86-189, :206-207)
    if (fp.flags & OVDET_PARSE_NO_NMS) {
        for (int k = tid; k < K; k += AM_NT) keep_sm[k] = (ne ? (ne[k] != 0) : 1) && (obj[k] > fp.conf);
    } else {
        const bool d2 = fp.flags & OVDET_NMS_2D;
        CornerSrc src{p.corners + (size_t)s * K * 24, obj, ne, cls_sm, d2 ? 1 : 0};
        nms_core(src, K, d2 ? 2 : 3, (fp.flags & OVDET_NMS_SAMECLS) != 0, (fp.flags & OVDET_NMS_OLD_TYPE) != 0, fp.nms_iou, 0.0, sh, nullptr, fp.dbg);
        const int na = sh.misc[0];
        for (int pos = tid; pos < na; pos += AM_NT)
            if (sh.picked[pos]) { const int k = sh.sidx[pos]; keep_sm[k] = obj[k] > fp.conf; }
    }
    __syncthreads();
    NSTAMP(fp.dbg, 11);
    if (fp.keep_out) for (int k = tid; k < K; k += AM_NT) fp.keep_out[(size_t)s * K + k] = keep_sm[k];
    // ---- matching on what is now in shared memory
    AmScene sc;
    const bool per_class = (fp.flags & OVDET_FRONT_PER_CLASS) != 0;
    sc.probs = per_class ? probs : nullptr;
    sc.ptile = (per_class && TILE) ? ptile : nullptr;
    sc.ptile_pitch = pitch;
    sc.score1 = (!per_class && (fp.flags & OVDET_FRONT_CLS_CONF)) ? clsp_sm : obj;
    sc.keep = keep_sm;
    sc.det_cls = per_class ? nullptr : cls_sm;
    // the match areas alias the NMS tables (dead now) and grow into the tile behind them only after the record phase,
    // its last reader
    am_scene_body(p, sc, un, hi, s);
}

// ------------------------------------------------- segmented radix sort + AP
constexpr int RS_NT = 256, RS_IPT = 8, RS_TILE = RS_NT * RS_IPT;

__global__ void __launch_bounds__(RS_NT) rs_prep_kernel(const float *__restrict__ score, const uint8_t *__restrict__ tp,
                                                        uint32_t *keys, uint8_t *vals, unsigned long long *nvalid, long long N)
{
    const int c = blockIdx.y;
    int local = 0;
    for (long long i = (long long)blockIdx.x * RS_NT + threadIdx.x; i < N; i += (long long)gridDim.x * RS_NT) {
        const float s = score[(size_t)c * N + i];
        keys[(size_t)c * N + i] = score_key(s);
        vals[(size_t)c * N + i] = tp[(size_t)c * N + i];
        local += (s > -INFINITY) ? 1 : 0;
    }
    for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&nvalid[c], (unsigned long long)local);
}

__global__ void __launch_bounds__(RS_NT) rs_hist_kernel(const uint32_t *__restrict__ keys, uint32_t *hist, long long N, int tiles, int shift)
{
    __shared__ unsigned int h[256];
    const int c = blockIdx.y, tile = blockIdx.x;
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)tile * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const long long i = base + j * RS_NT + threadIdx.x;
        if (i < N) atomicAdd(&h[(keys[(size_t)c * N + i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[((size_t)c * 256 + threadIdx.x) * tiles + tile] = h[threadIdx.x];
}

// exclusive scan of hist[c][0 .. 256*tiles) (digit-major), one CTA per class
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *hist, int tiles)
{
    __shared__ unsigned int wsum[32];
    __shared__ unsigned int carry_s;
    uint32_t *h = hist + (size_t)blockIdx.x * 256 * tiles;
    const int L = 256 * tiles;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < L; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned int v = i < L ? h[i] : 0u;
        unsigned int x = v;
        for (int off = 1; off < 32; off <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, off); if ((threadIdx.x & 31) >= off) x += y; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned int w = wsum[threadIdx.x];
            for (int off = 1; off < 32; off <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, w, off); if (threadIdx.x >= off) w += y; }
            wsum[threadIdx.x] = w;
        }
        __syncthreads();
        const unsigned int carry = carry_s;
        const unsigned int incl = x + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0u);
        if (i < L) h[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_NT) rs_scatter_kernel(const uint32_t *__restrict__ kin, const uint8_t *__restrict__ vin,
                                                           uint32_t *kout, uint8_t *vout, const uint32_t *__restrict__ hist,
                                                           long long N, int tiles, int shift)
{
    __shared__ unsigned int whist[RS_NT / 32][256];
    __shared__ unsigned int base[256];
    const int c = blockIdx.y, tile = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int w = 0; w < RS_NT / 32; ++w) whist[w][threadIdx.x] = 0;
    __syncthreads();
    // warp w owns the contiguous 256 elements [tile*2048 + w*256, +256): order (warp, round, lane) == index order
    const long long wbase = (long long)tile * RS_TILE + (long long)warp * (32 * RS_IPT);
    uint32_t key[RS_IPT]; uint8_t val[RS_IPT]; unsigned int rank[RS_IPT];
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const long long i = wbase + j * 32 + lane;
        const bool valid = i < N;
        key[j] = valid ? kin[(size_t)c * N + i] : 0xFFFFFFFFu;
        val[j] = valid ? vin[(size_t)c * N + i] : 0;
        const unsigned int digit = (key[j] >> shift) & 255u;
        const unsigned int peers = __match_any_sync(0xffffffffu, valid ? digit : 0x1000u + lane);
        const int lead = __ffs(peers) - 1;
        unsigned int old = 0;
        if (valid && lane == lead) { old = whist[warp][digit]; whist[warp][digit] = old + __popc(peers); }
        __syncwarp();
        old = __shfl_sync(0xffffffffu, old, lead);
        rank[j] = old + __popc(peers & ((1u << lane) - 1));
    }
    __syncthreads();
    {   // exclusive scan over warps for digit = threadIdx.x, plus the global base of this (digit, tile)
        const int d = threadIdx.x;
        unsigned int run = 0;
        for (int w = 0; w < RS_NT / 32; ++w) { const unsigned int t = whist[w][d]; whist[w][d] = run; run += t; }
        base[d] = hist[((size_t)c * 256 + d) * tiles + tile];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const long long i = wbase + j * 32 + lane;
        if (i < N) {
            const unsigned int digit = (key[j] >> shift) & 255u;
            const size_t pos = (size_t)c * N + base[digit] + whist[warp][digit] + rank[j];
            kout[pos] = key[j];
            vout[pos] = val[j];
        }
    }
}

// Cumulative TP/FP, precision/recall and VOC AP over one sorted class segment.
// grid = (C, nthr), 1024 threads.  Phase 1: per-chunk TP totals; phase 2: chunks
// in reverse with the precision envelope (running max from the right, eval_det.py:45-46)
// carried across chunks; AP = sum over TP positions of (rec_i - rec_{i-1}) * env_i (:50-54).
struct ApScanParams {
    const uint8_t *vals; const unsigned long long *nvalid; const long long *npos;
    long long N; int C, nthr, use07;
    double *ap, *recall, *rec_out, *prec_out; long long *ndet_out;
};

__global__ void __launch_bounds__(1024) ap_scan_kernel(ApScanParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned int *chunk_tp = reinterpret_cast<unsigned int *>(sm);  // [nchunks + 1] exclusive prefix
    __shared__ unsigned int wsum[32];
    __shared__ double wmax[32];
    __shared__ double red[32];
    __shared__ double carry_max_s;
    const int c = blockIdx.x, t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n = (long long)p.nvalid[c];
    const double npos = (double)p.npos[c];
    const uint8_t *v = p.vals + (size_t)c * p.N;
    const int nchunks = (int)((n + 1023) / 1024);
    const double eps = 2.220446049250313e-16;
    // ---- phase 1: chunk totals
    for (int b = warp; b < nchunks; b += 32) {
        unsigned int cnt = 0;
        for (int k = lane; k < 1024; k += 32) {
            const long long i = (long long)b * 1024 + k;
            if (i < n) cnt += (v[i] >> t) & 1u;
        }
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) chunk_tp[b + 1] = cnt;
    }
    if (tid == 0) { chunk_tp[0] = 0; carry_max_s = 0.0; }
    __syncthreads();
    if (tid == 0) for (int b = 0; b < nchunks; ++b) chunk_tp[b + 1] += chunk_tp[b];
    __syncthreads();
    const unsigned int total_tp = chunk_tp[nchunks];
    double ap_local = 0.0;
    double p11[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) p11[k] = 0.0;
    // ---- phase 2: reverse over chunks
    for (int b = nchunks - 1; b >= 0; --b) {
        const long long i = (long long)b * 1024 + tid;
        const bool valid = i < n;
        const unsigned int tp = valid ? ((v[i] >> t) & 1u) : 0u;
        unsigned int x = tp;
        for (int off = 1; off < 32; off <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (tid < 32) {
            unsigned int w = wsum[tid];
            for (int off = 1; off < 32; off <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, w, off); if (tid >= off) w += y; }
            wsum[tid] = w;
        }
        __syncthreads();
        const unsigned int ctp_u = chunk_tp[b] + x + (warp ? wsum[warp - 1] : 0u);
        const double ctp = (double)ctp_u;
        const double tot = (double)(i + 1);                 // tp + fp
        const double prec = valid ? __ddiv_rn(ctp, fmax(tot, eps)) : 0.0;
        const double rec = npos > 0.0 ? __ddiv_rn(ctp, npos) : 0.0;
        if (valid && p.rec_out) {
            p.rec_out[((size_t)t * p.C + c) * p.N + i] = rec;
            p.prec_out[((size_t)t * p.C + c) * p.N + i] = prec;
        }
        // reverse inclusive max scan of prec within the chunk, then fold the carry of later chunks
        double m = prec;
        for (int off = 1; off < 32; off <<= 1) { const double y = __shfl_down_sync(0xffffffffu, m, off); if (lane + off < 32) m = fmax(m, y); }
        if (lane == 0) wmax[warp] = m;
        __syncthreads();
        if (tid < 32) {
            double w = wmax[tid];
            for (int off = 1; off < 32; off <<= 1) { const double y = __shfl_down_sync(0xffffffffu, w, off); if (tid + off < 32) w = fmax(w, y); }
            wmax[tid] = w;
        }
        __syncthreads();
        const double carry = carry_max_s;
        double env = fmax(m, carry);
        if (warp < 31) env = fmax(env, wmax[warp + 1]);
        if (valid && tp) {
            const double rec_prev = npos > 0.0 ? __ddiv_rn(ctp - 1.0, npos) : 0.0;
            ap_local += __dmul_rn(__dsub_rn(rec, rec_prev), env);
        }
        if (valid && p.use07) {
#pragma unroll
            for (int k = 0; k < 11; ++k) if (rec >= k * 0.1) p11[k] = fmax(p11[k], prec);
        }
        __syncthreads();
        if (tid == 0) carry_max_s = fmax(carry, wmax[0]);
        __syncthreads();
    }
    // ---- deterministic block reductions
    double res;
    if (!p.use07) {
        double a = ap_local;
        for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
        if (lane == 0) red[warp] = a;
        __syncthreads();
        if (tid < 32) {
            double w = red[tid];
            for (int off = 16; off > 0; off >>= 1) w += __shfl_down_sync(0xffffffffu, w, off);
            if (tid == 0) red[0] = w;
        }
        __syncthreads();
        res = red[0];
    } else {
        res = 0.0;
        for (int k = 0; k < 11; ++k) {
            double a = p11[k];
            for (int off = 16; off > 0; off >>= 1) a = fmax(a, __shfl_down_sync(0xffffffffu, a, off));
            __syncthreads();
            if (lane == 0) red[warp] = a;
            __syncthreads();
            if (tid < 32) {
                double w = red[tid];
                for (int off = 16; off > 0; off >>= 1) w = fmax(w, __shfl_down_sync(0xffffffffu, w, off));
                if (tid == 0) red[0] = w;
            }
            __syncthreads();
            res = res + red[0] / 11.0;
        }
    }
    if (tid == 0) {
        p.ap[(size_t)t * p.C + c] = n > 0 ? res : 0.0;
        p.recall[(size_t)t * p.C + c] = (n > 0 && npos > 0.0) ? __ddiv_rn((double)total_tp, npos) : 0.0;
        if (p.ndet_out && t == 0) p.ndet_out[c] = n;
    }
}

static inline size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace ovdet

using namespace ovdet;

extern "C" int ovdet_box3d_iou_f64(const float *dets, const float *gts, const int32_t *nd, const int32_t *ng,
                                   int S, int D, int G, double *out, double *out2d, void *stream)
{
    OVDET_REQUIRE(S >= 0 && D >= 0 && G >= 0, "negative size");
    if (S == 0 || D == 0 || G == 0) return OVDET_OK;
    OVDET_REQUIRE(dets && gts && out, "null pointer");
    IouParams p{dets, gts, nd, ng, S, D, G, out, out2d, (long long)S * D * G};
    const size_t smem = sizeof(V2<double>) * 2 * SH_MAXV * EV_NT;
    OVDET_CUDA_TRY(ensure_dyn_smem(box3d_iou_kernel, smem));
    long long blocks = (p.total + EV_NT - 1) / EV_NT;
    if (blocks > 148 * 12) blocks = 148 * 12;
    box3d_iou_kernel<<<(unsigned)blocks, EV_NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return launch_ok("box3d_iou_kernel");
}

static int fill_match_params(MatchParams &p, const float *corners, const float *probs, const float *obj, const uint8_t *keep,
                             const int32_t *det_cls, const float *gt_corners, const int64_t *gt_labels, const uint8_t *gt_present,
                             int S, int K, int G, int C, const double *thr, int nthr,
                             double *iou_ws, float *rec_score, uint8_t *rec_tp, int64_t *npos)
{
    OVDET_REQUIRE(G == 0 || (gt_corners && gt_labels && gt_present && iou_ws), "null GT pointer");
    OVDET_REQUIRE(nthr >= 1 && nthr <= 8, "1..8 thresholds");
    OVDET_REQUIRE(K <= 32767 && G <= 32767, "K, G must fit int16");
    OVDET_REQUIRE((long long)K * G < 2147483647LL, "K*G too large");
    p.corners = corners; p.probs = probs; p.obj = obj; p.keep = keep; p.det_cls = det_cls; p.gt_corners = gt_corners;
    p.gt_labels = gt_labels; p.gt_present = gt_present; p.S = S; p.K = K; p.G = G; p.C = C; p.nthr = nthr;
    for (int t = 0; t < nthr; ++t) p.thr[t] = thr[t];
    p.iou_ws = iou_ws; p.rec_score = rec_score; p.rec_tp = rec_tp; p.npos = reinterpret_cast<unsigned long long *>(npos);
    p.tp_key = nullptr; p.tp_bits = nullptr; p.tp_cnt = nullptr; p.tp_cap = 0; p.gt_present_f32 = nullptr; p.iou_given = nullptr;
    { const char *e = getenv("OVDET_APMATCH_DBG_PTR"); p.dbg = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0)) : nullptr; }
    return OVDET_OK;
}

extern "C" int ovdet_ap_match(const float *corners, const float *probs, const float *obj, const uint8_t *keep,
                              const int32_t *det_cls, const float *gt_corners, const int64_t *gt_labels, const uint8_t *gt_present,
                              int S, int K, int G, int C, const double *thr, int nthr,
                              double *iou_ws, float *rec_score, uint8_t *rec_tp, int64_t *npos, void *stream)
{
    OVDET_REQUIRE(S >= 0 && K > 0 && G >= 0 && C > 0, "bad size");
    if (S == 0) return OVDET_OK;
    OVDET_REQUIRE(corners && obj && keep && rec_score && rec_tp && npos && thr, "null pointer");
    OVDET_REQUIRE(probs || det_cls, "need probs (per-class proposals) or det_cls");
    MatchParams p;
    { const int rc = fill_match_params(p, corners, probs, obj, keep, det_cls, gt_corners, gt_labels, gt_present, S, K, G, C, thr, nthr, iou_ws, rec_score, rec_tp, npos); if (rc) return rc; }
    const size_t smem = am_smem_bytes(K, G, nthr);
    OVDET_REQUIRE(smem <= 220 * 1024, "K + G too large for the shared-memory feature records");
    OVDET_CUDA_TRY(ensure_dyn_smem(ap_match_kernel, smem));
    ap_match_kernel<<<S, AM_NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return launch_ok("ap_match_kernel");
}

extern "C" int ovdet_aabb_iou_f64(const double *dets, const double *gts, const int32_t *nd, const int32_t *ng,
                                  int S, int D, int G, double *out, void *stream)
{
    OVDET_REQUIRE(S >= 0 && D >= 0 && G >= 0, "negative size");
    if (S == 0 || D == 0 || G == 0) return OVDET_OK;
    OVDET_REQUIRE(dets && gts && out, "null pointer");
    AabbParams p{dets, gts, nd, ng, S, D, G, out, (long long)S * D * G};
    long long blocks = (p.total + EV_NT - 1) / EV_NT;
    if (blocks > 148 * 16) blocks = 148 * 16;
    aabb_iou_kernel<<<(unsigned)blocks, EV_NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return launch_ok("aabb_iou_kernel");
}

extern "C" int ovdet_ap_match_iou(const double *iou, const float *probs, const float *obj, const uint8_t *keep,
                                  const int32_t *det_cls, const int64_t *gt_labels, const uint8_t *gt_present,
                                  int S, int K, int G, int C, const double *thr, int nthr,
                                  float *rec_score, uint8_t *rec_tp, int64_t *npos, void *stream)
{
    OVDET_REQUIRE(S >= 0 && K > 0 && G >= 0 && C > 0, "bad size");
    if (S == 0) return OVDET_OK;
    OVDET_REQUIRE(obj && keep && rec_score && rec_tp && npos && thr, "null pointer");
    OVDET_REQUIRE(probs || det_cls, "need probs (per-class proposals) or det_cls");
    OVDET_REQUIRE(G == 0 || iou, "null IoU matrix");
    MatchParams p;
    double dummy = 0.0;
    { const int rc = fill_match_params(p, nullptr, probs, obj, keep, det_cls, reinterpret_cast<const float *>(&dummy), gt_labels, gt_present, S, K, G, C, thr, nthr, &dummy, rec_score, rec_tp, npos); if (rc) return rc; }
    p.gt_corners = nullptr; p.iou_ws = nullptr;
    p.iou_given = G > 0 ? iou : &dummy;   // G == 0: nothing is ever read
    const size_t smem = am_smem_bytes(K, G, nthr);
    OVDET_REQUIRE(smem <= 220 * 1024, "K + G too large for the shared-memory tables");
    OVDET_CUDA_TRY(ensure_dyn_smem(ap_match_kernel, smem));
    ap_match_kernel<<<S, AM_NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return launch_ok("ap_match_kernel");
}

// the generic fused front end (any K <= 1024, any NMS threshold); ovdet_ap_front_f32 (ap_front.cu) picks between this and
// the lean thread-per-box kernel
int ovdet::front1_launch(const float *corners, const float *probs, const float *obj, const uint8_t *nonempty,
                         const float *gt_corners, const int64_t *gt_labels, const void *gt_present,
                         int S, int K, int G, int C, double nms_iou, float conf_thresh, unsigned flags,
                         const double *thr, int nthr, double *iou_ws, float *rec_score, uint8_t *rec_tp, int64_t *npos,
                         uint32_t *tp_key, uint8_t *tp_bits, int32_t *tp_cnt, int tp_cap, uint8_t *keep_out, void *stream)
{
    OVDET_REQUIRE(S >= 0 && K > 0 && G >= 0 && C > 0, "bad size");
    if (S == 0) return OVDET_OK;
    OVDET_REQUIRE(corners && probs && obj && rec_score && npos && thr, "null pointer");
    OVDET_REQUIRE(K <= NMS_MAXK, "K must be <= 1024");
    OVDET_REQUIRE((tp_key == nullptr) == (tp_bits == nullptr) && (tp_key == nullptr) == (tp_cnt == nullptr), "tp_key, tp_bits and tp_cnt go together");
    OVDET_REQUIRE(tp_key == nullptr || tp_cap > 0, "tp_cap must be positive");
    OVDET_REQUIRE(rec_tp || tp_key, "need rec_tp and/or a TP list to report the true positives");
    FrontParams fp;
    { const int rc = fill_match_params(fp.m, corners, probs, obj, nullptr, nullptr, gt_corners, gt_labels, static_cast<const uint8_t *>(gt_present), S, K, G, C, thr, nthr, iou_ws, rec_score, rec_tp, npos); if (rc) return rc; }
    fp.m.tp_key = tp_key; fp.m.tp_bits = tp_bits; fp.m.tp_cnt = tp_cnt; fp.m.tp_cap = tp_cap;
    if (flags & OVDET_FRONT_GT_PRESENT_F32) { fp.m.gt_present_f32 = reinterpret_cast<const float *>(gt_present); fp.m.gt_present = nullptr; }
    fp.nonempty = nonempty; fp.nms_iou = nms_iou; fp.conf = conf_thresh; fp.flags = flags; fp.keep_out = keep_out;
    { const char *e = getenv("OVDET_APFRONT_DBG_PTR"); fp.dbg = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0)) : nullptr; }
    // the probability tile stays in shared memory from the argmax to the record phase (nothing reads it later, so the
    // matching areas may grow into it) as long as the CTA stays small enough for >= 2 per SM
    const bool tile = front_smem_bytes(K, G, C, nthr, true) <= 100 * 1024;
    const size_t smem = front_smem_bytes(K, G, C, nthr, tile);
    OVDET_REQUIRE(smem <= 220 * 1024, "K + G too large for the shared-memory tables");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (tile) {
        OVDET_CUDA_TRY(ensure_dyn_smem(ap_front_kernel<true>, smem));
        ap_front_kernel<true><<<S, AM_NT, smem, st>>>(fp);
    } else {
        OVDET_CUDA_TRY(ensure_dyn_smem(ap_front_kernel<false>, smem));
        ap_front_kernel<false><<<S, AM_NT, smem, st>>>(fp);
    }
    return launch_ok("ap_front_kernel");
}

extern "C" size_t ovdet_ap_reduce_ws_bytes(int C, int64_t N)
{
    if (C <= 0 || N <= 0) return 256;
    const size_t cn = (size_t)C * (size_t)N;
    const size_t tiles = ((size_t)N + RS_TILE - 1) / RS_TILE;
    return 2 * a256(cn * 4) + 2 * a256(cn) + a256((size_t)C * 256 * tiles * 4) + a256((size_t)C * 8) + 256;
}

extern "C" int ovdet_ap_reduce(const float *rec_score, const uint8_t *rec_tp, const int64_t *npos,
                               int C, int64_t N, int nthr, int use_07_metric,
                               double *ap, double *recall, int64_t *n_det, double *rec_out, double *prec_out,
                               void *ws, size_t ws_bytes, void *stream)
{
    OVDET_REQUIRE(C > 0 && N >= 0 && nthr >= 1 && nthr <= 8, "bad size");
    OVDET_REQUIRE(ap && recall && npos, "null pointer");
    OVDET_REQUIRE((rec_out == nullptr) == (prec_out == nullptr), "rec_out and prec_out go together");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (N == 0) {
        OVDET_CUDA_TRY(cudaMemsetAsync(ap, 0, sizeof(double) * nthr * C, st));
        OVDET_CUDA_TRY(cudaMemsetAsync(recall, 0, sizeof(double) * nthr * C, st));
        if (n_det) OVDET_CUDA_TRY(cudaMemsetAsync(n_det, 0, sizeof(int64_t) * C, st));
        return OVDET_OK;
    }
    OVDET_REQUIRE(rec_score && rec_tp && ws, "null pointer");
    OVDET_REQUIRE(ws_bytes >= ovdet_ap_reduce_ws_bytes(C, N), "workspace too small");
    OVDET_REQUIRE(N < (1ll << 31), "N must be < 2^31");
    const size_t cn = (size_t)C * (size_t)N;
    const int tiles = (int)((N + RS_TILE - 1) / RS_TILE);
    char *w = static_cast<char *>(ws);
    uint32_t *kA = reinterpret_cast<uint32_t *>(w); w += a256(cn * 4);
    uint32_t *kB = reinterpret_cast<uint32_t *>(w); w += a256(cn * 4);
    uint8_t *vA = reinterpret_cast<uint8_t *>(w); w += a256(cn);
    uint8_t *vB = reinterpret_cast<uint8_t *>(w); w += a256(cn);
    uint32_t *hist = reinterpret_cast<uint32_t *>(w); w += a256((size_t)C * 256 * tiles * 4);
    unsigned long long *nvalid = reinterpret_cast<unsigned long long *>(w);
    OVDET_CUDA_TRY(cudaMemsetAsync(nvalid, 0, sizeof(unsigned long long) * C, st));
    {
        int gx = (int)((N + RS_NT - 1) / RS_NT);
        if (gx > 148 * 8) gx = 148 * 8;
        rs_prep_kernel<<<dim3(gx, C), RS_NT, 0, st>>>(rec_score, rec_tp, kA, vA, nvalid, N);
    }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        rs_hist_kernel<<<dim3(tiles, C), RS_NT, 0, st>>>(kA, hist, N, tiles, shift);
        rs_scan_kernel<<<C, 1024, 0, st>>>(hist, tiles);
        rs_scatter_kernel<<<dim3(tiles, C), RS_NT, 0, st>>>(kA, vA, kB, vB, hist, N, tiles, shift);
        uint32_t *tk = kA; kA = kB; kB = tk;
        uint8_t *tv = vA; vA = vB; vB = tv;
    }
    ApScanParams sp;
    sp.vals = vA; sp.nvalid = nvalid; sp.npos = reinterpret_cast<const long long *>(npos); sp.N = N; sp.C = C; sp.nthr = nthr;
    sp.use07 = use_07_metric; sp.ap = ap; sp.recall = recall; sp.rec_out = rec_out; sp.prec_out = prec_out;
    sp.ndet_out = reinterpret_cast<long long *>(n_det);
    const size_t smem = sizeof(unsigned int) * ((size_t)(N + 1023) / 1024 + 2);
    if (smem > 48 * 1024) OVDET_CUDA_TRY(ensure_dyn_smem(ap_scan_kernel, smem));
    OVDET_REQUIRE(smem <= 200 * 1024, "N too large for the chunk table");
    ap_scan_kernel<<<dim3(C, nthr), 1024, smem, st>>>(sp);
    return launch_ok("ap_reduce");
}
