// ap_front.cu -- the AP front end of one scene in one small CTA, thread k <-> predicted box k.
//
// parse_predictions (utils/ap_calculator.py:39-238: argmax class, AABB from corners, greedy NMS variant, confidence
// gate) and the AP matching of the same scene (utils/eval_det.py:117-140) for K <= 256 boxes, NMS threshold >= 0.
// Compared with the generic kernel in eval.cu (nms_core + am_scene_body, kept for everything else) nothing is staged
// or sorted:
//   load     thread k reads ITS box (six 16-byte loads) and ITS class-probability row, reduces them in registers to
//            the AABB, argmax class and the scores, and writes the score records of all classes straight away
//            (speculatively: a box that NMS or the confidence gate drops later overwrites its records with -inf);
//            threads g < G turn the present GT boxes into 64-byte feature records
//   nms      same-class boxes are found through per-class member bitmasks (one shared atomicOr per box); thread k tests
//            only the classmates that come BEFORE it in score order and keeps a bitmask of those that would suppress
//            it; the greedy pick is then the fixed point of  picked(k) <=> no picked box in sup(k),  reached in as many
//            rounds as the longest suppression chain (2-3), two barriers each -- no sort, no serial scan
//   match    thread k walks the present GT: exact rejects and a conservative fp32 IoU upper bound leave ~0.2 candidate
//            pairs per box; only their boxes get fp64 feature records and the fp64 Sutherland-Hodgman clip (8 lanes per
//            pair); claims and true positives per candidate as in eval.cu
// A scene with more candidates than the shared queue holds (or a negative IoU threshold, where every pair counts) is
// processed in slabs of boxes, claims first, true positives in a second sweep -- no global workspace.
#include <stdlib.h>

#include "ap_match.cuh"

namespace ovdet {

int front1_launch(const float *corners, const float *probs, const float *obj, const uint8_t *nonempty,
                  const float *gt_corners, const int64_t *gt_labels, const void *gt_present,
                  int S, int K, int G, int C, double nms_iou, float conf_thresh, unsigned flags,
                  const double *thr, int nthr, double *iou_ws, float *rec_score, uint8_t *rec_tp, int64_t *npos,
                  uint32_t *tp_key, uint8_t *tp_bits, int32_t *tp_cnt, int tp_cap, uint8_t *keep_out, void *stream);   // eval.cu

constexpr int F2_QCAP = 256;    // candidate pairs per slab
constexpr int F2_COOP = 64;     // up to this many candidates go through the 8-lane cooperative clipper

#define F2STAMP(i) do { if (p.dbg && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.dbg[(size_t)blockIdx.x * 16 + (i)] = t_; } } while (0)

struct Front2Params {
    const float *corners, *probs, *obj; const uint8_t *nonempty;
    const float *gt_corners; const int64_t *gt_labels; const uint8_t *gt_present; const float *gt_present_f32;
    int S, K, G, C, nthr; double thr[8]; double nms_iou; float conf; unsigned flags;
    float *rec_score; uint8_t *rec_tp; unsigned long long *npos;
    uint32_t *tp_key; uint8_t *tp_bits; int *tp_cnt; int tp_cap;
    uint8_t *keep_out; unsigned long long *dbg;
};

// Shared-memory layout as BYTE OFFSETS from the CTA's dynamic shared base (uniform values the compiler keeps in uniform
// registers or rematerialises; a struct of 19 generic pointers instead ended up in local memory and every shared access
// paid a local load first).
struct F2Smem {
    unsigned char *base;
    int o_cbox, o_scratch, o_qiou, o_vol, o_best, o_box6, o_score, o_gvlo, o_garea, o_qscore, o_cls, o_glab, o_queue, o_cmask,
        o_alive, o_picked, o_done, o_gmask;
    template <typename T> __device__ __forceinline__ T *at(int off) const { return reinterpret_cast<T *>(base + off); }
};
#define S_gbox (S.at<AmBox>(0))
#define S_cbox (S.at<AmBox>(S.o_cbox))
#define S_scratch (S.at<V2<double>>(S.o_scratch))
#define S_qiou (S.at<double>(S.o_qiou))
#define S_vol (S.at<double>(S.o_vol))
#define S_best (S.at<unsigned long long>(S.o_best))
#define S_box6 (S.at<float>(S.o_box6))
#define S_score (S.at<float>(S.o_score))
#define S_gvlo (S.at<float>(S.o_gvlo))
#define S_garea (S.at<float>(S.o_garea))
#define S_qscore (S.at<float>(S.o_qscore))
#define S_cls (S.at<int>(S.o_cls))
#define S_glab (S.at<int>(S.o_glab))
#define S_queue (S.at<unsigned>(S.o_queue))
#define S_cmask (S.at<uint32_t>(S.o_cmask))
#define S_alive (S.at<uint32_t>(S.o_alive))
#define S_picked (S.at<uint32_t>(S.o_picked))
#define S_done (S.at<uint32_t>(S.o_done))
#define S_gmask (S.at<uint32_t>(S.o_gmask))

__host__ __device__ inline size_t f2_scratch_bytes(int nt)
{
    const size_t coop = sizeof(V2<double>) * 8 * (size_t)(nt / 8), serial = sizeof(V2<double>) * 2 * SH_MAXV * AM_CLIP;
    return coop > serial ? coop : serial;
}
__host__ __device__ inline int f2_cmask_classes(int C, unsigned flags)
{
    return ((flags & OVDET_NMS_SAMECLS) && !(flags & OVDET_PARSE_NO_NMS)) ? C : 1;
}
__host__ __device__ inline size_t f2_smem_bytes(int K, int G, int C, int nthr, int nt, unsigned flags)
{
    const int W = (K + 31) / 32, WG = (G + 31) / 32;
    size_t b = sizeof(AmBox) * ((size_t)G + F2_COOP) + f2_scratch_bytes(nt) + sizeof(double) * (F2_QCAP + (size_t)K) + sizeof(unsigned long long) * (size_t)G * nthr;
    b += sizeof(float) * (7 * (size_t)K + 2 * (size_t)G + F2_QCAP) + sizeof(int) * ((size_t)K + G) + sizeof(unsigned) * F2_QCAP;
    b += sizeof(uint32_t) * ((size_t)f2_cmask_classes(C, flags) * W + 3 * W + WG);
    return (b + 15) & ~(size_t)15;
}
__device__ __forceinline__ F2Smem f2_carve(unsigned char *base, int K, int G, int C, int nthr, int nt, unsigned flags)
{
    const int W = (K + 31) / 32;
    F2Smem s;
    s.base = base;
    int o = (int)sizeof(AmBox) * G;
    s.o_cbox = o; o += (int)sizeof(AmBox) * F2_COOP;
    s.o_scratch = o; o += (int)f2_scratch_bytes(nt);
    s.o_qiou = o; o += 8 * F2_QCAP;
    s.o_vol = o; o += 8 * K;
    s.o_best = o; o += 8 * G * nthr;
    s.o_box6 = o; o += 24 * K;
    s.o_score = o; o += 4 * K;
    s.o_gvlo = o; o += 4 * G;
    s.o_garea = o; o += 4 * G;
    s.o_qscore = o; o += 4 * F2_QCAP;
    s.o_cls = o; o += 4 * K;
    s.o_glab = o; o += 4 * G;
    s.o_queue = o; o += 4 * F2_QCAP;
    s.o_cmask = o; o += 4 * f2_cmask_classes(C, flags) * W;
    s.o_alive = o; o += 4 * W;
    s.o_picked = o; o += 4 * W;
    s.o_done = o; o += 4 * W;
    s.o_gmask = o;
    return s;
}

// fp32 UPPER bound of the area of the BEV quad (corners 0..3, x/z) -- shoelace relative to corner 0 (well conditioned)
__device__ __forceinline__ float bev_area_hi(const float *c)
{
    const float x1 = c[3] - c[0], z1 = c[5] - c[2], x2 = c[6] - c[0], z2 = c[8] - c[2], x3 = c[9] - c[0], z3 = c[11] - c[2];
    return 0.5f * fabsf((x1 * z2 - x2 * z1) + (x2 * z3 - x3 * z2)) * (1.f + 1e-4f);
}

// The cold or repeated pieces are real calls: the kernel's hot loop then fits the instruction cache (the fully inlined
// version was 190 KB of SASS and spent 20-50 % of its stall samples waiting for instructions).
__device__ __noinline__ void f2_box_features(const float *g, AmBox *out)
{
    float c[24];
    am_load_box(g, c);
    am_features(c, *out);
}

struct CoopState { double vx, vy; int n; };
// one Sutherland-Hodgman pass against edge (e-1 mod 4 -> e) of the GT quad; the quad is read from the feature record
// here so that the caller keeps nothing but the polygon state alive across the call
__device__ __noinline__ CoopState f2_coop_pass(const AmBox *gt, int e, CoopState st, int gl, int gshift, V2<double> *gbuf)
{
    const int a = (e + 3) & 3;
    st.n = coop_pass(ClipEdge<double>((double)gt->qx[a], (double)gt->qz[a], (double)gt->qx[e], (double)gt->qz[e]), st.vx, st.vy, st.n, gl, gshift, gbuf);
    return st;
}

__device__ __noinline__ double f2_serial_iou(const float *det, const AmBox *gt, V2<double> *bufA, V2<double> *bufB)
{
    AmBox a;
    f2_box_features(det, &a);
    double sq[8], cq[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        sq[2 * t] = (double)a.qx[t]; sq[2 * t + 1] = (double)a.qz[t];
        cq[2 * t] = (double)gt->qx[t]; cq[2 * t + 1] = (double)gt->qz[t];
    }
    const int n = sh_clip_quads<double, AM_CLIP>(sq, cq, bufA, bufB);
    return am_finish_iou(area_f64<AM_CLIP>(bufB, n), a, *gt);
}

// overlap of two AABBs that do intersect (bounds already reduced in fp32, exact) against the NMS threshold, in the
// reference's fp64 operation order (utils/nms.py:79-117); vj = the earlier (picked) box, vk = the later one
__device__ __noinline__ bool f2_nms_suppresses(float l0, float h0, float l1, float h1, float l2, float h2, double vj, double vk,
                                               int dims, bool old_type, double thr)
{
    using A = Ar<double>;
    double inter = A::mul(A::sub((double)h0, (double)l0), A::sub((double)h1, (double)l1));
    if (dims == 3) inter = A::mul(inter, A::sub((double)h2, (double)l2));
    const double o = old_type ? A::div(inter, vk) : A::div(inter, A::sub(A::add(vj, vk), inter));
    return o > thr;
}

template <int NT>
__global__ void __launch_bounds__(NT, NT == 128 ? 8 : 4) ap_front2_kernel(Front2Params p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    using A = Ar<double>;
    const F2Smem S = f2_carve(sm, p.K, p.G, p.C, p.nthr, NT, p.flags);
    __shared__ int qn_s;
    const int s = blockIdx.x, k = threadIdx.x, lane = k & 31, warp = k >> 5;
    const int K = p.K, C = p.C, G = p.G, W = (K + 31) >> 5, WG = (G + 31) >> 5;
    const size_t N = (size_t)p.S * K, slot = (size_t)s * K + k;
    const bool per_class = (p.flags & OVDET_FRONT_PER_CLASS) != 0;
    const bool no_nms = (p.flags & OVDET_PARSE_NO_NMS) != 0;
    const bool samecls = (p.flags & OVDET_NMS_SAMECLS) && !no_nms;
    const bool d2 = (p.flags & OVDET_NMS_2D) != 0;
    const unsigned mybit = 1u << lane;
    F2STAMP(0);

    {   // zero what is accumulated with atomics
        const int ncm = f2_cmask_classes(C, p.flags) * W;
        for (int i = k; i < ncm; i += NT) S_cmask[i] = 0u;
        for (int i = k; i < 2 * W; i += NT) S_picked[i] = 0u;   // picked | done are adjacent
        for (int i = k; i < G * p.nthr; i += NT) S_best[i] = 0ull;
        if (k == 0) qn_s = 0;
    }
    __syncthreads();

    // ---- load: my box, my probability row
    bool alive = false;
    float obj = 0.f, ytop = 0.f, ybot = 0.f, lox = 0.f, hix = 0.f, loz = 0.f, hiz = 0.f, vlo = 0.f, ahi = 0.f, s1 = 0.f;
    float mn[3] = {0.f, 0.f, 0.f}, mx[3] = {0.f, 0.f, 0.f};
    int cls = 0;
    if (k < K) {
        alive = p.nonempty ? p.nonempty[slot] != 0 : true;
        obj = __ldg(p.obj + slot);
        {
            float c[24];
            am_load_box(p.corners + slot * 24, c);
#pragma unroll
            for (int a = 0; a < 3; ++a) {   // AABB over the 8 corners (ap_calculator.py:157-176)
                mn[a] = mx[a] = c[a];
#pragma unroll
                for (int i = 1; i < 8; ++i) { mn[a] = fminf(mn[a], c[3 * i + a]); mx[a] = fmaxf(mx[a], c[3 * i + a]); }
            }
            // what am_features keeps of a detection for the cheap rejects (BEV rectangle of corners 0..3, y of corners 0 / 4)
            ytop = c[1]; ybot = c[13];
            lox = fminf(fminf(c[0], c[3]), fminf(c[6], c[9])); hix = fmaxf(fmaxf(c[0], c[3]), fmaxf(c[6], c[9]));
            loz = fminf(fminf(c[2], c[5]), fminf(c[8], c[11])); hiz = fmaxf(fmaxf(c[2], c[5]), fmaxf(c[8], c[11]));
            // fp32 LOWER bound of the fp64 edge-length volume (box3d_vol): only used to widen the IoU upper bound
            auto edge = [&](int a0, int b0) {   // |corner a0 - corner b0| (constant indices: c stays in registers)
                const float dx = c[a0] - c[b0], dy = c[a0 + 1] - c[b0 + 1], dz = c[a0 + 2] - c[b0 + 2];
                return sqrtf(dx * dx + dy * dy + dz * dz);
            };
            const float v = edge(0, 3) * edge(3, 6) * edge(0, 12);   // edges (0,1), (1,2), (0,4) of box3d_vol
            vlo = v * (1.f - 2e-5f);
            ahi = bev_area_hi(c);
        }
        F2STAMP(7);
        // class probabilities: argmax = first maximum (ap_calculator.py:59-61); per-class layout: the records of all
        // classes are written right here (score = prob * objectness, :196-210), coalesced in k for each class
        const float *row = p.probs + slot * C;
        float bestp = -INFINITY;
        auto visit = [&](int c, float v) {
            if (v > bestp || c == 0) { bestp = v; cls = c; }
            if (per_class) p.rec_score[(size_t)c * N + slot] = alive ? __fmul_rn(v, obj) : -INFINITY;
        };
        if ((C & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
            for (int c = 0; c < C; c += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(row + c));
                visit(c, v.x); visit(c + 1, v.y); visit(c + 2, v.z); visit(c + 3, v.w);
            }
        } else {
            for (int c = 0; c < C; ++c) visit(c, __ldg(row + c));
        }
        s1 = (p.flags & OVDET_FRONT_CLS_CONF) ? bestp : obj;
        if (!per_class)   // single-class layouts (:212-236): the box only exists in its argmax class
            for (int c = 0; c < C; ++c) p.rec_score[(size_t)c * N + slot] = (alive && c == cls) ? s1 : -INFINITY;
        if (p.rec_tp) for (int c = 0; c < C; ++c) p.rec_tp[(size_t)c * N + slot] = 0;
        if (d2) { mn[1] = mn[2]; mx[1] = mx[2]; }   // 2D NMS works on the x and z extents (ap_calculator.py:92-104)
#pragma unroll
        for (int a = 0; a < 3; ++a) { S_box6[a * K + k] = mn[a]; S_box6[(3 + a) * K + k] = mx[a]; }
        {   // fp64 volume of the AABB as nms_3d_faster computes it (utils/nms.py:79-117 on the fp64 box rows)
            double v = A::sub((double)mx[0], (double)mn[0]);
            v = A::mul(v, A::sub((double)mx[1], (double)mn[1]));
            if (!d2) v = A::mul(v, A::sub((double)mx[2], (double)mn[2]));
            S_vol[k] = v;
        }
        F2STAMP(8);
        S_score[k] = obj;
        S_cls[k] = cls;
        if (alive) atomicOr(&S_cmask[(size_t)(samecls ? cls : 0) * W + warp], mybit);
    }
    {
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        if (lane == 0 && warp < W) S_alive[warp] = m;
    }
    // ---- present GT -> feature records (kept at their own index: the order of the present ones is what first-max needs)
    for (int g0 = warp * 32; g0 < G; g0 += NT) {
        const int g = g0 + lane;
        const size_t gs = (size_t)s * G + g;
        const bool present = g < G && (p.gt_present ? p.gt_present[gs] != 0 : p.gt_present_f32[gs] != 0.f);
        const unsigned m = __ballot_sync(0xffffffffu, present);
        if (lane == 0) S_gmask[g0 >> 5] = m;
        if (present) {
            f2_box_features(p.gt_corners + gs * 24, &S_gbox[g]);
            const long long lab = p.gt_labels[gs];
            S_glab[g] = (lab >= 0 && lab < C) ? (int)lab : -1;
            S_gvlo[g] = (float)S_gbox[g].vol * (1.f - 1e-6f);
            {   // upper bound of the BEV area from the feature record's quad (vertex order reversed: same area)
                const AmBox &b = S_gbox[g];
                const float q[12] = {b.qx[0], 0.f, b.qz[0], b.qx[1], 0.f, b.qz[1], b.qx[2], 0.f, b.qz[2], b.qx[3], 0.f, b.qz[3]};
                S_garea[g] = bev_area_hi(q);
            }
            if (lab >= 0 && lab < C) atomicAdd(&p.npos[lab], 1ull);
        }
    }
    F2STAMP(9);
    __syncthreads();
    F2STAMP(1);

    // ---- greedy NMS as a fixed point
    bool keep;
    if (no_nms) {
        keep = alive && obj > p.conf;
    } else {
        const int dims = d2 ? 2 : 3;
        const bool old_type = (p.flags & OVDET_NMS_OLD_TYPE) != 0;
        constexpr int F2_MAXW = NT / 32;   // K <= NT
        uint32_t sup[F2_MAXW];
#pragma unroll
        for (int w = 0; w < F2_MAXW; ++w) sup[w] = 0u;
        if (alive) {
            // my bounds in the NMS axis order; all comparisons of bounds are exact in fp32 (the reference's fp64 values are
            // these floats widened), fp64 only for the few pairs that really overlap
            const double vk = S_vol[k];
            const uint32_t *cm = S_cmask + (size_t)(samecls ? cls : 0) * W;
#pragma unroll
            for (int w = 0; w < F2_MAXW; ++w) {
                if (w >= W) break;
                uint32_t m = cm[w];
                if (w == warp) m &= ~mybit;
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const int j = 32 * w + b;
                    const float sj = S_score[j];
                    if (!(sj > obj || (sj == obj && j > k))) continue;   // only boxes picked before me can suppress me
                    float l[3], h[3];
                    bool empty = false;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {   // (mn, mx) still hold my bounds in the NMS axis order
                        l[a] = fmaxf(S_box6[a * K + j], mn[a]); h[a] = fminf(S_box6[(3 + a) * K + j], mx[a]);
                        if (a < dims) empty |= !(h[a] > l[a]);   // <=> max(0, min(hi) - max(lo)) == 0 in fp64
                    }
                    if (empty) continue;   // inter == 0: overlap 0 or NaN, never > thr (thr >= 0 here)
                    if (f2_nms_suppresses(l[0], h[0], l[1], h[1], l[2], h[2], S_vol[j], vk, dims, old_type, p.nms_iou)) sup[w] |= 1u << b;
                }
            }
        }
        F2STAMP(10);
        bool und = alive, picked_me = false;
        for (;;) {
            uint32_t anyp = 0u, pend = 0u;
#pragma unroll
            for (int w = 0; w < F2_MAXW; ++w) {
                if (w >= W) break;
                anyp |= sup[w] & S_picked[w];
                pend |= sup[w] & ~S_done[w];
            }
            __syncthreads();   // everybody has read this round's state
            if (und) {
                if (anyp) { atomicOr(&S_done[warp], mybit); und = false; }
                else if (!pend) { atomicOr(&S_picked[warp], mybit); atomicOr(&S_done[warp], mybit); und = false; picked_me = true; }
            }
            if (!__syncthreads_or(und ? 1 : 0)) break;
        }
        keep = picked_me && obj > p.conf;
        F2STAMP(11);
    }
    if (k < K) {
        if (alive && !keep) {   // take the speculative records back
            if (per_class) for (int c = 0; c < C; ++c) p.rec_score[(size_t)c * N + slot] = -INFINITY;
            else p.rec_score[(size_t)cls * N + slot] = -INFINITY;
        }
        if (p.keep_out) p.keep_out[slot] = keep ? 1 : 0;
    }
    F2STAMP(2);

    // ---- matching
    double thr_min = p.thr[0];
    for (int t = 1; t < p.nthr; ++t) thr_min = fmin(thr_min, p.thr[t]);
    const bool all_pairs = !(thr_min >= 0.0);
    const float thr_lo = all_pairs ? 0.f : (float)thr_min * (1.f - 1e-6f);
    int ng = 0;
    for (int w = 0; w < WG; ++w) ng += __popc(S_gmask[w]);
    if (ng == 0) return;   // uniform
    const float *myrow = p.probs + slot * C;
    auto enumerate = [&](int k0, int k1) {
        if (!(keep && k >= k0 && k < k1)) return;
        for (int w = 0; w < WG; ++w) {
            uint32_t m = S_gmask[w];
            while (m) {
                const int g = 32 * w + __ffs(m) - 1;
                m &= m - 1;
                const int c = S_glab[g];
                if (c < 0 || (!per_class && c != cls)) continue;   // can never match: not a candidate
                const AmBox &b = S_gbox[g];
                const float hh = fminf(ytop, b.ytop) - fmaxf(ybot, b.ybot);
                const float ox = fminf(hix, b.hix) - fmaxf(lox, b.lox), oz = fminf(hiz, b.hiz) - fmaxf(loz, b.loz);
                bool need = false, zero = false;
                if (all_pairs) {   // every pair counts; the exact rejects give IoU = 0 without a clip
                    need = true;
                    zero = !(fminf(ytop, b.ytop) > fmaxf(ybot, b.ybot) && !(hix < b.lox || b.hix < lox) && !(hiz < b.loz || b.hiz < loz));
                } else if (hh > 0.f && ox >= 0.f && oz >= 0.f) {
                    // IoU <= I / (V1 + V2 - I) with I = min(overlap of the BEV bounding rectangles, either BEV area) x height
                    // overlap >= the true intersection; fp32 with every rounding pushed to the safe side
                    const float I = fminf(ox * oz, fminf(ahi, S_garea[g])) * hh * (1.f + 1e-5f);
                    const float den = vlo + S_gvlo[g] - I;
                    need = !(den > 0.f) || I * (1.f + 1e-5f) >= thr_lo * den;
                }
                if (need) {
                    const int q = atomicAdd(&qn_s, 1);
                    if (q < F2_QCAP) {
                        S_queue[q] = (zero ? 0x80000000u : 0u) | ((unsigned)k << 16) | (unsigned)g;
                        S_qscore[q] = per_class ? __fmul_rn(__ldg(myrow + c), obj) : s1;
                    }
                }
            }
        }
    };
    auto clip = [&](int qn) {
        constexpr int NGROUP = NT / 8;
        if (qn <= F2_COOP) {
            for (int q = k; q < qn; q += NT) {
                const unsigned e = S_queue[q];
                if (e >> 31) { S_qiou[q] = 0.0; continue; }
                f2_box_features(p.corners + ((size_t)s * K + ((e >> 16) & 0x3fffu)) * 24, &S_cbox[q]);
            }
            __syncthreads();
            const int gli = lane & 7, gshift = lane & 24, group = k >> 3;
            V2<double> *gbuf = S_scratch + group * 8;
            for (int base = 0; base < qn; base += NGROUP) {
                if (base + warp * 4 >= qn) break;   // warp-uniform
                const int qi = base + group;
                unsigned e = 0x80000000u;
                if (qi < qn) e = S_queue[qi];
                const bool act = !(e >> 31);
                const AmBox &a = S_cbox[act ? qi : 0], &b = S_gbox[act ? (e & 0xffffu) : 0];
                CoopState st{(double)a.qx[gli & 3], (double)a.qz[gli & 3], act ? 4 : 0};
                for (int e = 0; e < 4; ++e) st = f2_coop_pass(&b, e, st, gli, gshift, gbuf);   // clip edges (3->0), (0->1), (1->2), (2->3)
                const double ia = coop_area_f64(st.vx, st.vy, st.n, gli);
                if (act && gli == 0) S_qiou[qi] = am_finish_iou(ia, a, b);
            }
        } else if (k < AM_CLIP) {
            V2<double> *bufA = S_scratch + k, *bufB = S_scratch + SH_MAXV * AM_CLIP + k;
            for (int qi = k; qi < qn; qi += AM_CLIP) {
                const unsigned e = S_queue[qi];
                if (e >> 31) { S_qiou[qi] = 0.0; continue; }
                S_qiou[qi] = f2_serial_iou(p.corners + ((size_t)s * K + ((e >> 16) & 0x3fffu)) * 24, &S_gbox[e & 0xffffu], bufA, bufB);
            }
        }
    };
    // pass 0: claims -- the highest-scoring detection whose first-max GT (within the class) it is wins the GT;
    // pass 1: a candidate is a true positive at threshold t iff it holds the claim (eval_det.py:117-140)
    auto claims = [&](int qn, int pass, bool recheck) {
        for (int q = k; q < qn; q += NT) {
            const double v = S_qiou[q];
            if (!(v > thr_min)) continue;
            const unsigned e = S_queue[q];
            const int i = (int)((e >> 16) & 0x3fffu), g = (int)(e & 0xffffu);
            const int c = S_glab[g];
            if (pass == 0 || recheck) {   // (slab mode re-enumerates for the second sweep: the remembered flag is gone)
                bool first_max = true;   // jmax of (det, class c): first GT of the class attaining the maximum (eval_det.py:121-126)
                for (int q2 = 0; q2 < qn; ++q2) {
                    const unsigned e2 = S_queue[q2];
                    if ((int)((e2 >> 16) & 0x3fffu) != i) continue;
                    const int g2 = (int)(e2 & 0xffffu);
                    if (S_glab[g2] != c) continue;
                    const double v2 = S_qiou[q2];
                    if (v2 > v || (v2 == v && g2 < g)) { first_max = false; break; }
                }
                if (!first_max) continue;
                if (pass == 0) S_queue[q] = e | 0x40000000u;   // remembered for the true-positive pass on the same queue
            }
            if (pass == 0) {
                // non-negative fp32 scores order like their bit patterns; lower det index wins ties
                const unsigned long long key = ((unsigned long long)__float_as_uint(S_qscore[q]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
                for (int t = 0; t < p.nthr; ++t)
                    if (v > p.thr[t]) atomicMax(&S_best[(size_t)g * p.nthr + t], key);
            } else {
                if (!recheck && !(e & 0x40000000u)) continue;
                unsigned char tp = 0;
                for (int t = 0; t < p.nthr; ++t)
                    if (v > p.thr[t]) {
                        const unsigned long long w = S_best[(size_t)g * p.nthr + t];
                        if ((unsigned)(0xFFFFFFFFu - (unsigned)(w & 0xFFFFFFFFull)) == (unsigned)i) tp |= (unsigned char)(1u << t);
                    }
                if (tp) {
                    const size_t dslot = (size_t)s * K + i;
                    if (p.rec_tp) p.rec_tp[(size_t)c * N + dslot] = tp;
                    if (p.tp_key) {
                        const int at = atomicAdd(&p.tp_cnt[c], 1);
                        if (at < p.tp_cap) { p.tp_key[(size_t)c * p.tp_cap + at] = score_key(S_qscore[q]); p.tp_bits[(size_t)c * p.tp_cap + at] = tp; }
                    }
                }
            }
        }
    };
    // One code instance of each stage.  Usual case: everything in one go (one "slab" = all boxes, claims then true
    // positives back to back).  Crowded scene / negative threshold (more candidates than the queue holds): restart in
    // slabs of boxes that cannot produce more than F2_QCAP candidates each -- a detection's candidates never straddle
    // slabs (first-max is per detection), the claims accumulate across slabs, true positives in a second sweep.
    bool multi = false;
    int ds = K;
    for (int sweep = 0; sweep < 2; ++sweep) {
        for (int k0 = 0; k0 < K; k0 += ds) {
            if (multi || sweep || k0) { __syncthreads(); if (k == 0) qn_s = 0; __syncthreads(); }
            enumerate(k0, k0 + ds);
            __syncthreads();
            const int qn = qn_s;
            F2STAMP(3);
            if (qn > F2_QCAP) { multi = true; ds = max(1, F2_QCAP / ng); k0 = -ds; continue; }   // only in the very first pass
            clip(qn);
            __syncthreads();
            F2STAMP(4);
            for (int pass = multi ? sweep : 0; pass <= (multi ? sweep : 1); ++pass) {
                claims(qn, pass, multi);
                if (!multi && pass == 0) __syncthreads();
            }
            if (p.dbg && k == 0) p.dbg[(size_t)blockIdx.x * 16 + 6] = (unsigned long long)qn;
        }
        if (!multi) break;
    }
    F2STAMP(5);
}

}  // namespace ovdet

using namespace ovdet;

extern "C" int ovdet_ap_front_f32(const float *corners, const float *probs, const float *obj, const uint8_t *nonempty,
                                  const float *gt_corners, const int64_t *gt_labels, const void *gt_present,
                                  int S, int K, int G, int C, double nms_iou, float conf_thresh, unsigned flags,
                                  const double *thr, int nthr, double *iou_ws, float *rec_score, uint8_t *rec_tp, int64_t *npos,
                                  uint32_t *tp_key, uint8_t *tp_bits, int32_t *tp_cnt, int tp_cap, uint8_t *keep_out, void *stream)
{
    OVDET_REQUIRE(S >= 0 && K > 0 && G >= 0 && C > 0, "bad size");
    if (S == 0) return OVDET_OK;
    OVDET_REQUIRE(corners && probs && obj && rec_score && npos && thr, "null pointer");
    OVDET_REQUIRE(G == 0 || (gt_corners && gt_labels && gt_present), "null GT pointer");
    OVDET_REQUIRE(nthr >= 1 && nthr <= 8, "1..8 thresholds");
    OVDET_REQUIRE((tp_key == nullptr) == (tp_bits == nullptr) && (tp_key == nullptr) == (tp_cnt == nullptr), "tp_key, tp_bits and tp_cnt go together");
    OVDET_REQUIRE(tp_key == nullptr || tp_cap > 0, "tp_cap must be positive");
    OVDET_REQUIRE(rec_tp || tp_key, "need rec_tp and/or a TP list to report the true positives");
    if (flags & OVDET_FRONT_RESET) {   // start of an evaluation: the lists' counters and the GT counts, one memset each
        char *lo = reinterpret_cast<char *>(tp_cnt), *hi = reinterpret_cast<char *>(npos);
        const size_t gap = sizeof(int32_t) * (size_t)((C + 1) / 2 * 2);
        if (tp_cnt && hi == lo + gap)   // npos right behind tp_cnt (8-byte aligned), as the host shim lays them out: one memset
            OVDET_CUDA_TRY(cudaMemsetAsync(lo, 0, gap + sizeof(int64_t) * C, reinterpret_cast<cudaStream_t>(stream)));
        else {
            if (tp_cnt) OVDET_CUDA_TRY(cudaMemsetAsync(tp_cnt, 0, sizeof(int32_t) * C, reinterpret_cast<cudaStream_t>(stream)));
            OVDET_CUDA_TRY(cudaMemsetAsync(npos, 0, sizeof(int64_t) * C, reinterpret_cast<cudaStream_t>(stream)));
        }
        flags &= ~OVDET_FRONT_RESET;
    }
    const int nt = K <= 128 ? 128 : 256;
    const bool lean = K <= 256 && G <= 32767 && ((flags & OVDET_PARSE_NO_NMS) || nms_iou >= 0.0) && !getenv("OVDET_APFRONT_GENERIC") &&
                      f2_smem_bytes(K, G, C, nthr, nt, flags) <= 100 * 1024;
    if (!lean)   // K > 256, a negative NMS threshold, very many classes: nms_core + am_scene_body (eval.cu)
        return front1_launch(corners, probs, obj, nonempty, gt_corners, gt_labels, gt_present, S, K, G, C, nms_iou, conf_thresh, flags,
                             thr, nthr, iou_ws, rec_score, rec_tp, npos, tp_key, tp_bits, tp_cnt, tp_cap, keep_out, stream);
    Front2Params p;
    p.corners = corners; p.probs = probs; p.obj = obj; p.nonempty = nonempty; p.gt_corners = gt_corners; p.gt_labels = gt_labels;
    p.gt_present = (flags & OVDET_FRONT_GT_PRESENT_F32) ? nullptr : static_cast<const uint8_t *>(gt_present);
    p.gt_present_f32 = (flags & OVDET_FRONT_GT_PRESENT_F32) ? static_cast<const float *>(gt_present) : nullptr;
    p.S = S; p.K = K; p.G = G; p.C = C; p.nthr = nthr;
    for (int t = 0; t < nthr; ++t) p.thr[t] = thr[t];
    p.nms_iou = nms_iou; p.conf = conf_thresh; p.flags = flags;
    p.rec_score = rec_score; p.rec_tp = rec_tp; p.npos = reinterpret_cast<unsigned long long *>(npos);
    p.tp_key = tp_key; p.tp_bits = tp_bits; p.tp_cnt = tp_cnt; p.tp_cap = tp_cap; p.keep_out = keep_out;
    { const char *e = getenv("OVDET_APFRONT_DBG_PTR"); p.dbg = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0)) : nullptr; }
    const size_t smem = f2_smem_bytes(K, G, C, nthr, nt, flags);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (nt == 128) {
        OVDET_CUDA_TRY(ensure_dyn_smem(ap_front2_kernel<128>, smem));
        ap_front2_kernel<128><<<S, 128, smem, st>>>(p);
    } else {
        OVDET_CUDA_TRY(ensure_dyn_smem(ap_front2_kernel<256>, smem));
        ap_front2_kernel<256><<<S, 256, smem, st>>>(p);
    }
    return launch_ok("ap_front2_kernel");
}
