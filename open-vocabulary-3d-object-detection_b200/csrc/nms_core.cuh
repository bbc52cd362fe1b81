// nms_core.cuh -- the per-scene greedy NMS core (shared by nms.cu and the fused AP front end, eval.cu).
// See nms.cu for the algorithm notes and the reference lines it replaces.
#pragma once
#include <math.h>

#include "common.cuh"

namespace ovdet {

#define NSTAMP(ptr, i) do { if ((ptr) && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); (ptr)[(size_t)blockIdx.x * 16 + (i)] = t_; } } while (0)

// compare-and-select min / max for the fp64 overlap tests: the library's fmin / fmax spend ~10 instructions per call on
// NaN handling (ncu: 14 % of the NMS kernels' instructions); these differ from them only for NaN operands and the sign of
// a zero result, neither of which can turn an overlap test (o > thr, e == 0) around.
__device__ __forceinline__ double dmin_cs(double x, double y) { return x < y ? x : y; }
__device__ __forceinline__ double dmax_cs(double x, double y) { return x > y ? x : y; }

constexpr int NMS_NT = 256;
constexpr int NMS_MAXK = 1024;
constexpr int NMS_MAXCLS = 256;   // class ids 0..255 take the per-class path; anything else the generic mask path

struct NmsSmem {
    double *lo[3], *hi[3], *vol, *cls, *skey;
    int *sidx;
    uint32_t *mask;
    int *misc;  // [0]=n_alive, [1]=npick, [2]=class ids not dense
    unsigned char *picked;  // by sorted position
    int *ccnt, *cstart;                  // [NMS_MAXCLS] per-class counts / segment starts
    unsigned short *grouped, *crank;     // [K] sorted positions grouped by class / rank of a position inside its class
};

__host__ __device__ inline int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

__host__ __device__ inline size_t nms_smem_bytes(int K)
{
    const int Kp = next_pow2(K < 32 ? 32 : K);
    const int W = (K + 31) / 32;
    size_t b = sizeof(double) * (size_t)(8 * K + Kp);  // lo3 hi3 vol cls + skey
    b += sizeof(int) * (size_t)Kp;                     // sidx
    b += sizeof(uint32_t) * (size_t)K * W;             // mask
    b += sizeof(int) * 4 + (size_t)((K + 3) & ~3) + 16;
    b += sizeof(int) * 2 * NMS_MAXCLS + sizeof(unsigned short) * 2 * (size_t)((K + 1) & ~1);
    return (b + 15) & ~(size_t)15;
}

__device__ inline NmsSmem nms_carve(unsigned char *base, int K)
{
    const int Kp = next_pow2(K < 32 ? 32 : K);
    const int W = (K + 31) / 32;
    NmsSmem s;
    double *d = reinterpret_cast<double *>(base);
    for (int a = 0; a < 3; ++a) { s.lo[a] = d; d += K; s.hi[a] = d; d += K; }
    s.vol = d; d += K;
    s.cls = d; d += K;
    s.skey = d; d += Kp;
    s.sidx = reinterpret_cast<int *>(d);
    s.mask = reinterpret_cast<uint32_t *>(s.sidx + Kp);
    s.misc = reinterpret_cast<int *>(s.mask + (size_t)K * W);
    s.picked = reinterpret_cast<unsigned char *>(s.misc + 4);
    s.ccnt = reinterpret_cast<int *>(s.picked + ((K + 3) & ~3));
    s.cstart = s.ccnt + NMS_MAXCLS;
    s.grouped = reinterpret_cast<unsigned short *>(s.cstart + NMS_MAXCLS);
    s.crank = s.grouped + ((K + 1) & ~1);
    return s;
}

// Src provides: bool alive(k); double score(k); double cls_of(k); void box(k, lo[3], hi[3], double &cls)
// On return: s.picked[pos] for positions pos < s.misc[0], s.sidx[pos] & 0x3fffffff = original index (bit 30 = dead),
// s.misc[1] = npick.  `order_out` (may be null) gets original indices in pick order.
template <typename Src>
__device__ void nms_core(const Src &src, int K, int dims, bool samecls, bool old_type, double thr, double eps,
                         NmsSmem &s, int32_t *order_out, unsigned long long *dbg = nullptr, bool lhs = false)
{
    using A = Ar<double>;
    const int tid = threadIdx.x;
    const int NT = blockDim.x;   // 128 for K <= 128 (more scenes in flight per SM), else 256
    const int Kp = next_pow2(K < 32 ? 32 : K);
    const int W = (K + 31) / 32;
    const int warp = tid >> 5, lane = tid & 31, nw = NT / 32;
    const bool fast = thr >= 0.0;
    // "a before b" = alive first, higher score, then higher index (stable-ascending-from-the-end)
    auto before = [](double ka, int ia, double kb, int ib) {
        const bool da = ia & 0x40000000, db = ib & 0x40000000;
        if (da != db) return !da;
        if (ka != kb) return ka > kb;
        return ia > ib;
    };
    // ---- 0. mode.  Class-wise NMS with thr >= 0 and dense class ids (0..255) splits into independent per-class
    // problems (steps 3'/4'); if the caller does not need the global pick order either, no global sort is needed at all:
    // boxes stay at their original positions and are ordered inside their class only (`identity`).
    bool classwise = fast && samecls && !lhs;   // the lhs re-pick needs the global score order of the suppressed boxes
    if (classwise) {
        bool bad = false;
        for (int k = tid; k < K; k += NT)
            if (src.alive(k)) { const double cd = src.cls_of(k); if (!(cd >= 0.0 && cd < (double)NMS_MAXCLS && (double)(int)cd == cd)) bad = true; }
        classwise = !__syncthreads_or(bad);
    }
    const bool identity = classwise && order_out == nullptr;
    if (identity) {
        for (int k = tid; k < K; k += NT) {
            const bool al = src.alive(k);
            s.skey[k] = al ? src.score(k) : -INFINITY;
            s.sidx[k] = al ? k : (k | 0x40000000);
        }
        __syncthreads();
    } else {
        // ---- 1. bitonic sort of (score, index) in shared memory
        for (int k = tid; k < Kp; k += NT) {
            const bool al = k < K && src.alive(k);
            s.skey[k] = al ? src.score(k) : -INFINITY;
            s.sidx[k] = al ? k : (k | 0x40000000);  // dead entries sort last
        }
        __syncthreads();
        for (int size = 2; size <= Kp; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < Kp / 2; t += NT) {
                    const int lo = 2 * t - (t & (stride - 1));
                    const int hi = lo + stride;
                    const bool up = ((lo & size) == 0);
                    const double ka = s.skey[lo], kb = s.skey[hi];
                    const int ia = s.sidx[lo], ib = s.sidx[hi];
                    const bool swap = up ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
                    if (swap) { s.skey[lo] = kb; s.skey[hi] = ka; s.sidx[lo] = ib; s.sidx[hi] = ia; }
                }
                __syncthreads();
            }
        }
    }
    NSTAMP(dbg, 2);
    // ---- 2. gather box extents at their (sorted or original) positions
    if (tid == 0) { s.misc[0] = 0; s.misc[1] = 0; }
    __syncthreads();
    int local_alive = 0;
    for (int pos = tid; pos < K; pos += NT) {
        const int k = s.sidx[pos];
        s.picked[pos] = 0;
        if (k & 0x40000000) continue;
        ++local_alive;
        double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, cl = 0;
        src.box(k, lo, hi, cl);
        double v = A::sub(hi[0], lo[0]);
        for (int a = 1; a < dims; ++a) v = A::mul(v, A::sub(hi[a], lo[a]));
        for (int a = 0; a < 3; ++a) { s.lo[a][pos] = lo[a]; s.hi[a][pos] = hi[a]; }
        s.vol[pos] = A::add(v, eps);
        s.cls[pos] = cl;
    }
    if (local_alive && !identity) atomicAdd(&s.misc[0], local_alive);
    if (identity && tid == 0) s.misc[0] = K;
    __syncthreads();
    const int n = s.misc[0];   // positions to consider: the alive prefix (sorted) or all K with holes (identity)
    NSTAMP(dbg, 3);
    // ---- 3'/4'. class-wise NMS with a non-negative threshold: independent per-class greedy loops
    if (classwise) {
        for (int c = tid; c < NMS_MAXCLS; c += NT) s.ccnt[c] = 0;
        if (tid == 0) s.misc[2] = 0;
        __syncthreads();
        if (warp == 0) {   // rank of every alive position inside its class, in position order
            for (int base = 0; base < n; base += 32) {
                const int pos = base + lane;
                int c = -1 - lane;   // dead / past the end: unique dummies
                if (pos < n && !(s.sidx[pos] & 0x40000000)) c = (int)s.cls[pos];
                const unsigned m = __match_any_sync(0xffffffffu, c);
                if (c >= 0) s.crank[pos] = (unsigned short)(s.ccnt[c] + __popc(m & ((1u << lane) - 1)));
                __syncwarp();
                if (c >= 0 && (m & ((1u << lane) - 1)) == 0) s.ccnt[c] += __popc(m);   // lowest lane of each class group
                __syncwarp();
            }
            // exclusive scan of the class counts (NMS_MAXCLS / 32 per lane)
            constexpr int PER = NMS_MAXCLS / 32;
            int loc[PER], sum = 0;
#pragma unroll
            for (int q = 0; q < PER; ++q) { loc[q] = s.ccnt[lane * PER + q]; sum += loc[q]; }
            int incl = sum;
            for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
            int run = incl - sum;
            int hi_c = -1; bool big = false;
#pragma unroll
            for (int q = 0; q < PER; ++q) { s.cstart[lane * PER + q] = run; run += loc[q]; if (loc[q] > 0) hi_c = lane * PER + q; if (loc[q] > 32) big = true; }
            for (int off = 16; off > 0; off >>= 1) hi_c = max(hi_c, __shfl_xor_sync(0xffffffffu, hi_c, off));
            big = __any_sync(0xffffffffu, big);
            if (lane == 0) { s.misc[3] = hi_c + 1; s.misc[2] = big ? 2 : 0; }   // class-loop bound; some class needs the warp scan
        }
        __syncthreads();
        const int ncls = s.misc[3];
        const bool any_big = s.misc[2] & 2;
        for (int pos = tid; pos < n; pos += NT)
            if (!(s.sidx[pos] & 0x40000000)) s.grouped[s.cstart[(int)s.cls[pos]] + s.crank[pos]] = (unsigned short)pos;
        __syncthreads();
        if (identity) {   // order the members of each class by score: rank = number of classmates that come before
            for (int pos = tid; pos < n; pos += NT) {
                const int k = s.sidx[pos];
                if (k & 0x40000000) continue;
                const int c = (int)s.cls[pos];
                const int q0 = s.cstart[c], nc = s.ccnt[c];
                const double key = s.skey[pos];
                int r = 0;
                for (int m = 0; m < nc; ++m) { const int p2 = s.grouped[q0 + m]; r += before(s.skey[p2], s.sidx[p2], key, k) ? 1 : 0; }
                s.crank[pos] = (unsigned short)r;
            }
            __syncthreads();
            for (int pos = tid; pos < n; pos += NT)
                if (!(s.sidx[pos] & 0x40000000)) s.grouped[s.cstart[(int)s.cls[pos]] + s.crank[pos]] = (unsigned short)pos;
            __syncthreads();
        }
        NSTAMP(dbg, 4);
        // (a) thread per box: its suppression words over the LATER boxes of its own class (class-local bit index);
        //     sum_c n_c^2/2 pair tests in total, rows independent of each other.
        // (b) the greedy pick per class is then an integer scan over those words: one thread per class when the
        //     class fits one word (the common case), else one warp with the removed bitset one word per lane.
        for (int pos = tid; pos < n; pos += NT) {
            if (s.sidx[pos] & 0x40000000) continue;
            const int i = pos;
            const int c = (int)s.cls[i];
            const int q0 = s.cstart[c], nc = s.ccnt[c];
            const unsigned short *g = s.grouped + q0;
            const int ii = s.crank[pos];
            const int q = q0 + ii;
            double li[3], hi_[3];
            for (int a = 0; a < 3; ++a) { li[a] = s.lo[a][i]; hi_[a] = s.hi[a][i]; }
            const double vi = s.vol[i];
            uint32_t word = 0;
            for (int w = 0; w < (ii >> 5); ++w) s.mask[(size_t)q * W + w] = 0u;
            for (int jj = ii + 1; jj < nc; ++jj) {
                if ((jj & 31) == 0) { s.mask[(size_t)q * W + ((jj - 1) >> 5)] = word; word = 0; }
                const int j = g[jj];
                double e[3] = {1.0, 1.0, 1.0};
                for (int a = 0; a < dims; ++a) e[a] = dmax_cs(0.0, A::sub(dmin_cs(hi_[a], s.hi[a][j]), dmax_cs(li[a], s.lo[a][j])));
                if (e[0] == 0.0 || e[1] == 0.0 || (dims == 3 && e[2] == 0.0)) continue;   // inter == 0: o is 0 or NaN, never > thr
                double inter = e[0];
                for (int a = 1; a < dims; ++a) inter = A::mul(inter, e[a]);
                const double o = old_type ? A::div(inter, s.vol[j]) : A::div(inter, A::sub(A::add(vi, s.vol[j]), inter));
                if (o > thr) word |= 1u << (jj & 31);
            }
            s.mask[(size_t)q * W + ((nc - 1) >> 5)] = word;
        }
        __syncthreads();
        NSTAMP(dbg, 8);
        for (int c = tid; c < ncls; c += NT) {   // classes of <= 32 boxes: one thread each
            const int nc = s.ccnt[c];
            if (nc == 0 || nc > 32) continue;
            const int q0 = s.cstart[c];
            uint32_t removed = 0;
            for (int ii = 0; ii < nc; ++ii)
                if (!((removed >> ii) & 1u)) { s.picked[s.grouped[q0 + ii]] = 1; removed |= s.mask[(size_t)(q0 + ii) * W]; }
        }
        NSTAMP(dbg, 9);
        if (any_big) for (int c = warp; c < ncls; c += nw) {   // larger classes: one warp each
            const int nc = s.ccnt[c];
            if (nc <= 32) continue;
            const int q0 = s.cstart[c];
            const int Wc = (nc + 31) >> 5;
            uint32_t removed = 0;
            for (int ii = 0; ii < nc; ++ii) {
                const uint32_t word = __shfl_sync(0xffffffffu, removed, ii >> 5);
                if (!((word >> (ii & 31)) & 1u)) {
                    if (lane == 0) s.picked[s.grouped[q0 + ii]] = 1;
                    if (lane < Wc && lane >= (ii >> 5)) removed |= s.mask[(size_t)(q0 + ii) * W + lane];
                }
            }
        }
        __syncthreads();
        NSTAMP(dbg, 5);
        if (warp == 0) {   // number of picks; their original indices in score order (sorted mode only)
            int np = 0;
            for (int base = 0; base < n; base += 32) {
                const int pos = base + lane;
                const bool pk = pos < n && s.picked[pos];
                const unsigned m = __ballot_sync(0xffffffffu, pk);
                if (pk && order_out) order_out[np + __popc(m & ((1u << lane) - 1))] = s.sidx[pos];
                np += __popc(m);
            }
            if (lane == 0) s.misc[1] = np;
        }
        __syncthreads();
        return;
    }
    // ---- 3. suppression bitmask (generic path: positions < n are the alive boxes in score order)
    for (int i = warp; i < n; i += nw) {
        double li[3], hi_[3];
        for (int a = 0; a < 3; ++a) { li[a] = s.lo[a][i]; hi_[a] = s.hi[a][i]; }
        const double vi = s.vol[i], ci = s.cls[i];
        for (int w = 0; w < W; ++w) {
            const int j = 32 * w + lane;
            bool sup = false;
            if (32 * w + 31 > i) {
                if (j > i && j < n) {
                    // exact shortcuts for thr >= 0: a different class (o*0) or an empty overlap on any axis
                    // (inter == 0) gives o in {0, NaN}, never > thr -- skips the fp64 divide for most pairs
                    bool maybe = !(fast && samecls && ci != s.cls[j]);
                    double e[3] = {1.0, 1.0, 1.0};
                    if (maybe) {
                        for (int a = 0; a < dims; ++a) e[a] = dmax_cs(0.0, A::sub(dmin_cs(hi_[a], s.hi[a][j]), dmax_cs(li[a], s.lo[a][j])));
                        if (fast && (e[0] == 0.0 || e[1] == 0.0 || (dims == 3 && e[2] == 0.0))) maybe = false;
                    }
                    if (maybe) {
                        double inter = e[0];
                        for (int a = 1; a < dims; ++a) inter = A::mul(inter, e[a]);
                        double o = old_type ? A::div(inter, s.vol[j]) : A::div(inter, A::sub(A::add(vi, s.vol[j]), inter));
                        if (samecls) o = A::mul(o, ci == s.cls[j] ? 1.0 : 0.0);
                        sup = o > thr;
                    }
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, sup);
            if (lane == 0) s.mask[(size_t)i * W + w] = m;
        }
    }
    __syncthreads();
    // ---- 4. ordered scan by warp 0
    if (warp == 0) {
        uint32_t removed = 0;
        int np = 0;
        for (int i = 0; i < n; ++i) {
            const uint32_t word = __shfl_sync(0xffffffffu, removed, i >> 5);
            if (!((word >> (i & 31)) & 1u)) {
                if (lane == 0) {
                    s.picked[i] = 1;
                    if (order_out) order_out[np] = s.sidx[i];
                }
                ++np;
                const uint32_t row = lane < W ? s.mask[(size_t)i * W + lane] : 0u;
                if (lhs) {
                    // tools variant (3DOVDet_tools/utils/box_3d_utils.py:113-116): the better-scoring half of the boxes this
                    // pick suppresses is picked too (in descending score = ascending position), and still leaves the race
                    const uint32_t newly = row & ~removed;
                    const int cnt = __popc(newly);
                    int incl = cnt;
                    for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += y; }
                    const int half = __shfl_sync(0xffffffffu, incl, 31) / 2;
                    int rank = incl - cnt;
                    uint32_t m = newly;
                    while (m && rank < half) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        const int pos = 32 * lane + b;
                        s.picked[pos] = 1;
                        if (order_out) order_out[np + rank] = s.sidx[pos];
                        ++rank;
                    }
                    np += half;
                }
                removed |= row;
            }
        }
        if (lane == 0) s.misc[1] = np;
    }
    __syncthreads();
}

struct CornerSrc {
    const float *corners; const float *obj; const uint8_t *nonempty; const int *cls; int dims2d;
    __device__ bool alive(int k) const { return nonempty ? nonempty[k] != 0 : true; }
    __device__ double score(int k) const { return (double)obj[k]; }
    __device__ double cls_of(int k) const { return (double)cls[k]; }
    __device__ void box(int k, double *lo, double *hi, double &cl) const
    {
        const float *c = corners + (size_t)k * 24;
        float mn[3], mx[3];
        for (int a = 0; a < 3; ++a) {
            mn[a] = mx[a] = __ldg(c + a);
            for (int i = 1; i < 8; ++i) { const float v = __ldg(c + 3 * i + a); mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
        }
        if (dims2d) {  // ap_calculator.py:92-104: x and z extents
            lo[0] = mn[0]; lo[1] = mn[2]; hi[0] = mx[0]; hi[1] = mx[2];
        } else {
            for (int a = 0; a < 3; ++a) { lo[a] = mn[a]; hi[a] = mx[a]; }
        }
        cl = (double)cls[k];
    }
};

}  // namespace ovdet
