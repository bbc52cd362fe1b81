// ap_compact.cu -- AP from (score, tp) records WITHOUT a global sort.
//
// VOC AP (utils/eval_det.py:23-54) only looks at the precision at the true-positive
// records: recall changes exactly there, and the running-max envelope of the
// precision is attained there.  So instead of sorting every record of a class by
// score (utils/eval_det.py:108-111; csrc/eval.cu's radix sort), it is enough to know,
// for each TP record, its 1-based position in the sorted order = the number of
// present records with a score >= its own:
//   1. collect   the records with any TP bit set into a small per-class list
//   2. sort      that list by descending score (bitonic, shared memory, <= 16384 entries)
//   3. hist      ONE streaming pass over all records: binary-search each score in the
//                sorted TP list, bump that bucket (shared-memory privatised histogram)
//   4. final     prefix-sum the buckets -> positions; cumulative TP, precision envelope, AP
// Traffic: the record stream is read twice (5 B/record) instead of ~10 radix passes.
// The stages are separate entry points because a scene-sharded multi-GPU caller
// exchanges between them: all-gather of the TP lists after (1), all-reduce of the
// bucket histogram after (3) -- KBs instead of the whole record stream.
// Results equal the sorted formulation on tie-free scores (same caveat as the reference's
// unstable argsort).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace ovdet {

constexpr int APC_NT = 256;
constexpr int APC_MAXCAP = 16384;

__device__ __forceinline__ uint32_t apc_score_key(float s)
{   // ascending key order == descending score; -inf (absent) maps to the largest finite-score key + ...
    const uint32_t b = __float_as_uint(s);
    const uint32_t ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~ord;
}

// 16 records per thread: four 16-byte score loads and one 16-byte tp load in flight at once (the pass is a pure stream:
// 5 B per record), valid records counted in registers, the rare TP records appended with one atomic each.
constexpr int APC_RPT = 16;

__global__ void __launch_bounds__(APC_NT) apc_collect_kernel(const float *__restrict__ score, const uint8_t *__restrict__ tp,
                                                             long long N, int cap, uint32_t *tp_key, uint8_t *tp_bits,
                                                             int *tp_cnt, unsigned long long *nvalid)
{
    const int c = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const float *sc = score + (size_t)c * N;
    const uint8_t *tb = tp + (size_t)c * N;
    const bool vec = ((reinterpret_cast<uintptr_t>(sc) & 15) == 0) && ((reinterpret_cast<uintptr_t>(tb) & 15) == 0);
    int local_valid = 0;
    auto emit = [&](float s, uint8_t t) {
        const int slot = atomicAdd(&tp_cnt[c], 1);
        if (slot < cap) { tp_key[(size_t)c * cap + slot] = apc_score_key(s); tp_bits[(size_t)c * cap + slot] = t; }
    };
    for (long long base = ((long long)blockIdx.x * APC_NT + threadIdx.x) * APC_RPT; base < N; base += (long long)gridDim.x * APC_NT * APC_RPT) {
        if (vec && base + APC_RPT <= N) {
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const float4 *>(sc + base) + q);
            const uint4 tv = __ldg(reinterpret_cast<const uint4 *>(tb + base));
            const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float f[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const bool present = f[r] > -INFINITY;
                    local_valid += present ? 1 : 0;
                    const uint8_t t = (uint8_t)((tw[q] >> (8 * r)) & 0xffu);
                    if (present && t) emit(f[r], t);
                }
            }
        } else {
            for (int r = 0; r < APC_RPT && base + r < N; ++r) {
                const float s = sc[base + r];
                const uint8_t t = tb[base + r];
                const bool present = s > -INFINITY;
                local_valid += present ? 1 : 0;
                if (present && t) emit(s, t);
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) local_valid += __shfl_xor_sync(0xffffffffu, local_valid, off);
    if (lane == 0 && local_valid) atomicAdd(&nvalid[c], (unsigned long long)local_valid);
}

// per class: bitonic sort of `cap` (power of two) (key, bits) entries, ascending key; unused slots hold key 0xFFFFFFFF
__global__ void __launch_bounds__(1024) apc_sort_kernel(uint32_t *tp_key, uint8_t *tp_bits, int cap)
{
    extern __shared__ __align__(16) unsigned char sm[];
    uint32_t *k = reinterpret_cast<uint32_t *>(sm);
    uint8_t *b = reinterpret_cast<uint8_t *>(k + cap);
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < cap; i += 1024) { k[i] = tp_key[(size_t)c * cap + i]; b[i] = tp_bits[(size_t)c * cap + i]; }
    __syncthreads();
    for (int size = 2; size <= cap; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < cap / 2; t += 1024) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const uint32_t ka = k[lo], kb = k[hi];
                if (up ? (ka > kb) : (ka < kb)) {
                    k[lo] = kb; k[hi] = ka;
                    const uint8_t ba = b[lo]; b[lo] = b[hi]; b[hi] = ba;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < cap; i += 1024) { tp_key[(size_t)c * cap + i] = k[i]; tp_bits[(size_t)c * cap + i] = b[i]; }
}

// ---- bucket of a record = number of TP-list keys strictly below its key (lower_bound over the sorted TP list).
// A binary search costs log2(cap) dependent shared loads per record (ncu: 33 short-scoreboard stalls per issue).
// Instead the KEY axis between the smallest and the largest key of the list is cut into <= APC_BINS equal bins
// (bin = (key - kmin) >> shift; the key is the order-preserving image of the fp32 score, i.e. a piecewise-logarithmic
// scale, which spreads the skewed score distributions of a detector evenly -- uniform bins in SCORE left ~10 entries
// per occupied bin); edge[b] = lower_bound(kmin + (b << shift)) is computed once per class (apc_edges_kernel) and a
// record only searches entries edge[b] .. edge[b+1].
constexpr int APC_BINS = 4096;
constexpr int APC_EHDR = 8;                           // u16 slots of the per-class header: kmin (2), shift, nb, ntp
constexpr int APC_ESTRIDE = APC_BINS + 1 + APC_EHDR;

__global__ void __launch_bounds__(1024) apc_edges_kernel(const uint32_t *__restrict__ tp_key, int cap, uint16_t *edge)
{
    extern __shared__ __align__(16) unsigned char sm[];
    uint32_t *k = reinterpret_cast<uint32_t *>(sm);
    __shared__ uint32_t kmin_s; __shared__ int shift_s, nb_s, ntp_s;
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < cap; i += 1024) k[i] = tp_key[(size_t)c * cap + i];
    __syncthreads();
    if (threadIdx.x == 0) {
        int lo = 0, n = cap;   // number of real entries = first empty (0xFFFFFFFF) slot
        while (n > 1) { const int half = n >> 1; lo += (k[lo + half - 1] < 0xFFFFFFFFu) ? half : 0; n -= half; }
        const int ntp = lo + ((k[lo] < 0xFFFFFFFFu) ? 1 : 0);
        uint32_t kmin = 0; int shift = 0, nb = 0;
        if (ntp > 0) {
            kmin = k[0];
            const uint32_t span = k[ntp - 1] - kmin;
            while ((span >> shift) >= (uint32_t)APC_BINS) ++shift;
            nb = (int)(span >> shift) + 1;
        }
        kmin_s = kmin; shift_s = shift; nb_s = nb; ntp_s = ntp;
        uint16_t *hdr = edge + (size_t)c * APC_ESTRIDE;
        hdr[0] = (uint16_t)(kmin & 0xffffu); hdr[1] = (uint16_t)(kmin >> 16); hdr[2] = (uint16_t)shift; hdr[3] = (uint16_t)nb; hdr[4] = (uint16_t)ntp;
    }
    __syncthreads();
    const int nb = nb_s, ntp = ntp_s, shift = shift_s;
    const uint32_t kmin = kmin_s;
    uint16_t *e = edge + (size_t)c * APC_ESTRIDE + APC_EHDR;
    for (int b = threadIdx.x; b <= nb; b += 1024) {
        const unsigned long long bound = (unsigned long long)kmin + ((unsigned long long)b << shift);
        int lo = 0, hi = ntp;   // entries with key < bound
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((unsigned long long)k[mid] < bound) lo = mid + 1; else hi = mid; }
        e[b] = (uint16_t)lo;
    }
}

__global__ void __launch_bounds__(APC_NT) apc_hist_kernel(const float *__restrict__ score, long long N, const uint32_t *__restrict__ tp_key,
                                                          const uint16_t *__restrict__ edge, int cap, uint32_t *hist)
{
    extern __shared__ __align__(16) unsigned char sm[];
    uint32_t *k = reinterpret_cast<uint32_t *>(sm);   // [cap]
    uint32_t *h = k + cap;                            // [cap + 1]
    uint16_t *e = reinterpret_cast<uint16_t *>(h + cap + 1);   // [APC_BINS + 1]
    const int c = blockIdx.y;
    const uint16_t *eg = edge + (size_t)c * APC_ESTRIDE;
    const uint32_t kmin = (uint32_t)eg[0] | ((uint32_t)eg[1] << 16);
    const int shift = eg[2], nb = eg[3], ntp = eg[4];
    for (int i = threadIdx.x; i < cap; i += APC_NT) k[i] = tp_key[(size_t)c * cap + i];
    for (int i = threadIdx.x; i <= cap; i += APC_NT) h[i] = 0;
    for (int i = threadIdx.x; i <= nb; i += APC_NT) e[i] = eg[APC_EHDR + i];
    __syncthreads();
    // Every record scoring below all TPs lands in the one bucket after the last real entry (the bulk of the false
    // positives): count those in a register instead of hammering one shared word.
    const uint32_t kmax = ntp > 0 ? k[ntp - 1] : 0u;
    unsigned int tail = 0;
    const float *sc = score + (size_t)c * N;
    auto place = [&](float s) {
        if (!(s > -INFINITY)) return;
        const uint32_t key = apc_score_key(s);
        if (ntp == 0 || key > kmax) { ++tail; return; }
        int lo = 0;
        if (key > kmin) {
            const int b = (int)((key - kmin) >> shift);
            lo = e[b];
            int hi = e[b + 1];
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (k[mid] < key) lo = mid + 1; else hi = mid; }
        }
        atomicAdd(&h[lo], 1u);
    };
    // four records per thread and step (one 16-byte load), 32-bit indexing when the class fits
    if (((reinterpret_cast<uintptr_t>(sc) & 15) == 0) && N < 0x7fffffffLL) {
        const int n4 = (int)(N >> 2);
        const int step = (int)(gridDim.x * APC_NT);
        for (int i = (int)(blockIdx.x * APC_NT + threadIdx.x); i < n4; i += step) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(sc) + i);
            place(v.x); place(v.y); place(v.z); place(v.w);
        }
        for (long long i = ((long long)n4 << 2) + (long long)blockIdx.x * APC_NT + threadIdx.x; i < N; i += (long long)gridDim.x * APC_NT) place(sc[i]);
    } else {
        for (long long i = (long long)blockIdx.x * APC_NT + threadIdx.x; i < N; i += (long long)gridDim.x * APC_NT) place(sc[i]);
    }
    for (int off = 16; off > 0; off >>= 1) tail += __shfl_xor_sync(0xffffffffu, tail, off);
    if ((threadIdx.x & 31) == 0 && tail) atomicAdd(&h[ntp], tail);
    __syncthreads();
    for (int i = threadIdx.x; i <= cap; i += APC_NT) { const uint32_t v = h[i]; if (v) atomicAdd(&hist[(size_t)c * (cap + 1) + i], v); }
}

struct ApcFinalParams {
    const uint8_t *tp_bits; const int *tp_cnt; const uint32_t *hist; const long long *npos; const unsigned long long *nvalid;
    int C, cap, nthr, use07;
    double *ap, *recall; long long *ndet; int *overflow;
};

__global__ void __launch_bounds__(1024) apc_final_kernel(ApcFinalParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned int *H = reinterpret_cast<unsigned int *>(sm);   // [cap] inclusive prefix of hist = 1-based sorted position
    unsigned int *ctp = H + p.cap;                            // [cap] inclusive count of TP(t) entries
    __shared__ unsigned int wsum[32], carry_u;
    __shared__ double wmax[32], red[32], carry_max_s;
    const int c = blockIdx.x, t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cap = p.cap;
    const double npos = (double)p.npos[c];
    const double eps = 2.220446049250313e-16;
    const uint8_t *bits = p.tp_bits + (size_t)c * cap;
    const uint32_t *hist = p.hist + (size_t)c * (cap + 1);
    // two forward chunked inclusive scans (hist -> H, tp flag -> ctp)
    for (int pass = 0; pass < 2; ++pass) {
        if (tid == 0) carry_u = 0;
        __syncthreads();
        for (int base = 0; base < cap; base += 1024) {
            const int i = base + tid;
            const unsigned int v = pass == 0 ? hist[i] : (unsigned int)((bits[i] >> t) & 1u);
            unsigned int x = v;
            for (int off = 1; off < 32; off <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
            if (lane == 31) wsum[warp] = x;
            __syncthreads();
            if (tid < 32) {
                unsigned int w = wsum[tid];
                for (int off = 1; off < 32; off <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, w, off); if (tid >= off) w += y; }
                wsum[tid] = w;
            }
            __syncthreads();
            const unsigned int incl = carry_u + x + (warp ? wsum[warp - 1] : 0u);
            (pass == 0 ? H : ctp)[i] = incl;
            __syncthreads();
            if (tid == 1023) carry_u = incl;
            __syncthreads();
        }
    }
    const unsigned int total_tp = ctp[cap - 1];
    if (tid == 0) carry_max_s = 0.0;
    __syncthreads();
    double ap_local = 0.0;
    double p11[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) p11[k] = 0.0;
    for (int base = cap - 1024; base >= 0; base -= 1024) {
        const int i = base + tid;
        const bool tp = (bits[i] >> t) & 1u;
        const double ct = (double)ctp[i];
        const double prec = tp ? __ddiv_rn(ct, fmax((double)H[i], eps)) : 0.0;   // precision at this TP record
        const double rec = npos > 0.0 ? __ddiv_rn(ct, npos) : 0.0;
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
        if (tp) {
            const double rec_prev = npos > 0.0 ? __ddiv_rn(ct - 1.0, npos) : 0.0;
            ap_local += __dmul_rn(__dsub_rn(rec, rec_prev), env);
            if (p.use07) {
#pragma unroll
                for (int k = 0; k < 11; ++k) if (rec >= k * 0.1) p11[k] = fmax(p11[k], prec);
            }
        }
        __syncthreads();
        if (tid == 0) carry_max_s = fmax(carry, wmax[0]);
        __syncthreads();
    }
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
        // VOC07: max precision over records with rec >= t; attained at a TP record unless t == 0, where the very
        // first record counts too -- its precision is 1 if it is a TP (covered) else 0 (covered by the 0 init).
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
        const long long n = (long long)p.nvalid[c];
        p.ap[(size_t)t * p.C + c] = n > 0 ? res : 0.0;
        p.recall[(size_t)t * p.C + c] = (n > 0 && npos > 0.0) ? __ddiv_rn((double)total_tp, npos) : 0.0;
        if (p.ndet && t == 0) p.ndet[c] = n;
        if (p.overflow && p.tp_cnt[c] > cap) atomicExch(p.overflow, 1);
    }
}

}  // namespace ovdet

using namespace ovdet;

static bool apc_cap_ok(int cap) { return cap >= 1024 && cap <= APC_MAXCAP && (cap & (cap - 1)) == 0; }

extern "C" int ovdet_apc_collect(const float *rec_score, const uint8_t *rec_tp, int C, int64_t N, int cap,
                                 uint32_t *tp_key, uint8_t *tp_bits, int32_t *tp_cnt, int64_t *nvalid, void *stream)
{
    // a rank's own list may be short (lists of several ranks are concatenated to >= 1024 before ovdet_apc_sort)
    OVDET_REQUIRE(C > 0 && N >= 0 && cap >= 32 && cap <= APC_MAXCAP && (cap & (cap - 1)) == 0, "bad size (cap must be a power of two in [32, 16384])");
    OVDET_REQUIRE(tp_key && tp_bits && tp_cnt && nvalid, "null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    OVDET_CUDA_TRY(cudaMemsetAsync(tp_key, 0xFF, sizeof(uint32_t) * (size_t)C * cap, st));
    OVDET_CUDA_TRY(cudaMemsetAsync(tp_bits, 0, (size_t)C * cap, st));
    OVDET_CUDA_TRY(cudaMemsetAsync(tp_cnt, 0, sizeof(int32_t) * C, st));
    OVDET_CUDA_TRY(cudaMemsetAsync(nvalid, 0, sizeof(int64_t) * C, st));
    if (N == 0) return OVDET_OK;
    OVDET_REQUIRE(rec_score && rec_tp, "null pointer");
    int gx = (int)((N + APC_NT * APC_RPT - 1) / (APC_NT * APC_RPT));
    if (gx < 1) gx = 1;
    apc_collect_kernel<<<dim3(gx, C), APC_NT, 0, st>>>(rec_score, rec_tp, N, cap, tp_key, tp_bits, tp_cnt,
                                                      reinterpret_cast<unsigned long long *>(nvalid));
    return launch_ok("apc_collect_kernel");
}

extern "C" int ovdet_apc_sort(uint32_t *tp_key, uint8_t *tp_bits, int C, int cap, void *stream)
{
    OVDET_REQUIRE(C > 0 && apc_cap_ok(cap), "bad size");
    OVDET_REQUIRE(tp_key && tp_bits, "null pointer");
    const size_t smem = (size_t)cap * 5;
    OVDET_CUDA_TRY(ensure_dyn_smem(apc_sort_kernel, smem));
    apc_sort_kernel<<<C, 1024, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tp_key, tp_bits, cap);
    return launch_ok("apc_sort_kernel");
}

extern "C" int ovdet_apc_hist(const float *rec_score, int C, int64_t N, const uint32_t *tp_key, int cap, uint32_t *hist, void *stream)
{
    OVDET_REQUIRE(C > 0 && N >= 0 && apc_cap_ok(cap), "bad size");
    OVDET_REQUIRE(tp_key && hist, "null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    OVDET_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (size_t)C * (cap + 1), st));
    if (N == 0) return OVDET_OK;
    OVDET_REQUIRE(rec_score, "null pointer");
    // per-class bin edges of the sorted list, in the library's grow-only device scratch
    uint16_t *edge = nullptr;
    { void *ws = nullptr; int rc = device_scratch().acquire(sizeof(uint16_t) * (size_t)C * APC_ESTRIDE, st, &ws); if (rc) return rc; edge = static_cast<uint16_t *>(ws); }
    OVDET_CUDA_TRY(ensure_dyn_smem(apc_edges_kernel, (sizeof(uint32_t) * cap)));
    apc_edges_kernel<<<C, 1024, sizeof(uint32_t) * cap, st>>>(tp_key, cap, edge);
    const size_t smem = sizeof(uint32_t) * (2 * (size_t)cap + 1) + sizeof(uint16_t) * (APC_BINS + 2);
    OVDET_CUDA_TRY(ensure_dyn_smem(apc_hist_kernel, smem));
    // enough CTAs to fill the chip a few times over, few enough that the histogram flush stays small
    int per_sm = (int)(220 * 1024 / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int gx = (148 * per_sm + C - 1) / C;
    const long long maxgx = (N + APC_NT * 4 - 1) / (APC_NT * 4);
    if (gx > maxgx) gx = (int)maxgx;
    if (gx < 1) gx = 1;
    apc_hist_kernel<<<dim3(gx, C), APC_NT, smem, st>>>(rec_score, N, tp_key, edge, cap, hist);
    { int rc = device_scratch().release(st); if (rc) return rc; }
    return launch_ok("apc_hist_kernel");
}

extern "C" int ovdet_apc_final(const uint8_t *tp_bits, const int32_t *tp_cnt, const uint32_t *hist, const int64_t *npos,
                               const int64_t *nvalid, int C, int cap, int nthr, int use_07_metric,
                               double *ap, double *recall, int64_t *n_det, int32_t *overflow, void *stream)
{
    OVDET_REQUIRE(C > 0 && apc_cap_ok(cap) && nthr >= 1 && nthr <= 8, "bad size");
    OVDET_REQUIRE(tp_bits && tp_cnt && hist && npos && nvalid && ap && recall, "null pointer");
    ApcFinalParams p;
    p.tp_bits = tp_bits; p.tp_cnt = tp_cnt; p.hist = hist; p.npos = reinterpret_cast<const long long *>(npos);
    p.nvalid = reinterpret_cast<const unsigned long long *>(nvalid); p.C = C; p.cap = cap; p.nthr = nthr; p.use07 = use_07_metric;
    p.ap = ap; p.recall = recall; p.ndet = reinterpret_cast<long long *>(n_det); p.overflow = overflow;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (overflow) OVDET_CUDA_TRY(cudaMemsetAsync(overflow, 0, sizeof(int32_t), st));
    const size_t smem = sizeof(unsigned int) * 2 * (size_t)cap;
    OVDET_CUDA_TRY(ensure_dyn_smem(apc_final_kernel, smem));
    apc_final_kernel<<<dim3(C, nthr), 1024, smem, st>>>(p);
    return launch_ok("apc_final_kernel");
}

// =====================================================================================================================
// Scene-sharded reduction with a device-side exchange (ovdet_apx_reduce): the same four stages, but
//   * the TP lists come straight from the fused front end (ovdet_ap_front_f32), no collect pass;
//   * ranks exchange by STORING into each other's symmetric buffer over NVLink (CUDA IPC mapping) and raising a flag
//     word per (source rank, class) -- consumers spin on the flags of the class they own, so there is no collective
//     launch, no host round trip, and class c of rank r can start as soon as ITS inputs have landed;
//   * the two halves of the symmetric buffer alternate by evaluation (epoch parity), so a fast rank can never overwrite
//     what a slow rank is still reading.
// =====================================================================================================================
namespace ovdet {

constexpr int APX_MAXW = 16;
constexpr int APX_BINS = 8192;     // bins of the key axis between the merged list's extremes (u16 entry: first list index | crowded << 15)
constexpr int APX_EHDR = 4;        // u32 header words per class: kmin, shift, number of bins, list length
constexpr size_t APX_ESTRIDE_BYTES = sizeof(uint32_t) * APX_EHDR + sizeof(uint16_t) * APX_BINS;

struct ApxLayout {          // byte offsets inside one parity half of a symmetric buffer
    size_t half, flags_l, flags_h, lists, list_stride, l_cnt, l_raw, l_npos, l_key, l_bits, hist, hist_stride;
    int hp;                 // u32 pitch of a histogram row (cap_total + 1 rounded up to 4)
};
static inline size_t a16(size_t x) { return (x + 15) & ~(size_t)15; }
// Every rank ships its list of a class as R sorted RUNS; run (rank, sub) lives in list slot rank * R + sub of every
// receiver.  R > 1 spreads a rank's sort over R CTAs, but measured on one B200 (5 050 scenes, 20 classes, R = 8) the
// 160-CTA push + 160-CTA cluster merge finished the merged lists at 52-70 us against 29 us for one radix-sorting CTA per
// class: two waves of one-CTA-per-SM kernels and a system-scope fence per run cost more than the sort saves.  So R = 1.
static int apx_runs_per_rank(int W) { (void)W; return 1; }
static ApxLayout apx_layout(int C, int cap, int W)
{
    ApxLayout L;
    const int NR = W * apx_runs_per_rank(W);
    L.hp = (cap + 1 + 3) & ~3;
    size_t o = 0;
    L.flags_l = o; o += a16(sizeof(uint32_t) * (size_t)NR * C);
    L.flags_h = o; o += a16(sizeof(uint32_t) * (size_t)W * C);
    size_t s = 0;
    L.l_cnt = s; s += a16(sizeof(int32_t) * (size_t)C);
    L.l_raw = s; s += a16(sizeof(int32_t) * (size_t)C);
    L.l_npos = s; s += a16(sizeof(int64_t) * (size_t)C);
    L.l_key = s; s += a16(sizeof(uint32_t) * (size_t)C * cap);
    L.l_bits = s; s += a16((size_t)C * cap);
    L.list_stride = s;
    L.lists = o; o += s * NR;
    L.hist_stride = a16(sizeof(uint32_t) * (size_t)C * L.hp);
    L.hist = o; o += L.hist_stride * W;
    L.half = (o + 255) & ~(size_t)255;
    return L;
}

struct ApxLocal {           // byte offsets inside the rank-local workspace
    size_t ctrl, done_hist, mcnt, npos_g, mkey, mbits, edge, hist, total;
};
// ctrl words: [0] epoch, [1] done_final, [2] overflow, [3] max per-rank count, [4] max merged count, [5] timeout
static ApxLocal apx_local(int C, int cap)
{
    ApxLocal L;
    const int hp = (cap + 1 + 3) & ~3;
    size_t o = 0;
    L.ctrl = o; o += 64;
    L.done_hist = o; o += a16(sizeof(uint32_t) * (size_t)C);
    L.mcnt = o; o += a16(sizeof(int32_t) * (size_t)C);
    L.npos_g = o; o += a16(sizeof(int64_t) * (size_t)C);
    L.mkey = o; o += a16(sizeof(uint32_t) * (size_t)C * cap);
    L.mbits = o; o += a16((size_t)C * cap);
    L.edge = o; o += a16((size_t)C * APX_ESTRIDE_BYTES);
    L.hist = o; o += a16(sizeof(uint32_t) * (size_t)C * hp);
    L.total = o;
    return L;
}

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spin until *flag has reached `tag` (wrap-safe); gives up after ~4 s so that a missing peer cannot hang the GPU
__device__ __forceinline__ bool wait_flag(const unsigned *flag, unsigned tag)
{
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (unsigned it = 0;; ++it) {
        if ((int)(ld_acquire_sys(flag) - tag) >= 0) return true;
        if ((it & 1023u) == 1023u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > 4000000000ull) return false;
        }
    }
}

struct ApxPeers { unsigned char *base[APX_MAXW]; };

// Programmatic dependent launch: the kernels of the chain are launched with the stream-serialisation attribute, so a
// successor's CTAs can take their SM slots while the predecessor is still running and start the moment it has finished
// and flushed (griddepcontrol.wait) -- the launch latency of the three dependent stages disappears from the chain.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

struct ApxParams {
    ApxPeers peers; ApxLayout sl; ApxLocal ll;
    unsigned char *local;
    const uint32_t *tp_key; const uint8_t *tp_bits; const int *tp_cnt; const long long *npos;
    int C, cap_list, cap, nthr, use07, rank, W, R, exchange;   // R = sorted runs per rank (list slots = W * R)
    double *result;
    double *result_m;   // the caller's pinned host result through its mapped device address (or null): no copy node needed
    unsigned long long *dbg;   // optional [C][8] globaltimer stamps of the merge stage (OVDET_APX_DBG_PTR; null in production)
};
#define XSTAMPC(cls, i) do { if (p.dbg && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.dbg[(size_t)(cls) * 32 + (i)] = t_; } } while (0)
#define XSTAMP(i) XSTAMPC(blockIdx.x, i)

__device__ __forceinline__ unsigned *apx_ctrl(const ApxParams &p) { return reinterpret_cast<unsigned *>(p.local + p.ll.ctrl); }

// Sort the (key, bits) entries [0, total) held in ksm[0..cap) / bsm[0..cap) by (key, bits, position); returns which half
// (0 / 1) of the two-buffer arrays ksm[2][cap] / bsm[2][cap] holds the result.  All 1024 threads of the CTA call it.
struct SortScratch { uint16_t *rnk, *whist; uint32_t *dbase, *wsum_s; int *skip_s; };
__device__ __forceinline__ int cta_sort(uint32_t *ksm, uint8_t *bsm, const SortScratch &sc, int cap, int total)
{
    uint16_t *rnk = sc.rnk, *whist = sc.whist;
    uint32_t *dbase = sc.dbase, *wsum_s = sc.wsum_s;
    int &skip_s = *sc.skip_s;
    const int tid = threadIdx.x;
    // Stable LSD radix sort of (key, bits) by 8-bit digits: the bits first (so that equal keys end up in a defined order
    // whatever order the lists were appended in), then the four key bytes.  Warp w owns the contiguous elements
    // [w*CH, (w+1)*CH): within a warp __match_any_sync ranks the elements of equal digit, one lane per digit bumps the
    // warp's count; a scan over (digit, warp) then gives every element its destination.  O(n) per pass -- the bitonic
    // network this replaces was 4x the instructions at 2 048 entries and 15x at 16 384.
    const int lane = tid & 31, warp = tid >> 5;
    const int CH = (((total + 31) >> 5) + 31) & ~31;
    const int nwarp = CH ? (total + CH - 1) / CH : 0;      // warps that own elements
    int cur = 0;
    if (total <= 384) {
        // short list (a rank's share of a small evaluation): one counting pass -- entry i goes to the number of entries
        // that sort before it by (key, bits, position); every thread reads the same entry at a time (broadcast)
        __syncthreads();
        if (tid < total) {
            const uint32_t ki = ksm[tid]; const uint8_t bi = bsm[tid];
            const unsigned long long vi = ((unsigned long long)ki << 8) | bi;
            int pos = 0;
#pragma unroll 8
            for (int j = 0; j < total; ++j) {   // unconditional broadcast loads and bitwise compares: nothing to branch on, the loads overlap
                const unsigned long long vj = ((unsigned long long)ksm[j] << 8) | bsm[j];
                pos += (int)(vj < vi) | ((int)(vj == vi) & (int)(j < tid));
            }
            ksm[cap + pos] = ki; bsm[cap + pos] = bi;
        }
        cur = 1;
        __syncthreads();
    } else
    for (int pass = 0; pass < 5; ++pass) {
        const int in0 = cur * cap, out0 = (cur ^ 1) * cap;
        for (int i = tid; i < nwarp * 128; i += 1024) reinterpret_cast<uint32_t *>(whist)[i] = 0u;
        if (tid == 0) skip_s = 0;
        __syncthreads();
        const int sh = 8 * (pass - 1);
        for (int r = 0; r < CH; r += 32) {
            const int i = warp * CH + r + lane;
            const bool valid = i < total;
            const unsigned d = valid ? (pass == 0 ? (unsigned)bsm[in0 + i] : ((ksm[in0 + i] >> sh) & 255u)) : (256u + lane);
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const int lead = __ffs(peers) - 1;
            unsigned old = 0;
            if (valid && lane == lead) { old = whist[warp * 256 + d]; whist[warp * 256 + d] = (uint16_t)(old + __popc(peers)); }
            __syncwarp();
            old = __shfl_sync(0xffffffffu, old, lead);
            if (valid) rnk[i] = (uint16_t)(old + __popc(peers & ((1u << lane) - 1u)));
        }
        __syncthreads();
        if (tid < 256) {   // digit tid: exclusive scan over the warps, then over the digits
            unsigned run = 0;
#pragma unroll 8
            for (int w2 = 0; w2 < nwarp; ++w2) { const unsigned t = whist[w2 * 256 + tid]; whist[w2 * 256 + tid] = (uint16_t)run; run += t; }
            if (run == (unsigned)total) skip_s = 1;   // every element has this digit: the pass would not move anything
            unsigned x = run;
            for (int off = 1; off < 32; off <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
            if (lane == 31) wsum_s[warp] = x;
            dbase[tid] = x - run;   // exclusive inside the warp of digits
        }
        __syncthreads();
        if (tid < 256) {
            unsigned add = 0;
            for (int w2 = 0; w2 < warp; ++w2) add += wsum_s[w2];
            dbase[tid] += add;
        }
        __syncthreads();
        if (!skip_s) {
            for (int r = 0; r < CH; r += 32) {
                const int i = warp * CH + r + lane;
                if (i < total) {
                    const uint32_t kk = ksm[in0 + i]; const uint8_t bb = bsm[in0 + i];
                    const unsigned d = pass == 0 ? (unsigned)bb : ((kk >> sh) & 255u);
                    const unsigned pos = dbase[d] + whist[warp * 256 + d] + rnk[i];
                    ksm[out0 + pos] = kk; bsm[out0 + pos] = bb;
                }
            }
            cur ^= 1;
        }
        __syncthreads();
    }
    return cur;
}

// ---- stage 1: SORT this rank's list of a class in R slices, push every sorted run to all peers.  grid (R, C), 1024 threads
// Sorting on the senders means W * R CTAs per class sort an (W * R)-th of the entries each, at the same time; the
// receivers only have to merge sorted runs (stage 2): a handful of binary searches per entry instead of a 5-pass sort of all.
__global__ void __launch_bounds__(1024, 1) apx_push_lists_kernel(ApxParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int cap = p.cap;
    uint32_t *ksm = reinterpret_cast<uint32_t *>(sm);                 // [2][cap]
    uint16_t *rnk = reinterpret_cast<uint16_t *>(ksm + 2 * cap);      // [cap]
    uint16_t *whist = rnk + cap;                                      // [32 warps][256 digits]
    uint8_t *bsm = reinterpret_cast<uint8_t *>(whist + 32 * 256);     // [2][cap]
    __shared__ uint32_t dbase[256], wsum_s[8];
    __shared__ int skip_s;
    const int sub = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    if (sub == 0) XSTAMPC(c, 8);
    pdl_wait();
    pdl_release();
    if (sub == 0) XSTAMPC(c, 9);
    const unsigned tag = apx_ctrl(p)[0] + 1u;
    const int raw = p.tp_cnt[c];
    const int nall = min(min(raw, p.cap_list), cap);
    const int i0 = (int)((long long)nall * sub / p.R), n = (int)((long long)nall * (sub + 1) / p.R) - i0;   // my slice of the list
    for (int i = tid; i < n; i += 1024) { ksm[i] = p.tp_key[(size_t)c * p.cap_list + i0 + i]; bsm[i] = p.tp_bits[(size_t)c * p.cap_list + i0 + i]; }
    const int cur = cta_sort(ksm, bsm, SortScratch{rnk, whist, dbase, wsum_s, &skip_s}, cap, n);
    if (sub == 0) XSTAMPC(c, 10);
    const uint4 *ks = reinterpret_cast<const uint4 *>(ksm + cur * cap);
    const uint4 *bs = reinterpret_cast<const uint4 *>(bsm + cur * cap);
    const long long np = sub == 0 ? p.npos[c] : 0;
    const int run = p.rank * p.R + sub;
    for (int dst = 0; dst < p.W; ++dst) {
        unsigned char *slot = p.peers.base[dst] + (size_t)(tag & 1u) * p.sl.half + p.sl.lists + (size_t)run * p.sl.list_stride;
        uint4 *kd = reinterpret_cast<uint4 *>(slot + p.sl.l_key + sizeof(uint32_t) * (size_t)c * cap);
        for (int i = tid; i < (n + 3) / 4; i += 1024) kd[i] = ks[i];
        uint4 *bd = reinterpret_cast<uint4 *>(slot + p.sl.l_bits + (size_t)c * cap);
        for (int i = tid; i < (n + 15) / 16; i += 1024) bd[i] = bs[i];
        if (tid == 0) {
            reinterpret_cast<int *>(slot + p.sl.l_cnt)[c] = n;
            reinterpret_cast<int *>(slot + p.sl.l_raw)[c] = sub == 0 ? raw : 0;   // the rank's true count (overflow check), once
            reinterpret_cast<long long *>(slot + p.sl.l_npos)[c] = np;
        }
    }
    // CTA barrier, then ONE system-scope fence per flag writer: fences are cumulative, so the stores of all 1024 threads
    // (ordered before the barrier) are visible to whoever acquires the flag.  A fence in every thread cost ~10 us here.
    __syncthreads();
    if (tid < p.W) {
        unsigned char *half = p.peers.base[tid] + (size_t)(tag & 1u) * p.sl.half;
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned *>(half + p.sl.flags_l) + (size_t)run * p.C + c, tag);
    }
    if (sub == 0) XSTAMPC(c, 11);
}

// ---- stage 2: per class, gather the W lists, sort, bin edges; zero the class's histogram.  grid C, 1024 threads
__global__ void __launch_bounds__(1024, 1) apx_merge_kernel(ApxParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    // two (key, bits) buffers for the radix passes (addressed by offset from one shared base: the compiler then knows
    // every access is a shared-memory access), the per-element rank inside (warp, digit), the per-warp digit counts
    const int cap = p.cap;
    uint32_t *ksm = reinterpret_cast<uint32_t *>(sm);                 // [2][cap]
    uint16_t *rnk = reinterpret_cast<uint16_t *>(ksm + 2 * cap);      // [cap]
    uint16_t *whist = rnk + cap;                                      // [32 warps][256 digits]
    uint8_t *bsm = reinterpret_cast<uint8_t *>(whist + 32 * 256);     // [2][cap]
    uint16_t *lo16 = reinterpret_cast<uint16_t *>(bsm + 2 * cap);      // [APX_BINS + 1] first list index of every bin
    __shared__ uint32_t dbase[256], wsum_s[8];
    __shared__ int skip_s;
    __shared__ int n_s[APX_MAXW], off_s[APX_MAXW + 1];
    __shared__ int bad_s;
    const int c = blockIdx.x, tid = threadIdx.x;
    XSTAMP(0);
    pdl_wait();
    pdl_release();     // 20 CTAs: the histogram CTAs may take the other SMs now and wait there
    XSTAMP(1);
    unsigned *ctrl = apx_ctrl(p);
    const unsigned tag = ctrl[0] + 1u;
    const unsigned char *half = p.exchange ? p.peers.base[p.rank] + (size_t)(tag & 1u) * p.sl.half : nullptr;
    if (tid == 0) bad_s = 0;
    __syncthreads();
    if (p.exchange && tid < p.W) {
        if (!wait_flag(reinterpret_cast<const unsigned *>(half + p.sl.flags_l) + (size_t)tid * p.C + c, tag)) bad_s = 1;
    }
    __syncthreads();
    if (bad_s) { if (tid == 0) atomicExch(&ctrl[5], 1u); }   // carry on with whatever is there: the result is flagged invalid
    XSTAMP(6);
    if (tid == 0) {
        int off = 0, mx = 0, raw_total = 0; long long np = 0;
        for (int r = 0; r < p.W; ++r) {
            int raw; long long q;
            if (p.exchange) {
                const unsigned char *slot = half + p.sl.lists + (size_t)r * p.sl.list_stride;
                raw = reinterpret_cast<const int *>(slot + p.sl.l_raw)[c];
                q = reinterpret_cast<const long long *>(slot + p.sl.l_npos)[c];
            } else { raw = p.tp_cnt[c]; q = p.npos[c]; }
            mx = max(mx, raw); raw_total += raw; np += q;
            int n = min(raw, min(p.cap_list, p.cap));   // what the producer could ship
            n = min(n, p.cap - off);                    // what still fits
            n_s[r] = n; off_s[r] = off; off += n;
        }
        off_s[p.W] = off;
        reinterpret_cast<int *>(p.local + p.ll.mcnt)[c] = off;
        reinterpret_cast<long long *>(p.local + p.ll.npos_g)[c] = np;
        if (raw_total > p.cap || mx > p.cap_list) atomicExch(&ctrl[2], 1u);
        atomicMax(&ctrl[3], (unsigned)mx);
        atomicMax(&ctrl[4], (unsigned)raw_total);
    }
    __syncthreads();
    const int total = off_s[p.W];
    XSTAMP(2);
    for (int r = 0; r < p.W; ++r) {
        const uint32_t *ksrc; const uint8_t *bsrc;
        if (p.exchange) {
            const unsigned char *slot = half + p.sl.lists + (size_t)r * p.sl.list_stride;
            ksrc = reinterpret_cast<const uint32_t *>(slot + p.sl.l_key) + (size_t)c * p.cap;
            bsrc = slot + p.sl.l_bits + (size_t)c * p.cap;
        } else { ksrc = p.tp_key + (size_t)c * p.cap_list; bsrc = p.tp_bits + (size_t)c * p.cap_list; }
        const int n = n_s[r], o = off_s[r];
        for (int i = tid; i < n; i += 1024) { ksm[o + i] = ksrc[i]; bsm[o + i] = bsrc[i]; }
    }
    int cur;
    if (!p.exchange) {
        cur = cta_sort(ksm, bsm, SortScratch{rnk, whist, dbase, wsum_s, &skip_s}, cap, total);
    } else {
        // the W runs arrive sorted (stage 1): an entry's place in the merged order is its index in its own run plus, for
        // every other run, the number of entries that sort before it -- ties broken by (bits, source rank), so the result
        // does not depend on arrival order.  Four entries of a thread are searched in lock step (independent probes).
        __syncthreads();
        const int W = p.W;
        for (int base = tid * 4; base < total; base += 4096) {
            unsigned long long v[4]; int pos[4], own[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = min(base + q, total - 1);
                v[q] = ((unsigned long long)ksm[i] << 8) | bsm[i];
                int r = 0;
                while (r + 1 < W && i >= off_s[r + 1]) ++r;
                own[q] = r; pos[q] = i - off_s[r];
            }
            for (int r2 = 0; r2 < W; ++r2) {
                const int o2 = off_s[r2], n2 = n_s[r2];
                if (n2 == 0) continue;
                int top = 1;
                while (top * 2 <= n2) top *= 2;
                int lo[4] = {0, 0, 0, 0};
                for (int step = top; step > 0; step >>= 1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int nx = lo[q] + step;
                        const int j = o2 + min(nx, n2) - 1;
                        const unsigned long long w = ((unsigned long long)ksm[j] << 8) | bsm[j];
                        // entries of an earlier run go first on ties (<=), those of a later run after (<)
                        const bool before = r2 < own[q] ? (w <= v[q]) : (w < v[q]);
                        lo[q] = ((nx <= n2) & before) ? nx : lo[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) pos[q] += (r2 != own[q]) ? lo[q] : 0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (base + q < total) { ksm[cap + pos[q]] = (uint32_t)(v[q] >> 8); bsm[cap + pos[q]] = (uint8_t)v[q]; }
        }
        cur = 1;
        __syncthreads();
    }
    XSTAMP(3);
    const uint32_t *ks = ksm + cur * cap; const uint8_t *bs = bsm + cur * cap;
    uint32_t *mk = reinterpret_cast<uint32_t *>(p.local + p.ll.mkey) + (size_t)c * p.cap;
    uint8_t *mb = p.local + p.ll.mbits + (size_t)c * p.cap;
    for (int i = tid; i < cap; i += 1024) { mk[i] = i < total ? ks[i] : 0xFFFFFFFFu; mb[i] = i < total ? bs[i] : 0; }
    // Bin table of the sorted list: the key axis between its extremes is cut into <= APX_BINS equal bins (the key is the
    // order-preserving image of the fp32 score, i.e. a piecewise-logarithmic scale that spreads a detector's skewed scores
    // evenly); entry b = index of the first list key >= the bin's lower bound, bit 15 set when the bin holds two or more
    // list keys.  A record then finds its bucket with one table look-up and one compare (apx_hist_kernel).  Built from
    // the entries: the first entry of a bin writes that bin and the empty bins in front of it.
    unsigned char *eb = p.local + p.ll.edge + (size_t)c * APX_ESTRIDE_BYTES;
    uint32_t kmin = 0; int shift = 0, nb = 0;
    if (total > 0) {
        kmin = ks[0];
        const uint32_t span = ks[total - 1] - kmin;
        while ((span >> shift) >= (uint32_t)APX_BINS) ++shift;
        nb = (int)(span >> shift) + 1;
    }
    if (tid == 0) {
        uint32_t *hdr = reinterpret_cast<uint32_t *>(eb);
        hdr[0] = kmin; hdr[1] = (uint32_t)shift; hdr[2] = (uint32_t)nb; hdr[3] = (uint32_t)total;
    }
    {
        // one thread per bin (a bin's first entry writing the run of empty bins in front of it would leave one thread
        // with thousands of stores when the scores have a gap): binary search of the bin's lower bound in the sorted keys
        uint16_t *e = reinterpret_cast<uint16_t *>(eb + sizeof(uint32_t) * APX_EHDR);
        // branch-free lower bounds, the (up to 9) bins of a thread in lock step so that their shared loads overlap
        constexpr int BPT = (APX_BINS + 1 + 1023) / 1024;
        int top = 1;
        while (top * 2 <= total) top *= 2;            // largest power of two <= total (1 when total <= 1)
        unsigned long long bound[BPT]; int lo[BPT];
#pragma unroll
        for (int q = 0; q < BPT; ++q) { bound[q] = (unsigned long long)kmin + ((unsigned long long)(tid + 1024 * q) << shift); lo[q] = 0; }
        for (int step = total > 0 ? top : 0; step > 0; step >>= 1) {
#pragma unroll
            for (int q = 0; q < BPT; ++q) {
                const int nx = lo[q] + step;
                const uint32_t kv = ks[min(nx, total) - 1];                                  // unconditional load (total >= 1 here)
                lo[q] = ((nx <= total) & ((unsigned long long)kv < bound[q])) ? nx : lo[q];   // entries with key < bound
            }
        }
#pragma unroll
        for (int q = 0; q < BPT; ++q) if (tid + 1024 * q <= nb) lo16[tid + 1024 * q] = (uint16_t)lo[q];
        __syncthreads();
        XSTAMP(4);
        for (int bi = tid; bi < nb; bi += 1024) e[bi] = (uint16_t)(lo16[bi] | (lo16[bi + 1] - lo16[bi] >= 2 ? 0x8000 : 0));
        if (tid == 0 && nb == 0) e[0] = 0;
    }
    uint32_t *h = reinterpret_cast<uint32_t *>(p.local + p.ll.hist) + (size_t)c * p.sl.hp;
    for (int i = tid; i < p.sl.hp; i += 1024) h[i] = 0;
    if (tid == 0) reinterpret_cast<unsigned *>(p.local + p.ll.done_hist)[c] = 0;
    XSTAMP(5);
}

// ---- stage 2 (several ranks): merge the W sorted runs of a class with a CLUSTER of 8 CTAs.  One SM's shared-memory
// bandwidth was the limit (151 us for 12 500 entries: every probe of every binary search is a shared load); each CTA
// of the cluster holds all runs as 8-byte (key << 8 | bits) words and places one eighth of the entries, writing them
// straight to their final position in global memory; after a cluster barrier each CTA reads the merged keys back and
// builds one eighth of the bin table.
constexpr int APX_CL = 8;
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(APX_CL, 1, 1) __launch_bounds__(1024, 1) apx_merge_runs_kernel(ApxParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int cap = p.cap;
    unsigned long long *v = reinterpret_cast<unsigned long long *>(sm);   // [cap] all runs, (key << 8) | bits
    uint32_t *ksm = reinterpret_cast<uint32_t *>(sm);                     // later: the merged keys, [cap]
    uint16_t *lo16 = reinterpret_cast<uint16_t *>(sm + sizeof(unsigned long long) * (size_t)cap);   // [APX_BINS / APX_CL + 8]
    __shared__ int n_s[APX_MAXW], off_s[APX_MAXW + 1];
    __shared__ int bad_s;
    const int c = blockIdx.x / APX_CL, part = blockIdx.x % APX_CL, tid = threadIdx.x;
    if (part == 0) XSTAMPC(c, 0);
    pdl_wait();
    pdl_release();
    unsigned *ctrl = apx_ctrl(p);
    const unsigned tag = ctrl[0] + 1u;
    const unsigned char *half = p.peers.base[p.rank] + (size_t)(tag & 1u) * p.sl.half;
    if (tid == 0) bad_s = 0;
    __syncthreads();
    const int NR = p.W * p.R;   // sorted runs to merge
    if (tid < NR) {
        if (!wait_flag(reinterpret_cast<const unsigned *>(half + p.sl.flags_l) + (size_t)tid * p.C + c, tag)) bad_s = 1;
    }
    __syncthreads();
    if (bad_s && tid == 0) atomicExch(&ctrl[5], 1u);   // carry on with whatever is there: the result is flagged invalid
    if (part == 0) XSTAMPC(c, 6);
    if (tid == 0) {
        int off = 0, mx = 0, raw_total = 0; long long np = 0;
        for (int r = 0; r < NR; ++r) {
            const unsigned char *slot = half + p.sl.lists + (size_t)r * p.sl.list_stride;
            const int raw = reinterpret_cast<const int *>(slot + p.sl.l_raw)[c];   // the source rank's true count (its first run carries it)
            np += reinterpret_cast<const long long *>(slot + p.sl.l_npos)[c];
            mx = max(mx, raw); raw_total += raw;
            int n = reinterpret_cast<const int *>(slot + p.sl.l_cnt)[c];           // entries in this run
            n = min(n, cap - off);                    // what still fits
            n_s[r] = n; off_s[r] = off; off += n;
        }
        off_s[NR] = off;
        if (part == 0) {
            reinterpret_cast<int *>(p.local + p.ll.mcnt)[c] = off;
            reinterpret_cast<long long *>(p.local + p.ll.npos_g)[c] = np;
            if (raw_total > cap || mx > p.cap_list) atomicExch(&ctrl[2], 1u);
            atomicMax(&ctrl[3], (unsigned)mx);
            atomicMax(&ctrl[4], (unsigned)raw_total);
        }
    }
    __syncthreads();
    const int total = off_s[NR], W = NR;   // (W = number of runs below)
    if (part == 0) XSTAMPC(c, 2);
    // one flat loop over the concatenated entries (not one loop per run: W dependent global round trips in a row)
#pragma unroll 4
    for (int i = tid; i < total; i += 1024) {
        int r = 0;
        while (r + 1 < W && i >= off_s[r + 1]) ++r;
        const unsigned char *slot = half + p.sl.lists + (size_t)r * p.sl.list_stride;
        const int j = i - off_s[r];
        const uint32_t kk = __ldcg(reinterpret_cast<const uint32_t *>(slot + p.sl.l_key) + (size_t)c * cap + j);
        const uint8_t bb = __ldcg(slot + p.sl.l_bits + (size_t)c * cap + j);
        v[i] = ((unsigned long long)kk << 8) | bb;
    }
    __syncthreads();
    uint32_t *mk = reinterpret_cast<uint32_t *>(p.local + p.ll.mkey) + (size_t)c * cap;
    uint8_t *mb = p.local + p.ll.mbits + (size_t)c * cap;
    // my eighth of the entries: place = index in the own run + entries of every other run that sort before it; ties by
    // (bits, source rank).  Four entries of a thread are searched in lock step (independent probes).
    const int s0 = (int)((long long)total * part / APX_CL), s1 = (int)((long long)total * (part + 1) / APX_CL);
    for (int base = s0 + tid * 4; base < s1; base += 4096) {
        unsigned long long x[4]; int pos[4], own[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = min(base + q, s1 - 1);
            x[q] = v[i];
            int r = 0;
            while (r + 1 < W && i >= off_s[r + 1]) ++r;
            own[q] = r; pos[q] = i - off_s[r];
        }
        for (int r2 = 0; r2 < W; ++r2) {
            const int o2 = off_s[r2], n2 = n_s[r2];
            if (n2 == 0) continue;
            int top = 1;
            while (top * 2 <= n2) top *= 2;
            int lo[4] = {0, 0, 0, 0};
            for (int step = top; step > 0; step >>= 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int nx = lo[q] + step;
                    const unsigned long long w = v[o2 + min(nx, n2) - 1];
                    const bool before = r2 < own[q] ? (w <= x[q]) : (w < x[q]);   // an earlier run goes first on ties
                    lo[q] = ((nx <= n2) & before) ? nx : lo[q];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) pos[q] += (r2 != own[q]) ? lo[q] : 0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (base + q < s1) { mk[pos[q]] = (uint32_t)(x[q] >> 8); mb[pos[q]] = (uint8_t)x[q]; }
    }
    // padding behind the list, the class's histogram row: an eighth each
    for (int i = total + part * 1024 + tid; i < cap; i += APX_CL * 1024) { mk[i] = 0xFFFFFFFFu; mb[i] = 0; }
    uint32_t *h = reinterpret_cast<uint32_t *>(p.local + p.ll.hist) + (size_t)c * p.sl.hp;
    for (int i = part * 1024 + tid; i < p.sl.hp; i += APX_CL * 1024) h[i] = 0;
    if (part == 0 && tid == 0) reinterpret_cast<unsigned *>(p.local + p.ll.done_hist)[c] = 0;
    __threadfence();
    cluster_sync_all();
    if (part == 0) XSTAMPC(c, 3);
    // ---- bin table (see apx_merge_kernel): the merged keys back into shared memory, an eighth of the bins per CTA
    for (int i = tid; i < total; i += 1024) ksm[i] = __ldcg(mk + i);
    __syncthreads();
    unsigned char *eb = p.local + p.ll.edge + (size_t)c * APX_ESTRIDE_BYTES;
    uint32_t kmin = 0; int shift = 0, nb = 0;
    if (total > 0) {
        kmin = ksm[0];
        const uint32_t span = ksm[total - 1] - kmin;
        while ((span >> shift) >= (uint32_t)APX_BINS) ++shift;
        nb = (int)(span >> shift) + 1;
    }
    if (part == 0 && tid == 0) {
        uint32_t *hdr = reinterpret_cast<uint32_t *>(eb);
        hdr[0] = kmin; hdr[1] = (uint32_t)shift; hdr[2] = (uint32_t)nb; hdr[3] = (uint32_t)total;
    }
    {
        uint16_t *e = reinterpret_cast<uint16_t *>(eb + sizeof(uint32_t) * APX_EHDR);
        const int b0 = part * (APX_BINS / APX_CL);                       // my bins [b0, b0 + 1024): thread t takes bin b0 + t
        int top = 1;
        while (top * 2 <= total) top *= 2;
        const int bi = b0 + tid;
        const unsigned long long bound = (unsigned long long)kmin + ((unsigned long long)bi << shift);
        int lo = 0;
        for (int step = total > 0 ? top : 0; step > 0; step >>= 1) {
            const int nx = lo + step;
            const uint32_t kv = ksm[min(nx, total) - 1];
            lo = ((nx <= total) & ((unsigned long long)kv < bound)) ? nx : lo;   // entries with key < bound
        }
        lo16[tid] = (uint16_t)lo;
        if (tid == 0) {   // the bound of the bin after my last one
            const unsigned long long bnd = (unsigned long long)kmin + ((unsigned long long)(b0 + 1024) << shift);
            int l2 = 0;
            for (int step = total > 0 ? top : 0; step > 0; step >>= 1) {
                const int nx = l2 + step;
                const uint32_t kv = ksm[min(nx, total) - 1];
                l2 = ((nx <= total) & ((unsigned long long)kv < bnd)) ? nx : l2;
            }
            lo16[1024] = (uint16_t)l2;
        }
        __syncthreads();
        if (bi < nb) e[bi] = (uint16_t)(lo16[tid] | (lo16[tid + 1] - lo16[tid] >= 2 ? 0x8000 : 0));
        if (part == 0 && tid == 0 && nb == 0) e[0] = 0;
    }
    if (part == 0) XSTAMPC(c, 5);
}

// ---- stage 3: histogram of the local records over the merged list; the last CTA of a class ships the class's row
__global__ void __launch_bounds__(1024) apx_hist_kernel(ApxParams p, const float *__restrict__ score, long long N, int last_block)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int cap = p.cap;
    const int NT = (int)blockDim.x;   // 256, or 1024 when the tables of a long merged list leave room for one CTA per SM only
    uint32_t *k = reinterpret_cast<uint32_t *>(sm);   // [cap + 1] sorted keys of the merged list + a 0xFFFFFFFF sentinel
    uint32_t *h = k + cap + 1;                        // [cap + 1] private histogram
    uint16_t *bins = reinterpret_cast<uint16_t *>(h + cap + 1);   // [APX_BINS] bin -> first list entry | crowded << 15
    __shared__ int last_s;
    const int c = blockIdx.y;
    pdl_wait();
    pdl_release();
    const unsigned char *eb = p.local + p.ll.edge + (size_t)c * APX_ESTRIDE_BYTES;
    const uint32_t *hdr = reinterpret_cast<const uint32_t *>(eb);
    const uint16_t *eg = reinterpret_cast<const uint16_t *>(eb + sizeof(uint32_t) * APX_EHDR);
    const uint32_t *tp_key = reinterpret_cast<const uint32_t *>(p.local + p.ll.mkey);
    uint32_t *hist = reinterpret_cast<uint32_t *>(p.local + p.ll.hist) + (size_t)c * p.sl.hp;
    const uint32_t kmin = hdr[0];
    const int shift = (int)hdr[1], nb = (int)hdr[2], ntp = (int)hdr[3];
    for (int i = threadIdx.x; i < ntp; i += NT) k[i] = tp_key[(size_t)c * cap + i];
    if (threadIdx.x == 0) k[ntp] = 0xFFFFFFFFu;
    for (int i = threadIdx.x; i <= ntp; i += NT) h[i] = 0;
    for (int i = threadIdx.x; i < (nb + 1) / 2 + (nb == 0 ? 1 : 0); i += NT) reinterpret_cast<uint32_t *>(bins)[i] = reinterpret_cast<const uint32_t *>(eg)[i];   // two u16 entries per load
    __syncthreads();
    // Every record scoring below all TPs lands in the one bucket after the last list entry (the bulk of the false
    // positives): counted in a register instead of hammering one shared word.
    const uint32_t kmax = ntp > 0 ? k[ntp - 1] : 0u;   // empty list: every valid key (>= 0x00800000) is "below"
    const uint32_t last_bin = (uint32_t)(nb > 0 ? nb - 1 : 0);
    unsigned int tail = 0;
    const float *sc = score + (size_t)c * N;
    // four records at a time in lock step: bin look-up, one compare against the bin's first list key (the sentinel /
    // the next bin's key when the bin is empty: no effect), and only for a crowded bin a short forward scan
    auto place4 = [&](const float (&sv)[4]) {
        uint32_t key[4], ent[4]; int lo[4]; bool cnt[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            key[q] = apc_score_key(sv[q]);
            const bool valid = sv[q] > -INFINITY, below = key[q] > kmax;
            tail += (valid && below) ? 1u : 0u;
            cnt[q] = valid && !below;
            const uint32_t d = key[q] > kmin ? key[q] - kmin : 0u;
            ent[q] = bins[min(d >> shift, last_bin)];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { lo[q] = (int)(ent[q] & 0x7fffu); lo[q] += (k[lo[q]] < key[q]) ? 1 : 0; }
        if (((ent[0] | ent[1] | ent[2] | ent[3]) & 0x8000u) != 0u) {   // some record sits in a crowded bin: finish its scan (the sentinel ends it)
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (ent[q] & 0x8000u) while (k[lo[q]] < key[q]) ++lo[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (cnt[q]) atomicAdd(&h[lo[q]], 1u);
    };
    auto place = [&](float s) { const float sv[4] = {s, -INFINITY, -INFINITY, -INFINITY}; place4(sv); };
    if (((reinterpret_cast<uintptr_t>(sc) & 15) == 0) && N < 0x7fffffffLL) {
        const int n4 = (int)(N >> 2);
        const int step = (int)(gridDim.x * NT);
        for (int i = (int)(blockIdx.x * NT + threadIdx.x); i < n4; i += step) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(sc) + i);
            const float sv[4] = {v.x, v.y, v.z, v.w};
            place4(sv);
        }
        for (long long i = ((long long)n4 << 2) + (long long)blockIdx.x * NT + threadIdx.x; i < N; i += (long long)gridDim.x * NT) place(sc[i]);
    } else {
        for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < N; i += (long long)gridDim.x * NT) place(sc[i]);
    }
    for (int off = 16; off > 0; off >>= 1) tail += __shfl_xor_sync(0xffffffffu, tail, off);
    if ((threadIdx.x & 31) == 0 && tail) atomicAdd(&h[ntp], tail);
    __syncthreads();
    for (int i = threadIdx.x; i <= ntp; i += NT) { const uint32_t v = h[i]; if (v) atomicAdd(&hist[i], v); }
    (void)last_block; (void)last_s;   // shipping to the peers is a separate stage (apx_ship_hist_kernel): one CTA per (class, peer)
}

// ---- stage 3, long merged list (> 4096 entries: several ranks' true positives): the private histogram packs two 16-bit
// counters per shared word (a CTA never sees more than 60 000 records), so two 512-thread CTAs fit an SM next to the
// 64 KB of list keys instead of one; a bin now holds 1-2 list keys on average, so the bucket is a short branch-free
// binary search inside the bin's range.  (Keys left in global memory instead: 3x slower, every probe is an L1/L2 trip.)
__global__ void __launch_bounds__(512) apx_hist_big_kernel(ApxParams p, const float *__restrict__ score, long long N, int last_block)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int cap = p.cap;
    const int NT = (int)blockDim.x;
    uint32_t *k = reinterpret_cast<uint32_t *>(sm);                         // [cap + 1] sorted keys of the merged list
    uint32_t *hw = k + cap + 1;                                             // [(cap + 2) / 2] two u16 counters per word
    uint16_t *bins = reinterpret_cast<uint16_t *>(hw + (cap + 2) / 2);      // [APX_BINS]
    __shared__ int last_s;
    const int c = blockIdx.y;
    pdl_wait();
    pdl_release();
    const unsigned char *eb = p.local + p.ll.edge + (size_t)c * APX_ESTRIDE_BYTES;
    const uint32_t *hdr = reinterpret_cast<const uint32_t *>(eb);
    const uint16_t *eg = reinterpret_cast<const uint16_t *>(eb + sizeof(uint32_t) * APX_EHDR);
    const uint32_t *kg = reinterpret_cast<const uint32_t *>(p.local + p.ll.mkey) + (size_t)c * cap;
    uint32_t *hist = reinterpret_cast<uint32_t *>(p.local + p.ll.hist) + (size_t)c * p.sl.hp;
    const uint32_t kmin = hdr[0];
    const int shift = (int)hdr[1], nb = (int)hdr[2], ntp = (int)hdr[3];
    for (int i = threadIdx.x; i < ntp; i += NT) k[i] = kg[i];
    for (int i = threadIdx.x; i < (ntp + 2) / 2; i += NT) hw[i] = 0;
    for (int i = threadIdx.x; i < (nb + 1) / 2 + (nb == 0 ? 1 : 0); i += NT) reinterpret_cast<uint32_t *>(bins)[i] = reinterpret_cast<const uint32_t *>(eg)[i];
    __syncthreads();
    const uint32_t kmax = ntp > 0 ? k[ntp - 1] : 0u;
    const uint32_t last_bin = (uint32_t)(nb > 0 ? nb - 1 : 0);
    unsigned int tail = 0;
    const float *sc = score + (size_t)c * N;
    auto place4 = [&](const float (&sv)[4]) {
        uint32_t key[4]; int lo[4], hi[4]; bool cnt[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            key[q] = apc_score_key(sv[q]);
            const bool valid = sv[q] > -INFINITY, below = key[q] > kmax;
            tail += (valid && below) ? 1u : 0u;
            cnt[q] = valid && !below;
            const uint32_t d = key[q] > kmin ? key[q] - kmin : 0u;
            const uint32_t bi = min(d >> shift, last_bin);
            lo[q] = (int)(bins[bi] & 0x7fffu);
            hi[q] = bi + 1 < (uint32_t)nb ? (int)(bins[bi + 1] & 0x7fffu) : ntp;
        }
        // lower bound inside [lo, hi): three branch-free steps cover ranges of up to 7 entries
#pragma unroll
        for (int step = 4; step > 0; step >>= 1)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int nx = lo[q] + step;
                const uint32_t kv = k[min(nx, hi[q]) - (hi[q] > 0 ? 1 : 0)];
                lo[q] = ((nx <= hi[q]) & (kv < key[q])) ? nx : lo[q];
            }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            while (cnt[q] && lo[q] < hi[q] && k[lo[q]] < key[q]) ++lo[q];   // a bin of more than 7 entries: finish the scan
            if (cnt[q]) atomicAdd(&hw[lo[q] >> 1], 1u << ((lo[q] & 1) * 16));
        }
    };
    auto place = [&](float s) { const float sv[4] = {s, -INFINITY, -INFINITY, -INFINITY}; place4(sv); };
    // The private counters are 16 bits wide: after at most 60 000 records of this CTA they are flushed to the class's
    // global row and cleared, so a CTA may take any number of records and the grid stays one balanced wave.
    auto flush = [&](bool clear) {
        __syncthreads();
        for (int i = threadIdx.x; i < (ntp + 2) / 2; i += NT) {
            const uint32_t w = hw[i];
            if (w) {
                if (w & 0xffffu) atomicAdd(&hist[2 * i], w & 0xffffu);
                if (w >> 16) atomicAdd(&hist[2 * i + 1], w >> 16);
                if (clear) hw[i] = 0;
            }
        }
        if (clear) __syncthreads();
    };
    if (((reinterpret_cast<uintptr_t>(sc) & 15) == 0) && N < 0x7fffffffLL) {
        const int n4 = (int)(N >> 2);
        const int step = (int)(gridDim.x * NT);
        const int iters = (n4 + step - 1) / step;               // the same trip count for every thread of the grid
        const int per_flush = max(1, 60000 / (4 * NT));         // iterations of this CTA between two flushes
        for (int it = 0; it < iters; ++it) {
            if (it && it % per_flush == 0) flush(true);
            const int i = it * step + (int)(blockIdx.x * NT + threadIdx.x);
            if (i < n4) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(sc) + i);
                const float sv[4] = {v.x, v.y, v.z, v.w};
                place4(sv);
            }
        }
        for (long long i = ((long long)n4 << 2) + (long long)blockIdx.x * NT + threadIdx.x; i < N; i += (long long)gridDim.x * NT) place(sc[i]);   // < 4 records
    } else {
        const long long step = (long long)gridDim.x * NT;
        const long long iters = (N + step - 1) / step;
        const int per_flush = max(1, 60000 / NT);
        for (long long it = 0; it < iters; ++it) {
            if (it && it % per_flush == 0) flush(true);
            const long long i = it * step + (long long)blockIdx.x * NT + threadIdx.x;
            if (i < N) place(sc[i]);
        }
    }
    for (int off = 16; off > 0; off >>= 1) tail += __shfl_xor_sync(0xffffffffu, tail, off);
    if ((threadIdx.x & 31) == 0 && tail) atomicAdd(&hist[ntp], tail);
    flush(false);
    (void)last_block; (void)last_s;   // shipping to the peers is a separate stage (apx_ship_hist_kernel): one CTA per (class, peer)
}

// ---- stage 3b: ship the finished histogram rows.  grid (W, C): CTA (dst, c) copies row c into peer dst's slot of this
// rank and raises the flag -- 160 CTAs move what one "last CTA" per class used to push alone (33 us at 16 384 buckets).
__global__ void __launch_bounds__(256) apx_ship_hist_kernel(ApxParams p)
{
    const int dst = blockIdx.x, c = blockIdx.y;
    pdl_wait();
    pdl_release();
    if (dst == 0) XSTAMPC(c, 12);
    const unsigned tag = apx_ctrl(p)[0] + 1u;
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(p.local + p.ll.hist) + (size_t)c * p.sl.hp);
    const int ntp = reinterpret_cast<const int *>(p.local + p.ll.mcnt)[c];
    const int span = min(max(1, (ntp + 1023) / 1024) * 1024, p.cap);      // what the final stage scans (plus the bucket behind it)
    const int n4 = min(p.sl.hp, span + 4) / 4;
    unsigned char *half = p.peers.base[dst] + (size_t)(tag & 1u) * p.sl.half;
    uint4 *out = reinterpret_cast<uint4 *>(half + p.sl.hist + (size_t)p.rank * p.sl.hist_stride + sizeof(uint32_t) * (size_t)c * p.sl.hp);
    for (int i = threadIdx.x; i < n4; i += 256) out[i] = __ldcg(src + i);
    __syncthreads();
    if (threadIdx.x == 0) {   // barrier + one cumulative system-scope fence (see apx_push_lists_kernel)
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned *>(half + p.sl.flags_h) + (size_t)p.rank * p.C + c, tag);
    }
    if (dst == 0) XSTAMPC(c, 13);
}

// ---- stage 4: per (class, threshold): sum the W histogram rows, positions, precision envelope, AP.  grid (C, nthr)
// Thread t owns E = span / 1024 CONSECUTIVE list entries: the prefix sums (positions, cumulative TP) are a serial walk
// over its own entries plus one block scan of the per-thread totals, the precision envelope (running maximum from the
// right, utils/eval_det.py:45-46) a reverse walk plus one block suffix-max -- eight barriers whatever the list length.
__global__ void __launch_bounds__(1024) apx_final_kernel(ApxParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned int *Hs = reinterpret_cast<unsigned int *>(sm);   // [span] bucket counts summed over the ranks
    uint8_t *Bs = reinterpret_cast<uint8_t *>(Hs + p.cap);     // [span] TP flag of the entry at this threshold
    __shared__ unsigned int wsH[32], wsT[32];
    __shared__ double wmax[32], red[32];
    __shared__ int bad_s, last_s;
    const int c = blockIdx.x, t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (t == 0) XSTAMPC(c, 16);
    pdl_wait();
    if (t == 0) XSTAMPC(c, 17);
    unsigned *ctrl = apx_ctrl(p);
    const unsigned tag = ctrl[0] + 1u;
    const unsigned char *half = p.exchange ? p.peers.base[p.rank] + (size_t)(tag & 1u) * p.sl.half : nullptr;
    if (tid == 0) bad_s = 0;
    __syncthreads();
    if (p.exchange && tid < p.W) {
        if (!wait_flag(reinterpret_cast<const unsigned *>(half + p.sl.flags_h) + (size_t)tid * p.C + c, tag)) bad_s = 1;
    }
    __syncthreads();
    if (bad_s && tid == 0) atomicExch(&ctrl[5], 1u);
    if (t == 0) XSTAMPC(c, 18);
    const double npos = (double)reinterpret_cast<const long long *>(p.local + p.ll.npos_g)[c];
    const int ntp = reinterpret_cast<const int *>(p.local + p.ll.mcnt)[c];
    const double eps = 2.220446049250313e-16;
    const uint8_t *bits = p.local + p.ll.mbits + (size_t)c * p.cap;
    const uint32_t *hl = reinterpret_cast<const uint32_t *>(p.local + p.ll.hist) + (size_t)c * p.sl.hp;
    auto hist_at = [&](int i) -> unsigned int {
        if (!p.exchange) return hl[i];
        unsigned int v = 0;
#pragma unroll 8
        for (int r = 0; r < p.W; ++r)   // unrolled: the W slot reads of a bucket are independent loads, issued together
            v += __ldcg(reinterpret_cast<const uint32_t *>(half + p.sl.hist + (size_t)r * p.sl.hist_stride) + (size_t)c * p.sl.hp + i);
        return v;
    };
    const int nch = max(1, (ntp + 1023) / 1024);
    const int span = nch * 1024 <= p.cap ? nch * 1024 : p.cap;   // entries that matter (the rest of the list is padding)
    const int E = span >> 10;
    for (int i = tid; i < span; i += 1024) { Hs[i] = hist_at(i); Bs[i] = (uint8_t)((bits[i] >> t) & 1u); }
    __syncthreads();
    // ---- forward: per-thread totals, block exclusive scan
    const int j0 = tid * E;
    unsigned int sumH = 0, sumT = 0;
    for (int j = 0; j < E; ++j) { sumH += Hs[j0 + j]; sumT += Bs[j0 + j]; }
    unsigned int xH = sumH, xT = sumT;
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned int yH = __shfl_up_sync(0xffffffffu, xH, off), yT = __shfl_up_sync(0xffffffffu, xT, off);
        if (lane >= off) { xH += yH; xT += yT; }
    }
    if (lane == 31) { wsH[warp] = xH; wsT[warp] = xT; }
    __syncthreads();
    if (tid < 32) {
        unsigned int wH = wsH[tid], wT = wsT[tid];
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int yH = __shfl_up_sync(0xffffffffu, wH, off), yT = __shfl_up_sync(0xffffffffu, wT, off);
            if (tid >= off) { wH += yH; wT += yT; }
        }
        wsH[tid] = wH; wsT[tid] = wT;
    }
    __syncthreads();
    const unsigned int exH = xH - sumH + (warp ? wsH[warp - 1] : 0u), exT = xT - sumT + (warp ? wsT[warp - 1] : 0u);
    const unsigned int total_tp = wsT[31];
    // number of present records = all buckets, including the one behind the last list entry
    unsigned long long nvalid = (unsigned long long)wsH[31];
    if (ntp >= span) nvalid += hist_at(ntp);   // ntp == span: the tail bucket sits just past the scanned range
    // ---- my entries: precision at the TP records, its maximum
    double pm = 0.0;
    {
        unsigned int H = exH, ct = exT;
        for (int j = 0; j < E; ++j) {
            H += Hs[j0 + j];
            if (Bs[j0 + j]) { ++ct; pm = fmax(pm, __ddiv_rn((double)ct, fmax((double)H, eps))); }
        }
    }
    // block suffix maximum (exclusive: threads after me)
    double m = pm;
    for (int off = 1; off < 32; off <<= 1) { const double y = __shfl_down_sync(0xffffffffu, m, off); if (lane + off < 32) m = fmax(m, y); }
    if (lane == 0) wmax[warp] = m;
    __syncthreads();
    if (tid < 32) {
        double w = wmax[tid];
        for (int off = 1; off < 32; off <<= 1) { const double y = __shfl_down_sync(0xffffffffu, w, off); if (tid + off < 32) w = fmax(w, y); }
        wmax[tid] = w;
    }
    __syncthreads();
    double env = warp < 31 ? wmax[warp + 1] : 0.0;                   // warps after mine
    {
        const double nx = __shfl_down_sync(0xffffffffu, m, 1);       // inclusive suffix max of the next lane = lanes after me
        if (lane < 31) env = fmax(env, nx);
    }
    // ---- reverse walk: AP = sum over TP records of (recall step) x (max precision at or after the record)
    double ap_local = 0.0;
    double p11[11];
#pragma unroll
    for (int kk = 0; kk < 11; ++kk) p11[kk] = 0.0;
    {
        unsigned int H = exH + sumH, ct = exT + sumT;
        for (int j = E - 1; j >= 0; --j) {
            if (Bs[j0 + j]) {
                const double ctd = (double)ct;
                const double prec = __ddiv_rn(ctd, fmax((double)H, eps));
                const double rec = npos > 0.0 ? __ddiv_rn(ctd, npos) : 0.0;
                const double rec_prev = npos > 0.0 ? __ddiv_rn(ctd - 1.0, npos) : 0.0;
                env = fmax(env, prec);
                ap_local += __dmul_rn(__dsub_rn(rec, rec_prev), env);
                if (p.use07) {
#pragma unroll
                    for (int kk = 0; kk < 11; ++kk) if (rec >= kk * 0.1) p11[kk] = fmax(p11[kk], prec);
                }
                --ct;
            }
            H -= Hs[j0 + j];
        }
    }
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
        // VOC07: max precision over records with rec >= t; attained at a TP record unless t == 0, where the very
        // first record counts too -- its precision is 1 if it is a TP (covered) else 0 (covered by the 0 init).
        res = 0.0;
        for (int kk = 0; kk < 11; ++kk) {
            double a = p11[kk];
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
    if (t == 0) XSTAMPC(c, 19);
    if (tid == 0) {
        const size_t kx = (size_t)p.nthr * p.C;
        const double v_ap = nvalid > 0 ? res : 0.0, v_rec = (nvalid > 0 && npos > 0.0) ? __ddiv_rn((double)total_tp, npos) : 0.0;
        p.result[(size_t)t * p.C + c] = v_ap;
        p.result[kx + (size_t)t * p.C + c] = v_rec;
        if (t == 0) p.result[2 * kx + c] = (double)nvalid;
        if (p.result_m) {
            p.result_m[(size_t)t * p.C + c] = v_ap;
            p.result_m[kx + (size_t)t * p.C + c] = v_rec;
            if (t == 0) p.result_m[2 * kx + c] = (double)nvalid;
        }
        __threadfence();
        const unsigned tk = atomicAdd(&ctrl[1], 1u);
        last_s = (tk == gridDim.x * gridDim.y - 1);
        if (last_s) {   // everything of this evaluation is done on this rank: publish the status words, advance the epoch
            __threadfence();
            // written by the merge stage (an earlier kernel): plain L2 reads, one 16-byte store to clear them
            const uint2 s01 = __ldcg(reinterpret_cast<const uint2 *>(ctrl + 2)), s23 = __ldcg(reinterpret_cast<const uint2 *>(ctrl + 4));
            const unsigned ovf = s01.x, mx = s01.y, mt = s23.x, to = s23.y;
            ctrl[2] = 0u; ctrl[3] = 0u; ctrl[4] = 0u; ctrl[5] = 0u;
            p.result[2 * kx + p.C] = to ? -1.0 : (double)ovf;
            p.result[2 * kx + p.C + 1] = (double)mx;
            p.result[2 * kx + p.C + 2] = (double)mt;
            if (p.result_m) {
                p.result_m[2 * kx + p.C] = to ? -1.0 : (double)ovf;
                p.result_m[2 * kx + p.C + 1] = (double)mx;
                p.result_m[2 * kx + p.C + 2] = (double)mt;
            }
            ctrl[1] = 0u;
            __threadfence();
            ctrl[0] = tag;
        }
    }
}

}  // namespace ovdet

extern "C" size_t ovdet_apx_local_bytes(int C, int cap_total)
{
    if (C <= 0 || !apc_cap_ok(cap_total)) return 0;
    return apx_local(C, cap_total).total;
}

extern "C" size_t ovdet_apx_symm_bytes(int C, int cap_total, int world)
{
    if (C <= 0 || !apc_cap_ok(cap_total) || world < 1 || world > APX_MAXW) return 0;
    return 2 * apx_layout(C, cap_total, world).half;
}

extern "C" int ovdet_apx_reduce(const void *const *blocks, const int64_t *block_n, int nblocks, int C,
                                const uint32_t *tp_key, const uint8_t *tp_bits, const int32_t *tp_cnt, const int64_t *npos,
                                int cap_list, int cap_total, int nthr, unsigned flags, int rank, int world,
                                void *const *peers, void *local_ws, double *result, double *result_host, void *stream)
{
    OVDET_REQUIRE(C > 0 && nblocks >= 0 && apc_cap_ok(cap_total) && nthr >= 1 && nthr <= 8, "bad size (cap_total must be a power of two in [1024, 16384])");
    OVDET_REQUIRE(cap_list >= 16 && cap_list % 16 == 0, "cap_list must be a positive multiple of 16");
    OVDET_REQUIRE(world >= 1 && world <= APX_MAXW && rank >= 0 && rank < world, "bad rank / world");
    OVDET_REQUIRE(tp_key && tp_bits && tp_cnt && npos && local_ws && result, "null pointer");
    OVDET_REQUIRE(nblocks == 0 || (blocks && block_n), "null block table");
    const bool exchange = world > 1 || (flags & OVDET_APX_FORCE_EXCHANGE);
    OVDET_REQUIRE(!exchange || peers, "the exchange needs the table of symmetric buffers");
    const bool via_slots = peers != nullptr;   // lists travel as sorted runs through the (symmetric-layout) buffer, also on one rank
    ApxParams p;
    memset(&p, 0, sizeof(p));
    p.sl = apx_layout(C, cap_total, world);
    p.ll = apx_local(C, cap_total);
    if (via_slots) for (int r = 0; r < world; ++r) { OVDET_REQUIRE(peers[r], "null symmetric buffer"); p.peers.base[r] = static_cast<unsigned char *>(peers[r]); }
    p.local = static_cast<unsigned char *>(local_ws);
    p.tp_key = tp_key; p.tp_bits = tp_bits; p.tp_cnt = tp_cnt; p.npos = reinterpret_cast<const long long *>(npos);
    p.C = C; p.cap_list = cap_list; p.cap = cap_total; p.nthr = nthr; p.use07 = (flags & OVDET_APX_USE_07_METRIC) ? 1 : 0;
    p.rank = rank; p.W = world; p.R = apx_runs_per_rank(world); p.exchange = exchange ? 1 : 0; p.result = result;
    p.result_m = nullptr;
    if (result_host) {   // pinned host memory is written by the final kernel itself; anything else gets a copy node
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, result_host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) p.result_m = static_cast<double *>(at.devicePointer);
        else cudaGetLastError();
    }
    { const char *e = getenv("OVDET_APX_DBG_PTR"); p.dbg = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0)) : nullptr; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned stages = flags & (OVDET_APX_STAGE_PUSH | OVDET_APX_STAGE_MERGE_HIST | OVDET_APX_STAGE_FINAL);
    if (!stages) stages = OVDET_APX_STAGE_PUSH | OVDET_APX_STAGE_MERGE_HIST | OVDET_APX_STAGE_FINAL;
    if (via_slots && (stages & OVDET_APX_STAGE_PUSH)) {
        const size_t psmem = (size_t)cap_total * 12 + sizeof(uint16_t) * 32 * 256;
        OVDET_CUDA_TRY(ensure_dyn_smem(apx_push_lists_kernel, psmem));
        OVDET_CUDA_TRY(launch_pdl(apx_push_lists_kernel, dim3(p.R, C), dim3(1024), psmem, st, p));
        { const int rc = launch_ok("apx_push_lists_kernel"); if (rc) return rc; }
    }
    if ((stages & OVDET_APX_STAGE_MERGE_HIST) && !via_slots) {
        const size_t smem = (size_t)cap_total * 12 + sizeof(uint16_t) * (32 * 256 + APX_BINS + 8);
        OVDET_CUDA_TRY(ensure_dyn_smem(apx_merge_kernel, smem));
        OVDET_CUDA_TRY(launch_pdl(apx_merge_kernel, dim3(C), dim3(1024), smem, st, p));
        { const int rc = launch_ok("apx_merge_kernel"); if (rc) return rc; }
    }
    if ((stages & OVDET_APX_STAGE_MERGE_HIST) && via_slots) {   // the runs arrive sorted: a cluster of 8 CTAs per class merges them
        const size_t smem = (size_t)cap_total * 8 + sizeof(uint16_t) * (APX_BINS / APX_CL + 8);
        OVDET_CUDA_TRY(ensure_dyn_smem(apx_merge_runs_kernel, smem));
        OVDET_CUDA_TRY(launch_pdl(apx_merge_runs_kernel, dim3(APX_CL * C), dim3(1024), smem, st, p));   // cluster dims are compiled in
        { const int rc = launch_ok("apx_merge_runs_kernel"); if (rc) return rc; }
    }
    if (stages & OVDET_APX_STAGE_MERGE_HIST) {
        const bool big = cap_total > 4096;
        const size_t smem = big ? sizeof(uint32_t) * ((size_t)cap_total + 1 + ((size_t)cap_total + 2) / 2) + sizeof(uint16_t) * APX_BINS
                                : sizeof(uint32_t) * 2 * ((size_t)cap_total + 1) + sizeof(uint16_t) * APX_BINS;
        if (big) OVDET_CUDA_TRY(ensure_dyn_smem(apx_hist_big_kernel, smem));
        else OVDET_CUDA_TRY(ensure_dyn_smem(apx_hist_kernel, smem));
        // CTAs that fit one SM: 228 KB of shared memory, 1 KB reserved per CTA, 2048 threads.  The grid is ONE balanced wave:
        // a few CTAs more than the slots would leave most SMs idle behind a second wave (measured on 8 GPUs: 220 CTAs of
        // the long-list kernel on 148 SMs, two resident on half of them, 114 us against 51 us for the short-list kernel)
        int per_sm = (int)(233472 / (smem + 1024 + 64));
        const int by_threads = 2048 / (big ? 512 : APC_NT);
        if (per_sm > by_threads) per_sm = by_threads;
        if (per_sm > 8) per_sm = 8;
        if (per_sm < 1) per_sm = 1;
        int gx_full = 148 * per_sm / C;
        if (gx_full < 1) gx_full = 1;
        auto launch_hist = [&](const float *blk, long long N, int last) -> int {
            // a CTA pays the set-up of its tables (list keys, 16 KB of bins) and the flush of its histogram: with few records
            // per class use 1024-thread CTAs (the set-up is spread over 4x the threads) of >= 16 records per thread
            int nt = big ? 512 : APC_NT;
            long long gx = (N + 8191) / 8192;
            if (!big && gx < gx_full / 2) { nt = 1024; gx = (N + 16383) / 16384; }
            if (gx > gx_full) gx = gx_full;
            if (gx < 1) gx = 1;
            if (big) OVDET_CUDA_TRY(launch_pdl(apx_hist_big_kernel, dim3((unsigned)gx, C), dim3(nt), smem, st, p, blk, N, last));
            else OVDET_CUDA_TRY(launch_pdl(apx_hist_kernel, dim3((unsigned)gx, C), dim3(nt), smem, st, p, blk, N, last));
            return OVDET_OK;
        };
        for (int b = 0; b < nblocks; ++b) {
            const long long N = block_n[b];
            OVDET_REQUIRE(N >= 0, "negative block size");
            const bool last = (b == nblocks - 1);
            if (N == 0 && !(exchange && last)) continue;
            OVDET_REQUIRE(N == 0 || blocks[b], "null record block");
            { const int rc = launch_hist(static_cast<const float *>(blocks[b]), N, last ? 1 : 0); if (rc) return rc; }
        }
        if (exchange) {   // every rank ships its rows (all zero when it had no records) so that the peers' final stage can run
            OVDET_CUDA_TRY(launch_pdl(apx_ship_hist_kernel, dim3(world, C), dim3(256), 0, st, p));
            { const int rc = launch_ok("apx_ship_hist_kernel"); if (rc) return rc; }
        }
    }
    if (stages & OVDET_APX_STAGE_FINAL) {
        const size_t smem = (sizeof(unsigned int) + 1) * (size_t)cap_total;
        OVDET_CUDA_TRY(ensure_dyn_smem(apx_final_kernel, smem));
        OVDET_CUDA_TRY(launch_pdl(apx_final_kernel, dim3(C, nthr), dim3(1024), smem, st, p));
        { const int rc = launch_ok("apx_final_kernel"); if (rc) return rc; }
    }
    if (result_host && !p.result_m && (stages & OVDET_APX_STAGE_FINAL))
        OVDET_CUDA_TRY(cudaMemcpyAsync(result_host, result, sizeof(double) * (2 * (size_t)nthr * C + C + 3), cudaMemcpyDeviceToHost, st));
    return OVDET_OK;
}
