// lsap.cu -- on-device linear sum assignment for Matcher.forward (criterion.py:76-86).
//
// The reference moves the [B,Q,G] cost matrix to the host and calls
// scipy.optimize.linear_sum_assignment per sample on the first nactual_gt[b]
// columns.  scipy (1.18.1; not vendored in the reference) implements the
// shortest-augmenting-path algorithm of Crouse, "On implementing 2D rectangular
// assignment algorithms" (2016), in fp64, transposing so that rows <= columns.
// This kernel restates that published algorithm with one 256-thread CTA per sample:
// the cost slab is staged transposed in shared memory, the threads stride over the
// columns for the Dijkstra relaxation, and a shuffle argmin per warp + 8 warp winners
// through a double-buffered shared array (one barrier per step) replace the serial scan.  On tie-free costs the optimum is
// unique, so assignments equal scipy's (north_star: bit-exact except ties).
// Tie-break used here: lowest reduced cost, then unassigned column, then lowest index.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace ovdet {

struct LsapParams {
    const float *cost;
    const int64_t *nactual;
    int64_t *inds;
    float *mask;
    int32_t *col_to_row;
    int B, Q, G, M;   // M = max(Q, G): capacity of the per-column arrays
    int stage_cost;   // 1: the [min(Q,n), max(Q,n)] cost slab is staged (transposed if needed) in shared memory
    unsigned long long *dbg;   // optional [B][4]: start, after staging, end (globaltimer), Dijkstra steps (OVDET_LSAP_DBG_PTR)
};
#define LSTAMP(i) do { if (p.dbg && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.dbg[(size_t)blockIdx.x * 4 + (i)] = t_; } } while (0)

constexpr int LS_NT = 256;

struct Cand { double v; int j; int asg; };
__device__ __forceinline__ bool cand_less(const Cand &a, const Cand &b)
{   // lowest reduced cost, then unassigned column, then lowest index
    return a.v < b.v || (a.v == b.v && (a.asg < b.asg || (a.asg == b.asg && a.j < b.j)));
}

// The same order on integers, so that the arg-min is three redux.sync per warp instead of five rounds of fp64 shuffles and
// compares (the arg-min is the critical path of a Dijkstra step): kv = order-preserving image of the reduced cost
// (zero canonicalised to +0 first, so that equal doubles have equal images), kj = (assigned << 31) | column.
struct Key { unsigned long long kv; unsigned kj; };
__device__ __forceinline__ unsigned long long dkey(double v)
{
    const long long b = __double_as_longlong(__dadd_rn(v, 0.0));
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
__device__ __forceinline__ double dkey_inv(unsigned long long k)
{
    const long long b = (k & 0x8000000000000000ull) ? (long long)(k ^ 0x8000000000000000ull) : (long long)~k;
    return __longlong_as_double(b);
}
__device__ __forceinline__ bool key_less(const Key &a, const Key &b) { return a.kv < b.kv || (a.kv == b.kv && a.kj < b.kj); }
__device__ __forceinline__ Key key_of(const Cand &c) { return Key{dkey(c.v), ((unsigned)c.asg << 31) | (unsigned)c.j}; }
__device__ __forceinline__ Key warp_min_key(Key k)
{
    const unsigned full = 0xffffffffu;
    const unsigned hi = (unsigned)(k.kv >> 32), lo = (unsigned)k.kv;
    const unsigned mh = __reduce_min_sync(full, hi);
    const unsigned ml = __reduce_min_sync(full, hi == mh ? lo : 0xffffffffu);
    const unsigned mj = __reduce_min_sync(full, (hi == mh && lo == ml) ? k.kj : 0xffffffffu);
    return Key{((unsigned long long)mh << 32) | ml, mj};
}

// One CTA (8 warps) per sample.  Thread t owns columns t, t+256, ...: it relaxes them, the per-warp argmin
// goes through shuffles, the 8 warp winners through a double-buffered shared array (one barrier per Dijkstra step).
__global__ void __launch_bounds__(LS_NT) lsap_kernel(LsapParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    double *u = reinterpret_cast<double *>(sm);      // [M] row duals
    double *v = u + p.M;                              // [M] column duals
    double *spc = v + p.M;                            // [M] shortest path costs
    int *path = reinterpret_cast<int *>(spc + p.M);   // [M]
    int *row4col = path + p.M;                        // [M]
    int *col4row = row4col + p.M;                     // [M]
    unsigned char *SR = reinterpret_cast<unsigned char *>(col4row + p.M);  // [M]
    unsigned char *SC = SR + p.M;                                         // [M]
    float *sc = reinterpret_cast<float *>(sm + (((size_t)p.M * (3 * sizeof(double) + 3 * sizeof(int) + 2)) + 15 & ~(size_t)15));
    __shared__ Key wbest[2][LS_NT / 32];

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    LSTAMP(0);
    int nsteps = 0;
    long long nn = p.nactual ? p.nactual[b] : p.G;
    const int n = nn < 0 ? 0 : (nn > p.G ? p.G : (int)nn);
    const float *cb = p.cost + (size_t)b * p.Q * p.G;

    for (int q = tid; q < p.Q; q += LS_NT) { p.inds[(size_t)b * p.Q + q] = 0; p.mask[(size_t)b * p.Q + q] = 0.f; }
    if (p.col_to_row) for (int g = tid; g < p.G; g += LS_NT) p.col_to_row[(size_t)b * p.G + g] = -1;
    if (n == 0) return;

    const bool transposed = n < p.Q;           // scipy: transpose when nc < nr
    const int nr = transposed ? n : p.Q;
    const int nc = transposed ? p.Q : n;
    const int ldc = nc + 1;
    // scipy.optimize.linear_sum_assignment raises ValueError on NaN / -inf entries ("matrix contains invalid numeric
    // entries") and on infeasible problems: such a sample is not solved here either -- its mask is set to -1 and its
    // col_to_row entries to -2, which the Python shim turns into the same exception (criterion.lsap)
    bool bad = false;
    if (p.stage_cost) {   // coalesced global read (16 bytes per thread when the rows allow it), conflict-free (odd stride) shared write
        if ((p.G & 3) == 0 && (reinterpret_cast<uintptr_t>(cb) & 15) == 0) {
            const int n4 = (n + 3) >> 2;
            for (int idx = tid; idx < p.Q * n4; idx += LS_NT) {
                const int q = idx / n4, g = (idx - q * n4) << 2;
                const float4 v = __ldg(reinterpret_cast<const float4 *>(cb + (size_t)q * p.G + g));   // g + 3 < G: columns >= n are just not stored
                const float c4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (g + t < n) { bad |= !(c4[t] > -INFINITY) ; if (transposed) sc[(g + t) * ldc + q] = c4[t]; else sc[q * ldc + g + t] = c4[t]; }
            }
        } else {
            for (int idx = tid; idx < p.Q * n; idx += LS_NT) {
                const int q = idx / n, g = idx - q * n;
                const float c = __ldg(cb + (size_t)q * p.G + g);
                bad |= !(c > -INFINITY);
                if (transposed) sc[g * ldc + q] = c; else sc[q * ldc + g] = c;
            }
        }
    }
    else {
        for (int idx = tid; idx < p.Q * n; idx += LS_NT) { const int q = idx / n, g = idx - q * n; bad |= !(__ldg(cb + (size_t)q * p.G + g) > -INFINITY); }
    }
    auto fail = [&]() {
        for (int q = tid; q < p.Q; q += LS_NT) p.mask[(size_t)b * p.Q + q] = -1.f;
        if (p.col_to_row) for (int g = tid; g < p.G; g += LS_NT) p.col_to_row[(size_t)b * p.G + g] = -2;
    };
    if (__syncthreads_or(bad ? 1 : 0)) { fail(); return; }   // NaN or -inf somewhere in cost[b, :, :n]
    auto C = [&](int i, int j) -> double {
        if (p.stage_cost) return (double)sc[i * ldc + j];
        return transposed ? (double)__ldg(cb + (size_t)j * p.G + i) : (double)__ldg(cb + (size_t)i * p.G + j);
    };
    for (int i = tid; i < nr; i += LS_NT) { u[i] = 0.0; col4row[i] = -1; }
    for (int j = tid; j < nc; j += LS_NT) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
    __syncthreads();

    // per-row state is reset once here and then together with the path walk of a row that needed the general search
    for (int i = tid; i < nr; i += LS_NT) SR[i] = 0;
    for (int j = tid; j < nc; j += LS_NT) { SC[j] = 0; spc[j] = INFINITY; }
    __syncthreads();
    LSTAMP(1);
    int buf = 0;
    for (int cur = 0; cur < nr; ++cur) {
        double minVal = 0.0;
        int i = cur, sink = -1;
        bool first = true;   // first Dijkstra step of this row: nothing stored yet
        while (sink == -1) {
            ++nsteps;
            const double ui = u[i];
            Cand best{INFINITY, 0x7fffffff, 1};
            if (first) {
                // every column is unscanned and its shortest-path cost is the reduced cost itself: keep it in registers.
                // Measured on config 2: 1.01 steps per row -- the cheapest column of a new row is almost always free.
                for (int j = tid; j < nc; j += LS_NT) {
                    const double r = __dsub_rn(__dsub_rn(__dadd_rn(minVal, C(i, j)), ui), v[j]);
                    const Cand c{r, j, row4col[j] != -1};
                    if (cand_less(c, best)) best = c;
                }
            } else {
                if (tid == 0) SR[i] = 1;
                for (int j = tid; j < nc; j += LS_NT) {   // own columns only: spc/path/SC[j] are private to this thread
                    if (SC[j]) continue;
                    const double r = __dsub_rn(__dsub_rn(__dadd_rn(minVal, C(i, j)), ui), v[j]);
                    double s = spc[j];
                    if (r < s) { s = r; spc[j] = r; path[j] = i; }
                    const Cand c{s, j, row4col[j] != -1};
                    if (cand_less(c, best)) best = c;
                }
            }
            {
                Key kb = warp_min_key(key_of(best));
                if (lane == 0) wbest[buf][warp] = kb;
                __syncthreads();
                // the 8 warp winners: lanes 0..7 of every warp take one each and the same three redux.sync finish the job
                kb = lane < LS_NT / 32 ? wbest[buf][lane] : Key{~0ull, 0xffffffffu};
                kb = warp_min_key(kb);
                buf ^= 1;
                best.v = dkey_inv(kb.kv); best.j = (int)(kb.kj & 0x7fffffffu); best.asg = (int)(kb.kj >> 31);
            }
            if (best.j == 0x7fffffff || best.v == INFINITY) { sink = -2; break; }   // infeasible (non-finite costs)
            const int j = best.j;
            if (first) {
                if (row4col[j] == -1) {
                    // Free column at the first step: the path is (cur -> j), SR = {cur}, SC = {j}, and the dual updates reduce
                    // to u[cur] += minVal (v[j] -= minVal - spc[j] = +0).  The column's owner applies them; nobody else reads
                    // these words before the next barrier, so the row costs ONE barrier and no shared-memory state.
                    if ((j % LS_NT) == tid) { row4col[j] = cur; col4row[cur] = j; u[cur] = __dadd_rn(u[cur], best.v); }
                    sink = -3;
                    break;
                }
                // materialise what the general search expects after its first step
                for (int jj = tid; jj < nc; jj += LS_NT) {
                    spc[jj] = __dsub_rn(__dsub_rn(__dadd_rn(minVal, C(i, jj)), ui), v[jj]);
                    path[jj] = i;
                }
                if (tid == 0) SR[i] = 1;
                first = false;
            }
            minVal = best.v;
            if ((j % LS_NT) == tid) SC[j] = 1;     // the owner marks its column scanned
            if (row4col[j] == -1) sink = j; else i = row4col[j];
        }
        if (sink == -3) continue;
        // no barrier here: everything the dual updates read across threads (SR, spc) was written before the barrier of
        // the last step; SC / spc / v of a column are touched by its owner only
        if (sink < 0) break;
        // dual updates (Crouse 2016, step 4)
        if (tid == 0) u[cur] = __dadd_rn(u[cur], minVal);
        for (int r = tid; r < nr; r += LS_NT)
            if (SR[r] && r != cur) u[r] = __dadd_rn(u[r], __dsub_rn(minVal, spc[col4row[r]]));
        for (int j = tid; j < nc; j += LS_NT)
            if (SC[j]) v[j] = __dsub_rn(v[j], __dsub_rn(minVal, spc[j]));
        __syncthreads();   // duals done (they read spc / col4row / SR / SC)
        // thread 0 augments along the alternating path (path[], row4col, col4row) while everybody resets the per-row state
        if (tid == 0) {
            int j = sink;
            while (true) {
                const int r = path[j];
                row4col[j] = r;
                const int t = col4row[r]; col4row[r] = j; j = t;
                if (r == cur) break;
            }
        }
        for (int i2 = tid; i2 < nr; i2 += LS_NT) SR[i2] = 0;
        for (int j = tid; j < nc; j += LS_NT) { SC[j] = 0; spc[j] = INFINITY; }
        __syncthreads();
    }
    __syncthreads();
    LSTAMP(2);
    {   // a row without a column: the search ran out of finite costs (+inf rows): infeasible
        bool unassigned = false;
        for (int r = tid; r < nr; r += LS_NT) unassigned |= col4row[r] < 0;
        if (__syncthreads_or(unassigned ? 1 : 0)) { fail(); return; }
    }
    if (p.dbg && tid == 0) p.dbg[(size_t)b * 4 + 3] = ((unsigned long long)n << 32) | (unsigned)nsteps;
    for (int r = tid; r < nr; r += LS_NT) {
        const int c = col4row[r];
        if (c < 0) continue;
        const int q = transposed ? c : r, g = transposed ? r : c;
        p.inds[(size_t)b * p.Q + q] = g;
        p.mask[(size_t)b * p.Q + q] = 1.f;
        if (p.col_to_row) p.col_to_row[(size_t)b * p.G + g] = q;
    }
}

}  // namespace ovdet

using namespace ovdet;

extern "C" int ovdet_lsap_f32(const float *cost, const int64_t *nactual_gt, int B, int Q, int G,
                              int64_t *per_prop_gt_inds, float *proposal_matched_mask, int32_t *col_to_row, void *stream)
{
    OVDET_REQUIRE(B >= 0 && Q >= 0 && G >= 0, "negative size");
    if (B == 0 || Q == 0) return OVDET_OK;
    OVDET_REQUIRE(cost || G == 0, "null cost");
    OVDET_REQUIRE(per_prop_gt_inds && proposal_matched_mask, "null output");
    LsapParams p;
    p.cost = cost; p.nactual = nactual_gt; p.inds = per_prop_gt_inds; p.mask = proposal_matched_mask;
    p.col_to_row = col_to_row; p.B = B; p.Q = Q; p.G = G; p.M = Q > G ? Q : G;
    OVDET_REQUIRE(p.M <= 4096, "Q and G must be <= 4096");
    { const char *e = getenv("OVDET_LSAP_DBG_PTR"); p.dbg = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0)) : nullptr; }
    size_t smem = (((size_t)p.M * (3 * sizeof(double) + 3 * sizeof(int) + 2)) + 15) & ~(size_t)15;
    const size_t slab = sizeof(float) * (size_t)(Q < G ? Q : G) * ((size_t)p.M + 1);
    p.stage_cost = (smem + slab <= 200 * 1024) ? 1 : 0;
    if (p.stage_cost) smem += slab;
    OVDET_CUDA_TRY(ensure_dyn_smem(lsap_kernel, smem));
    lsap_kernel<<<B, LS_NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return launch_ok("lsap_kernel");
}

// One call for what criterion.py:348-361 + Matcher.forward do per decoder layer (or for all layers batched along B): the
// fused GIoU + L1-centre + class + objectness cost kernel, then the per-sample assignment -- two launches, one trip
// through the host shim.
extern "C" int ovdet_matcher_cost_f32(const float *sem_cls_prob, const float *objectness, const float *center_dist,
                                      const float *center_q, const float *center_g, const float *gious,
                                      const float *corners1, const float *corners2, const int64_t *gt_labels,
                                      const int64_t *nactual_gt, int B, int Q, int G, int C,
                                      float w_class, float w_obj, float w_center, float w_giou,
                                      unsigned giou_flags, int k2_cap, float *gious_out, float *cost, void *stream);

extern "C" int ovdet_matcher_step_f32(const float *sem_cls_prob, const float *objectness, const float *center_q, const float *center_g,
                                      const float *corners1, const float *corners2, const int64_t *gt_labels,
                                      const int64_t *nactual_gt, int B, int Q, int G, int C,
                                      float w_class, float w_obj, float w_center, float w_giou,
                                      unsigned giou_flags, int k2_cap, float *gious_out, float *cost,
                                      int64_t *per_prop_gt_inds, float *proposal_matched_mask, int32_t *col_to_row, void *stream)
{
    OVDET_REQUIRE(cost && per_prop_gt_inds && proposal_matched_mask, "null output");
    int rc = ovdet_matcher_cost_f32(sem_cls_prob, objectness, nullptr, center_q, center_g, nullptr, corners1, corners2, gt_labels, nactual_gt,
                                    B, Q, G, C, w_class, w_obj, w_center, w_giou, giou_flags, k2_cap, gious_out, cost, stream);
    if (rc) return rc;
    return ovdet_lsap_f32(cost, nactual_gt, B, Q, G, per_prop_gt_inds, proposal_matched_mask, col_to_row, stream);
}
