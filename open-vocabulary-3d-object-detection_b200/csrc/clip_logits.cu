// clip_logits.cu -- open-vocabulary logits for sm_100a: tcgen05 + TMEM + TMA.
//
// Replaces  cls_logits = sem_cls_head(visual_embeds)            (models/model_3detr.py:237-238,
//           the frozen CLIP text matrix as an nn.Linear weight, :151-154)
//           prob = softmax(cls_logits); sem_cls_prob = prob[..., :-1];
//           objectness = 1 - prob[..., -1]                       (:58-62)
// and, with OVDET_LOGITS_L2NORM, the normalise + temperature form of
// utils/ulip_losses.py:39-47 (logits = scale * x/|x| . t/|t|).
//
// One CTA computes a 128 x BLOCK_N tile of logits = X[128,K] * T[BLOCK_N,K]^T:
//   warp 0   TMA producer: 128B-swizzled [128 x 64] bf16 A boxes and [BLOCK_N x 64]
//            B boxes through a STAGES-deep mbarrier ring
//   warp 1   one elected thread issues tcgen05.mma (cta_group::1, kind::f16,
//            M=128, N=BLOCK_N, K=16) with the fp32 accumulator in TMEM;
//            tcgen05.commit releases ring slots and finally signals the epilogue
//   warp 2   TMEM allocate / free
//   warps 4-7 epilogue: thread r owns accumulator row r (tcgen05.ld 32x32b)
// A softmax row spans all N columns, i.e. several N-tiles: the NC CTAs of a
// thread-block CLUSTER hold the N-tiles of the same 128 rows, each computes
// (row max, sum exp) over its tile while the logits stay in TMEM, the pairs are
// exchanged through distributed shared memory (st.shared::cluster), and after one
// cluster barrier every CTA normalises its own tile out of TMEM.  No logits round
// trip through HBM, no second GEMM.  Probabilities leave through a padded shared
// tile with 16-byte coalesced row stores.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ovdet {

constexpr int GM_M = 128;        // rows per CTA (UMMA_M)
constexpr int GM_K = 64;         // bf16 per 128B swizzle row
constexpr int GM_THREADS = 256;
constexpr int GM_MAX_NC = 8;     // portable cluster size

struct LogitsParams {
    int M, N, K, block_n, nc, num_kb, stages, tmem_cols;
    unsigned flags;
    float scale;
    float *logits; int ld_logits;
    __nv_bfloat16 *prob; int ld_prob;
    float *objectness;
    const float *inv_nx, *inv_nt;
    unsigned long long *dbg;   // optional [gridDim.x][8] globaltimer stamps of warp 4 lane 0 (profiling; NULL in production)
};

__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define STAMP(i) do { if (p.dbg && threadIdx.x == 128) p.dbg[(size_t)blockIdx.x * 8 + (i)] = gtimer(); } while (0)

__global__ void __launch_bounds__(GM_THREADS, 2)
clip_logits_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const LogitsParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) float2 stats[GM_MAX_NC][GM_M];   // (row max, row sum-exp) from every CTA of the cluster
    __shared__ __align__(8) float2 part[2][GM_M];            // per-row partials of the two column groups
    __shared__ __align__(16) float colscale[256 + 32];       // 1/|t_col| of this tile's columns (L2NORM only)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int m_tile = blockIdx.x / p.nc;
    const int m0 = m_tile * GM_M;
    const int n0 = (int)rank * p.block_n;
    const uint32_t a_bytes = GM_M * GM_K * 2, b_bytes = (uint32_t)p.block_n * GM_K * 2;
    const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);

    STAMP(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        // A is multicast: a ring slot may be refilled only when EVERY CTA of the cluster has consumed it,
        // so each MMA commit arrives on the empty barrier of all nc CTAs
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], (uint32_t)p.nc); }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
    if (p.inv_nt) for (int c = threadIdx.x; c < 256 + 32; c += GM_THREADS) colscale[c] = __ldg(p.inv_nt + min(n0 + c, p.N - 1));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();   // every CTA of the cluster is resident before any DSMEM traffic
    STAMP(1);
    const uint32_t tmem_base = tmem_base_s;
    const uint16_t mc_mask = (uint16_t)((1u << p.nc) - 1u);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                unsigned char *sa = smem + (size_t)s * stage_bytes;
                mbar_expect_tx(&full_bar[s], a_bytes + b_bytes);   // 16 KB of A arrive as nc multicast slices, B is ours
                if (p.nc > 1) {
                    const int rows = GM_M / p.nc;                  // this CTA fetches rows [rank*rows, +rows) of the A tile for everyone
                    tma_load_2d_mc(&tmA, &full_bar[s], sa + (size_t)rank * rows * (GM_K * 2), kb * GM_K, m0 + (int)rank * rows, mc_mask);
                } else {
                    tma_load_2d(&tmA, &full_bar[s], sa, kb * GM_K, m0);
                }
                tma_load_2d(&tmB, &full_bar[s], sa + a_bytes, kb * GM_K, n0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b BF16 (1<<7, 1<<10), K-major both, N>>3 at 17, M>>4 at 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(GM_M >> 4) << 24);
        for (int kb = 0; kb < p.num_kb; ++kb) {
            const int s = kb % p.stages;
            const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sa + a_bytes);
#pragma unroll
                for (int k = 0; k < GM_K / 16; ++k)   // +32 B (>>4 = 2) per UMMA_K=16 step inside the swizzle atom
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                if (p.nc > 1) umma_commit_mc(&empty_bar[s], mc_mask);   // slot free (cluster-wide) once these MMAs retire
                else umma_commit(&empty_bar[s]);
                if (kb == p.num_kb - 1) umma_commit(&tmem_full_bar);  // accumulator complete
            }
            __syncwarp();
        }
    }

    // ===================== epilogue: all 8 warps =====================
    // TMEM lane quarter q = warp % 4 holds accumulator rows 32q..32q+31 (hardware rule: a warp may only
    // touch lanes 32*(warpid%4)..+31).  Warps 4-7 take the even 32-column chunks of the tile, warps 0-3
    // (free once their producer / MMA / alloc roles end) the odd ones; per-row partial (max, sum-exp) pairs
    // are merged through shared memory, published to the whole cluster through DSMEM, and after one cluster
    // barrier each warp normalises its own chunks straight out of TMEM.  Everything is in the log2 domain:
    // y = logit * log2(e), p = 2^(y - max) / sum.
    const int grp = (warp >> 2) ^ 1;                   // warps 4-7 -> group 0 (even chunks), warps 0-3 -> group 1
    const int row = (warp & 3) * 32 + lane;            // accumulator row == TMEM lane
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    const bool row_ok = (m0 + row) < p.M;
    float rs = p.scale;
    if (p.inv_nx && row_ok) rs *= __ldg(p.inv_nx + m0 + row);
    const float a2 = rs * LOG2E;
    float rmax = -INFINITY, rsum = 0.f;

    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    STAMP(2);
    // ---- pass 1: online (max, sum-exp) over this warp's chunks.  Each chunk's exponentials 2^(y - m_c) (m_c = the
    // running max after that chunk) are written BACK into the accumulator's TMEM columns, so pass 2 only rescales:
    // one MUFU.EX2 per element in total (the epilogue is MUFU-bound: 16 exp/clk/SM).
    float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // m_c of this warp's chunks (BLOCK_N <= 256 -> <= 4)
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
        const int c0 = grp * 32 + ci * 64;
        const int w = min(32, p.block_n - c0);
        const int nv = min(w, p.N - (n0 + c0));      // valid columns in this chunk (warp-uniform)
        if (c0 < p.block_n && nv > 0) {
            float v[32];
            tmem_ld32(trow + (uint32_t)c0, v);   // the TMEM allocation is a power of two >= BLOCK_N: a 16-column tail chunk may read (and mask) 32
            if (p.inv_nt) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 cs = *reinterpret_cast<const float4 *>(colscale + c0 + i);
                    v[i] *= cs.x; v[i + 1] *= cs.y; v[i + 2] *= cs.z; v[i + 3] *= cs.w;
                }
            }
            if (p.logits && row_ok) {
                float *o = p.logits + (size_t)(m0 + row) * p.ld_logits + n0 + c0;
#pragma unroll
                for (int i = 0; i < 32; ++i) if (i < nv) o[i] = v[i] * rs;
            }
            float cmax = -INFINITY;
            if (nv == 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) { v[i] *= a2; cmax = fmaxf(cmax, v[i]); }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) { v[i] = i < nv ? v[i] * a2 : -INFINITY; cmax = fmaxf(cmax, v[i]); }
            }
            const float nm = fmaxf(rmax, cmax);
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                v[i] = fast_exp2(v[i] - nm); v[i + 1] = fast_exp2(v[i + 1] - nm);
                acc0 += v[i]; acc1 += v[i + 1];
            }
            tmem_st32(trow + (uint32_t)c0, v);
            rsum = rsum * fast_exp2(rmax - nm) + (acc0 + acc1);
            rmax = nm;
            cm[ci] = nm;
        }
    }
    tmem_wait_st();
    part[grp][row] = make_float2(rmax, rsum);
    STAMP(3);
    __syncthreads();
    {   // merge the two column groups, publish to every CTA of the cluster (each group serves half of the ranks)
        const float2 s0 = part[0][row], s1 = part[1][row];
        const float m = fmaxf(s0.x, s1.x);
        float sm = 0.f;
        if (m > -INFINITY) sm = s0.y * fast_exp2(s0.x - m) + s1.y * fast_exp2(s1.x - m);
        for (int r = grp; r < p.nc; r += 2) st_cluster_f32x2(&stats[rank][row], (uint32_t)r, m, sm);
    }
    cluster_sync_all();   // release/acquire: all stats visible cluster-wide
    STAMP(4);

    if (p.prob || p.objectness) {
        float gmax = -INFINITY;
        for (int r = 0; r < p.nc; ++r) gmax = fmaxf(gmax, stats[r][row].x);
        float gsum = 0.f;
        for (int r = 0; r < p.nc; ++r) {
            const float2 s = stats[r][row];
            if (s.x > -INFINITY) gsum += s.y * fast_exp2(s.x - gmax);
        }
        const float inv = 1.f / gsum;
        // ---- pass 2: normalise out of TMEM into the padded shared tile (pipeline smem is idle now)
        const int tile_ld = p.block_n + 8;   // bf16 elements; +16 B keeps the 16-byte row writes conflict-free
        __nv_bfloat16 *tile = reinterpret_cast<__nv_bfloat16 *>(smem);
        const int obj_c = p.N - 1 - n0;      // tile column holding the background class, if in this tile
        // ---- pass 2: p = e_c * 2^(m_c - gmax) / gsum -- no transcendental per element
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
            const int c0 = grp * 32 + ci * 64;
            if (c0 >= p.block_n) continue;
            const int w = min(32, p.block_n - c0);
            const int nv = min(w, p.N - (n0 + c0));
            float v[32];
            if (nv > 0) {
                const float f = fast_exp2(cm[ci] - gmax) * inv;
                tmem_ld32(trow + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (nv == 32 || i < nv) ? v[i] * f : 0.f;
                if (p.objectness && obj_c >= c0 && obj_c < c0 + 32) {          // warp-uniform: one extra 1-column TMEM read
                    const float pb = tmem_ld1(trow + (uint32_t)obj_c) * f;
                    if (row_ok) p.objectness[m0 + row] = 1.f - pb;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            if (p.prob) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    if (i < w) {
                        __nv_bfloat162 h[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
                        *reinterpret_cast<uint4 *>(tile + (size_t)row * tile_ld + c0 + i) = *reinterpret_cast<uint4 *>(h);
                    }
                }
            }
        }
        STAMP(5);
        if (p.prob) {
            __syncthreads();
            // one row per warp and step, one 16-byte piece per lane (no index division)
            for (int c8 = lane * 8; c8 < p.block_n; c8 += 256) {
                const int gcol = n0 + c8;
                if (gcol >= p.ld_prob) continue;
#pragma unroll 4
                for (int r = warp; r < GM_M; r += GM_THREADS / 32) {
                    const int grow = m0 + r;
                    if (grow >= p.M) break;
                    const uint4 val = *reinterpret_cast<const uint4 *>(tile + (size_t)r * tile_ld + c8);
                    *reinterpret_cast<uint4 *>(p.prob + (size_t)grow * p.ld_prob + gcol) = val;
                }
            }
        }
    }
    STAMP(6);
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
    STAMP(7);
}

// inverse L2 norms of bf16 rows (F.normalize(x, dim=-1, p=2): x / max(|x|, 1e-12)), one warp per row
__global__ void __launch_bounds__(256) inv_norm_kernel(const __nv_bfloat16 *__restrict__ x, int R, int K, float *__restrict__ out)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= R) return;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) { const float v = __bfloat162float(x[(size_t)row * K + k]); s += v * v; }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) out[row] = 1.f / fmaxf(sqrtf(s), 1e-12f);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int make_bf16_map(CUtensorMap *map, const void *base, int rows, int K, int box_rows)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return OVDET_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)GM_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return OVDET_ERR_CUDA; }
    return OVDET_OK;
}

int clip_logits_persistent_launch(const void *x, const void *text, int M, int K, int N, int nc, int bn, unsigned flags, float scale,
                                  float *logits, int ld_logits, void *prob, int ld_prob, float *objectness,
                                  const float *inv_nx, const float *inv_nt, cudaStream_t st);   // clip_logits_persistent.cu
int clip_logits_persistent_choose(int M, int K, int N, int *nc_out, int *bn_out);

}  // namespace ovdet

using namespace ovdet;

extern "C" int ovdet_clip_logits_bf16(const void *x, const void *text, int M, int K, int N, unsigned flags, float scale,
                                      float *logits, int ld_logits, void *prob, int ld_prob, float *objectness, void *stream)
{
    OVDET_REQUIRE(M >= 0 && K > 0 && N > 0, "bad size");
    if (M == 0) return OVDET_OK;
    OVDET_REQUIRE(x && text, "null pointer");
    OVDET_REQUIRE(K % GM_K == 0, "K must be a multiple of 64");
    OVDET_REQUIRE(N <= GM_MAX_NC * 256, "N must be <= 2048");
    OVDET_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(text) & 15) == 0, "x/text must be 16-byte aligned");
    OVDET_REQUIRE(!logits || ld_logits >= N, "ld_logits < N");
    OVDET_REQUIRE(!prob || (ld_prob >= N && ld_prob % 8 == 0 && (reinterpret_cast<uintptr_t>(prob) & 15) == 0),
                  "prob needs ld_prob >= N, ld_prob % 8 == 0 and 16-byte alignment");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // split N over a cluster of nc CTAs (power of two), BLOCK_N = multiple of 16 <= 256
    int nc = 1;
    while (nc < GM_MAX_NC && (N + nc - 1) / nc > 256) nc <<= 1;
    int bn = (((N + nc - 1) / nc) + 15) / 16 * 16;
    if (bn < 16) bn = 16;
    LogitsParams p;
    p.M = M; p.N = N; p.K = K; p.block_n = bn; p.nc = nc; p.num_kb = K / GM_K;
    p.stages = 2;   // two CTAs are co-resident per SM (4 stages in flight), one in its mainloop while the other normalises
    { const char *e = getenv("OVDET_LOGITS_STAGES"); if (e) p.stages = atoi(e); }   // profiling experiments
    if (p.stages > p.num_kb) p.stages = p.num_kb;
    if (p.stages > 8) p.stages = 8;
    p.tmem_cols = 32; while (p.tmem_cols < bn) p.tmem_cols <<= 1;
    p.flags = flags; p.scale = scale;
    p.logits = logits; p.ld_logits = ld_logits; p.prob = static_cast<__nv_bfloat16 *>(prob); p.ld_prob = ld_prob;
    p.objectness = objectness; p.inv_nx = nullptr; p.inv_nt = nullptr;
    { const char *e = getenv("OVDET_LOGITS_DBG_PTR"); p.dbg = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0)) : nullptr; }
    float *norms = nullptr;
    if (flags & OVDET_LOGITS_L2NORM) {
        { void *ws = nullptr; int rc0 = device_scratch().acquire(sizeof(float) * ((size_t)M + N), st, &ws); if (rc0) return rc0; norms = static_cast<float *>(ws); }
        inv_norm_kernel<<<(M + 7) / 8, 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(x), M, K, norms);
        inv_norm_kernel<<<(N + 7) / 8, 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(text), N, K, norms + M);
        p.inv_nx = norms; p.inv_nt = norms + M;
    }
    // Two kernels.  Persistent (csrc/clip_logits_persistent.cu: one CTA per SM, double-buffered TMEM, 16 epilogue warps
    // working on tile i under the mainloop of tile i+1, cluster size chosen for co-residency) when there are at least two
    // rounds of M-tiles: 25.8 us at config 4 (cluster of 6 x 208 columns, 22 clusters resident) against 30.6 us for the
    // one-tile-per-CTA kernel below, which remains the choice for small M.  OVDET_LOGITS_PERSISTENT=0/1 forces either.
    static int persistent = -1;
    if (persistent < 0) { const char *e = getenv("OVDET_LOGITS_PERSISTENT"); persistent = e ? atoi(e) : 2; }
    int pnc = nc, pbn = bn;
    if (persistent && clip_logits_persistent_choose(M, K, N, &pnc, &pbn)) {
        int prc = clip_logits_persistent_launch(x, text, M, K, N, pnc, pbn, flags, scale, logits, ld_logits, prob, ld_prob, objectness,
                                                p.inv_nx, p.inv_nt, st);
        if (norms) { int rc0 = device_scratch().release(st); if (rc0) return rc0; }
        return prc;
    }
    CUtensorMap tmA, tmB;
    int rc = make_bf16_map(&tmA, x, M, K, GM_M / nc);   // each CTA of the cluster fetches (and multicasts) 128/nc rows of the A tile
    if (rc) return rc;
    rc = make_bf16_map(&tmB, text, N, K, bn);
    if (rc) return rc;
    const size_t stage_bytes = (size_t)GM_M * GM_K * 2 + (((size_t)bn * GM_K * 2 + 1023) & ~(size_t)1023);
    size_t smem = stage_bytes * p.stages;
    const size_t tile_bytes = (size_t)GM_M * (bn + 8) * 2;
    if (smem < tile_bytes) smem = tile_bytes;
    smem += 1024;  // manual 1024-byte alignment for the 128B swizzle
    OVDET_CUDA_TRY(cudaFuncSetAttribute(clip_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    const int m_tiles = (M + GM_M - 1) / GM_M;
    cfg.gridDim = dim3((unsigned)(m_tiles * nc));
    cfg.blockDim = dim3(GM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    OVDET_CUDA_TRY(cudaLaunchKernelEx(&cfg, clip_logits_kernel, tmA, tmB, p));
    if (norms) { int rc0 = device_scratch().release(st); if (rc0) return rc0; }
    return launch_ok("clip_logits_kernel");
}
