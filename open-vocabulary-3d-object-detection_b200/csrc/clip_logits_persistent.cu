// clip_logits_persistent.cu -- persistent, warp-specialised form of the open-vocabulary logits kernel.
//
// Same maths and outputs as clip_logits_kernel (csrc/clip_logits.cu: logits = X T^T on tcgen05, row softmax finished
// across the N-tiles of a thread-block cluster through DSMEM), restructured after profiling the one-tile-per-CTA
// version (profiles/r1_notes.md): there a CTA lived 13.5 us -- 6.2 us of TMA-latency-bound mainloop with only two
// ring stages, then 6.2 us of epilogue during which its TMA and tensor pipe idled -- and 512 CTAs made two full waves.
// Here one CTA per SM stays resident and walks the M-tiles of its cluster:
//   warp 0      TMA producer, ring of STAGES slots running ahead across tiles (A slice multicast to the cluster)
//   warp 1      MMA issuer into one of TWO TMEM accumulators (tcgen05.mma, commit -> ring slot / accumulator ready)
//   warp 2      TMEM alloc / free
//   warps 4-19  epilogue of the PREVIOUS tile while the next one is being multiplied (16 warps: four per TMEM lane
//               quarter, each taking every fourth 16-column chunk): one TMEM read per element,
//               exponentials parked as bf16 in a shared tile,
//               accumulator released early, per-row (max, sum) merged across the cluster with DSMEM stores +
//               st.async stores that complete bytes on the peer's mbarrier (no CTA-wide barrier.cluster, no fences in the steady state), then a coalesced,
//               rescaling copy-out.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ovdet {

constexpr int PK_M = 128, PK_K = 64;
#ifndef PK_EPI_WARPS
#define PK_EPI_WARPS 16                          // epilogue warps (multiple of 4).  Measured at config 4: 16 -> 24.0 us, 28 (1024 threads, 64 regs) -> 25.8 us: the epilogue is throughput-, not latency-bound
#endif
constexpr int PK_EPI = 32 * PK_EPI_WARPS, PK_THREADS = 128 + PK_EPI;   // 4 role warps + the epilogue warps
constexpr int PK_NG = PK_EPI / 128;              // column groups (epilogue warps per TMEM lane quarter)
constexpr int PK_CPG = (16 + PK_NG - 1) / PK_NG; // 16-column chunks per group at BLOCK_N = 256
constexpr int PK_MAX_NC = 8, PK_MAX_STAGES = 6;
constexpr int PK_ACC_STRIDE = 256;   // TMEM columns between the two accumulators (512 allocated)

struct PLogitsParams {
    int M, N, K, block_n, nc, nslice, num_kb, stages, m_tiles, num_clusters;
    float scale;
    float *logits; int ld_logits;
    __nv_bfloat16 *prob; int ld_prob;
    float *objectness;
    const float *inv_nx, *inv_nt;
    unsigned long long *dbg;   // optional [gridDim.x][8 tiles][8] globaltimer stamps of epilogue thread 0 (profiling)
};

__device__ __forceinline__ unsigned long long pk_gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PSTAMP(i) do { if (p.dbg && et == 0 && it < 8) p.dbg[((size_t)blockIdx.x * 8 + it) * 8 + (i)] = pk_gtimer(); } while (0)

__device__ __forceinline__ void mbar_arrive_local(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *local_bar, uint32_t rank)
{   // release at cluster scope: this thread's earlier DSMEM stores are visible to whoever acquires the phase
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// DSMEM store that completes `8` tx bytes on the destination CTA's mbarrier: the producer/consumer hand-off needs no fence
__device__ __forceinline__ void st_async_f32x2(const void *local_ptr, uint64_t *local_bar, uint32_t rank, float a, float b)
{
    uint32_t raddr, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_ptr)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "f"(a), "f"(b), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(PK_EPI) : "memory"); }

__global__ void __launch_bounds__(PK_THREADS, 1)
clip_logits_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PLogitsParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[PK_MAX_STAGES], empty_bar[PK_MAX_STAGES], tmem_full_bar[2], tmem_empty_bar[2], stats_full_bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) float2 stats[2][PK_MAX_NC][PK_M];   // [buffer][source CTA][row] = (max, sum-exp), log2 domain
    __shared__ __align__(8) float2 part[PK_NG][PK_M];           // the column groups of this CTA
    __shared__ float ftab[PK_M][17];                            // per row, per 16-column chunk: 2^(m_c - gmax) / gsum (padded)
    __shared__ __align__(16) float colscale[256 + 32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.dbg && threadIdx.x == 0) p.dbg[((size_t)blockIdx.x * 8 + 7) * 8 + 0] = pk_gtimer();   // kernel entry
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x / p.nc;
    const int n0 = (int)rank * p.block_n;
    const uint32_t a_bytes = PK_M * PK_K * 2, b_bytes = (uint32_t)p.block_n * PK_K * 2;
    const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int tile_ld = p.block_n + 8;
    __nv_bfloat16 *tile = reinterpret_cast<__nv_bfloat16 *>(smem + (size_t)p.stages * stage_bytes);
    const uint16_t mc_mask = (uint16_t)((1u << p.nc) - 1u);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], (uint32_t)p.nc); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full_bar[b], 1);
            mbar_init(&tmem_empty_bar[b], PK_EPI / 32);          // one arrive per epilogue warp
            mbar_init(&stats_full_bar[b], 1);   // one local expect_tx arrive; the nc*128 remote st.async complete the bytes
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&tmem_base_s, 512);
    if (p.inv_nt) for (int c = threadIdx.x; c < 256 + 32; c += PK_THREADS) colscale[c] = __ldg(p.inv_nt + min(n0 + c, p.N - 1));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();   // all CTAs of the cluster resident, barriers initialised, before any multicast / DSMEM traffic
    if (p.dbg && threadIdx.x == 0) p.dbg[((size_t)blockIdx.x * 8 + 7) * 8 + 2] = pk_gtimer();   // prologue done
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== TMA producer: runs ahead across tiles =====================
        if (lane == 0) {
            uint32_t kc = 0;
            const int rows = PK_M / p.nslice;   // the A tile is fetched as nslice (power of two <= nc) multicast slices
            for (int t = cluster_id; t < p.m_tiles; t += p.num_clusters) {
                const int m0 = t * PK_M;
                for (int kb = 0; kb < p.num_kb; ++kb, ++kc) {
                    const int s = kc % p.stages;
                    const uint32_t ph = (kc / p.stages) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    unsigned char *sa = smem + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full_bar[s], a_bytes + b_bytes);
                    if (p.nc > 1) {
                        if ((int)rank < p.nslice)
                            tma_load_2d_mc(&tmA, &full_bar[s], sa + (size_t)rank * rows * (PK_K * 2), kb * PK_K, m0 + (int)rank * rows, mc_mask);
                    } else tma_load_2d(&tmA, &full_bar[s], sa, kb * PK_K, m0);
                    tma_load_2d(&tmB, &full_bar[s], sa + a_bytes, kb * PK_K, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: alternates between the two accumulators =====================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(PK_M >> 4) << 24);
        uint32_t kc = 0;
        int it = 0;
        for (int t = cluster_id; t < p.m_tiles; t += p.num_clusters, ++it) {
            const int acc = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            mbar_wait(&tmem_empty_bar[acc], aph ^ 1u);   // epilogue has drained this accumulator (passes on first use)
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(acc * PK_ACC_STRIDE);
            for (int kb = 0; kb < p.num_kb; ++kb, ++kc) {
                const int s = kc % p.stages;
                const uint32_t ph = (kc / p.stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sa + a_bytes);
#pragma unroll
                    for (int k = 0; k < PK_K / 16; ++k)
                        umma_bf16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                    if (p.nc > 1) umma_commit_mc(&empty_bar[s], mc_mask); else umma_commit(&empty_bar[s]);
                    if (kb == p.num_kb - 1) umma_commit(&tmem_full_bar[acc]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int et = threadIdx.x - 128;
        const int grp = (warp - 4) >> 2;                   // column group: 16-column chunks grp, grp+NG, grp+2NG, ..
        const int row = (warp & 3) * 32 + lane;            // accumulator row == TMEM lane (quarter = warp % 4)
        const float LOG2E = 1.4426950408889634f;
        const int obj_c = p.N - 1 - n0;                    // tile column of the background class, if in this tile
        const int pieces = p.block_n / 8;
        int it = 0;
        for (int t = cluster_id; t < p.m_tiles; t += p.num_clusters, ++it) {
            const int acc = it & 1, buf = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            const int m0 = t * PK_M;
            const bool row_ok = (m0 + row) < p.M;
            float rs = p.scale;
            if (p.inv_nx && row_ok) rs *= __ldg(p.inv_nx + m0 + row);
            const float a2 = rs * LOG2E;
            const uint32_t trow = tmem_base + (uint32_t)(acc * PK_ACC_STRIDE) + ((uint32_t)((warp & 3) * 32) << 16);
            float rmax = -INFINITY, rsum = 0.f, eobj = 0.f;
            float cm[PK_CPG];
#pragma unroll
            for (int ci = 0; ci < PK_CPG; ++ci) cm[ci] = -INFINITY;

            PSTAMP(0);
            mbar_wait(&tmem_full_bar[acc], ph);
            tc_fence_after();
            PSTAMP(1);
            // ---- the only pass over TMEM: online (max, sum-exp); 2^(y - m_c) goes to the shared tile as bf16
#pragma unroll
            for (int ci = 0; ci < PK_CPG; ++ci) {
                const int c0 = (grp + PK_NG * ci) * 16;
                if (c0 >= p.block_n) continue;             // block_n is a multiple of 16: chunks are never partial in width
                const int nv = min(16, p.N - (n0 + c0));   // valid columns (warp-uniform)
                float v[16];
                if (nv > 0) {
                    tmem_ld16(trow + (uint32_t)c0, v);
                    if (p.inv_nt) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 cs = *reinterpret_cast<const float4 *>(colscale + c0 + i);
                            v[i] *= cs.x; v[i + 1] *= cs.y; v[i + 2] *= cs.z; v[i + 3] *= cs.w;
                        }
                    }
                    if (p.logits && row_ok) {
                        float *o = p.logits + (size_t)(m0 + row) * p.ld_logits + n0 + c0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) if (i < nv) o[i] = v[i] * rs;
                    }
                    float cmax = -INFINITY;
                    if (nv == 16) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) { v[i] *= a2; cmax = fmaxf(cmax, v[i]); }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) { v[i] = i < nv ? v[i] * a2 : -INFINITY; cmax = fmaxf(cmax, v[i]); }
                    }
                    const float nm = fmaxf(rmax, cmax);
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        v[i] = fast_exp2(v[i] - nm); v[i + 1] = fast_exp2(v[i + 1] - nm);
                        acc0 += v[i]; acc1 += v[i + 1];
                    }
                    rsum = rsum * fast_exp2(rmax - nm) + (acc0 + acc1);
                    rmax = nm;
                    cm[ci] = nm;
                    if (p.objectness && obj_c >= c0 && obj_c < c0 + 16) {   // warp-uniform: fp32 copy of the background exponential
                        float x = tmem_ld1(trow + (uint32_t)obj_c);
                        if (p.inv_nt) x *= colscale[obj_c];
                        eobj = fast_exp2(x * a2 - nm);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                if (p.prob) {
#pragma unroll
                    for (int i = 0; i < 16; i += 8) {
                        __nv_bfloat162 h[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
                        *reinterpret_cast<uint4 *>(tile + (size_t)row * tile_ld + c0 + i) = *reinterpret_cast<uint4 *>(h);
                    }
                }
            }
            // accumulator drained: hand it back to the MMA warp before the (slow) cluster exchange
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_local(&tmem_empty_bar[acc]);

            PSTAMP(2);
            part[grp][row] = make_float2(rmax, rsum);
            epi_barrier();
            {   // merge the column groups; group g publishes to CTAs g, g+4, .. of the cluster (DSMEM store completing tx bytes)
                float m = -INFINITY;
#pragma unroll
                for (int g = 0; g < PK_NG; ++g) m = fmaxf(m, part[g][row].x);
                float sm = 0.f;
                if (m > -INFINITY) {
#pragma unroll
                    for (int g = 0; g < PK_NG; ++g) { const float2 sg = part[g][row]; if (sg.x > -INFINITY) sm += sg.y * fast_exp2(sg.x - m); }
                }
                if (et == 0) mbar_expect_tx(&stats_full_bar[buf], (uint32_t)p.nc * PK_M * 8u);   // nc sources x 128 rows x 8 B land here
                for (int r = grp; r < p.nc; r += PK_NG)
                    st_async_f32x2(&stats[buf][rank][row], &stats_full_bar[buf], (uint32_t)r, m, sm);
            }
            PSTAMP(3);
            mbar_wait_cluster(&stats_full_bar[buf], ph);   // all nc*128 (source, row) pairs have landed here
            PSTAMP(4);
            float gmax = -INFINITY;
            for (int r = 0; r < p.nc; ++r) gmax = fmaxf(gmax, stats[buf][r][row].x);
            float gsum = 0.f;
            for (int r = 0; r < p.nc; ++r) {
                const float2 sr = stats[buf][r][row];
                if (sr.x > -INFINITY) gsum += sr.y * fast_exp2(sr.x - gmax);
            }
            const float inv = 1.f / gsum;
#pragma unroll
            for (int ci = 0; ci < PK_CPG; ++ci) {
                const int j = grp + PK_NG * ci, c0 = j * 16;
                if (c0 >= p.block_n) continue;
                const float f = cm[ci] > -INFINITY ? fast_exp2(cm[ci] - gmax) * inv : 0.f;
                ftab[row][j] = f;
                if (p.objectness && row_ok && obj_c >= c0 && obj_c < c0 + 16) p.objectness[m0 + row] = 1.f - eobj * f;
            }
            epi_barrier();   // tile + factor table complete
            PSTAMP(5);
            if (p.prob) {    // coalesced copy-out, rescaling 8 bf16 at a time: one row per warp and step, one 16-byte piece per lane
                const int ew = warp - 4;
                for (int c8 = lane * 8; c8 < p.block_n; c8 += 256) {     // block_n <= 256: a single trip
                    const int gcol = n0 + c8;
                    if (gcol >= p.ld_prob) continue;
#pragma unroll 4
                    for (int r = ew; r < PK_M; r += PK_EPI / 32) {
                        const int grow = m0 + r;
                        if (grow >= p.M) break;
                        const float f = ftab[r][c8 >> 4];
                        uint4 val = *reinterpret_cast<const uint4 *>(tile + (size_t)r * tile_ld + c8);
                        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&val);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 e = __bfloat1622float2(h[j]);
                            h[j] = __floats2bfloat162_rn(e.x * f, e.y * f);
                        }
                        *reinterpret_cast<uint4 *>(p.prob + (size_t)grow * p.ld_prob + gcol) = val;
                    }
                }
            }
            PSTAMP(6);
            epi_barrier();   // tile, ftab and part are free for the next M-tile
            PSTAMP(7);
        }
    }
    // teardown: nobody may exit while a peer can still multicast-arrive on its barriers
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
    if (p.dbg && threadIdx.x == 0) p.dbg[((size_t)blockIdx.x * 8 + 7) * 8 + 1] = pk_gtimer();   // kernel exit
}

int make_bf16_map(CUtensorMap *map, const void *base, int rows, int K, int box_rows);   // clip_logits.cu

// shared-memory plan of one (nc, bn) configuration
struct PkPlan { int stages; size_t stage_bytes, tile_bytes, smem; };
static bool pk_plan(int bn, int num_kb, PkPlan *pl)
{
    pl->stage_bytes = (size_t)PK_M * PK_K * 2 + (((size_t)bn * PK_K * 2 + 1023) & ~(size_t)1023);
    pl->tile_bytes = (size_t)PK_M * (bn + 8) * 2;
    const size_t budget = 227 * 1024 - 36 * 1024 - 1024 - pl->tile_bytes;   // static: barriers + stats 16K + part <= 7K + ftab 8.5K + colscale
    int stages = (int)(budget / pl->stage_bytes);
    if (stages > PK_MAX_STAGES) stages = PK_MAX_STAGES;
    if (stages > num_kb * 2) stages = num_kb * 2;
    pl->stages = stages;
    pl->smem = pl->stage_bytes * stages + pl->tile_bytes + 1024;
    return stages >= 2;
}

static void pk_launch_cfg(cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attr, int nc, int clusters, size_t smem, cudaStream_t st)
{
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3((unsigned)(clusters * nc));
    cfg->blockDim = dim3(PK_THREADS);
    cfg->dynamicSmemBytes = smem;
    cfg->stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr; cfg->numAttrs = 1;
}

// A persistent grid must be fully co-resident, and a cluster is placed inside one GPC: fewer than 148/nc clusters fit
// (measured on B200: 15 of 8, 18 of 7, 22 of 6, 26 of 5).  Ask the driver, once per (nc, smem).
static int pk_max_clusters(int nc, size_t smem)
{
    static int cache[PK_MAX_NC + 1] = {0};
    static size_t cache_smem[PK_MAX_NC + 1] = {0};
    if (cache[nc] && cache_smem[nc] == smem) return cache[nc];
    if (cudaFuncSetAttribute(clip_logits_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    pk_launch_cfg(&cfg, attr, nc, 148 / nc, smem, nullptr);
    int mc = 0;
    if (cudaOccupancyMaxActiveClusters(&mc, clip_logits_persistent_kernel, &cfg) != cudaSuccess || mc <= 0) { cudaGetLastError(); mc = 0; }
    cache[nc] = mc; cache_smem[nc] = smem;
    return mc;
}

// Pick the cluster size for the persistent kernel: rounds of M-tiles x (tile width + fixed per-tile cost), the epilogue
// being the per-tile bound.  Returns 0 when the persistent kernel should not be used (fewer than two rounds of work).
int clip_logits_persistent_choose(int M, int K, int N, int *nc_out, int *bn_out)
{
    const int m_tiles = (M + PK_M - 1) / PK_M;
    static int force = -1;
    if (force < 0) { const char *e = getenv("OVDET_LOGITS_NC"); force = e ? atoi(e) : 0; }
    long best = -1;
    for (int nc = 1; nc <= PK_MAX_NC; ++nc) {
        if (force >= 1 && force <= PK_MAX_NC && nc != force) continue;
        const int bn = (((N + nc - 1) / nc) + 15) / 16 * 16;
        if (bn > 256 || bn < 16) continue;
        PkPlan pl;
        if (!pk_plan(bn, K / PK_K, &pl)) continue;
        const int clusters = pk_max_clusters(nc, pl.smem);
        if (clusters <= 0) continue;
        const int used = clusters < m_tiles ? clusters : m_tiles;
        const int rounds = (m_tiles + used - 1) / used;
        if (rounds < 2) continue;
        const long cost = (long)rounds * (bn + 80);
        if (best < 0 || cost < best) { best = cost; *nc_out = nc; *bn_out = bn; }
    }
    return best >= 0;
}

int clip_logits_persistent_launch(const void *x, const void *text, int M, int K, int N, int nc, int bn, unsigned flags, float scale,
                                  float *logits, int ld_logits, void *prob, int ld_prob, float *objectness,
                                  const float *inv_nx, const float *inv_nt, cudaStream_t st)
{
    PLogitsParams p;
    p.M = M; p.N = N; p.K = K; p.block_n = bn; p.nc = nc; p.num_kb = K / PK_K;
    p.m_tiles = (M + PK_M - 1) / PK_M;
    p.nslice = 1;
    while (p.nslice * 2 <= nc) p.nslice *= 2;
    p.num_clusters = 148 / nc;
    if (p.num_clusters > p.m_tiles) p.num_clusters = p.m_tiles;
    p.scale = scale; p.logits = logits; p.ld_logits = ld_logits; p.prob = static_cast<__nv_bfloat16 *>(prob); p.ld_prob = ld_prob;
    p.objectness = objectness; p.inv_nx = inv_nx; p.inv_nt = inv_nt;
    { const char *e = getenv("OVDET_LOGITS_DBG_PTR"); p.dbg = e ? reinterpret_cast<unsigned long long *>(strtoull(e, nullptr, 0)) : nullptr; }
    PkPlan pl;
    if (!pk_plan(bn, p.num_kb, &pl)) { set_error("clip_logits: tile does not fit the persistent kernel"); return OVDET_ERR_UNSUPPORTED; }
    p.stages = pl.stages;
    CUtensorMap tmA, tmB;
    int rc = make_bf16_map(&tmA, x, M, K, PK_M / p.nslice);
    if (rc) return rc;
    rc = make_bf16_map(&tmB, text, N, K, bn);
    if (rc) return rc;
    const int mc = pk_max_clusters(nc, pl.smem);
    if (mc <= 0) { set_error("clip_logits: persistent kernel cannot be made resident"); return OVDET_ERR_UNSUPPORTED; }
    if (p.num_clusters > mc) p.num_clusters = mc;
    OVDET_CUDA_TRY(cudaFuncSetAttribute(clip_logits_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    pk_launch_cfg(&cfg, attr, nc, p.num_clusters, pl.smem, st);
    OVDET_CUDA_TRY(cudaLaunchKernelEx(&cfg, clip_logits_persistent_kernel, tmA, tmB, p));
    return launch_ok("clip_logits_persistent_kernel");
}

}  // namespace ovdet
