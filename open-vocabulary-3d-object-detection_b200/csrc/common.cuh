// common.cuh -- shared device/host helpers for libovdet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ovdet_b200.h"

namespace ovdet {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define OVDET_CUDA_TRY(expr)                                  \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return ::ovdet::cuda_fail(_e, #expr); \
    } while (0)

#define OVDET_REQUIRE(cond, msg)                              \
    do {                                                      \
        if (!(cond)) {                                        \
            ::ovdet::set_error("invalid argument: %s (%s)", msg, #cond); \
            return OVDET_ERR_INVALID;                         \
        }                                                     \
    } while (0)

static inline int launch_ok(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return OVDET_OK;
}

// Grow-only device scratch owned by the library, one per (host thread, device), used only
// by the *_host entry points (the device-pointer API never allocates).
struct HostStaging {
    void *dev = nullptr;
    size_t cap = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_d2h = nullptr;   // second stream: result read-back of chunk i overlaps upload + kernel of chunk i+1
    cudaEvent_t ev[8] = {};
    void *pinned = nullptr;      // grow-only pinned host buffer: small host arguments are packed here and travel in ONE copy
    size_t pinned_cap = 0;
    int ensure(size_t bytes);
    int ensure_pinned(size_t bytes);
};
HostStaging &host_staging();

// Grow-only device scratch for kernels that need a small internal table (logits row norms, AP bin edges): one per
// (host thread, device), reused across calls.  Captured CUDA graphs pin the buffer: once handed out inside a stream
// capture it can no longer grow (acquire() then fails loudly instead of freeing what a replay would touch).  cudaMallocAsync is not used for this: with the default pool's release threshold of
// zero every synchronisation returns the memory to the OS and the next call pays a multi-millisecond re-map.
// Stream order across calls is kept with an event: acquire() makes `stream` wait for the previous user, release()
// records the new last use.
constexpr int OVDET_MAX_DEVICES = 64;   // library scratch / staging objects are kept per (host thread, device)
struct DeviceScratch {
    void *dev = nullptr;
    size_t cap = 0;
    cudaEvent_t last_use = nullptr;
    bool pinned_by_graph = false;   // handed out inside a stream capture: the address is baked into a graph, never free it
    int acquire(size_t bytes, cudaStream_t stream, void **out);
    int release(cudaStream_t stream);
};
DeviceScratch &device_scratch();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when a kernel needs MORE than what was already set for it on the
// current device (capi.cu): the attribute call is ~1 us of host time, and the AP path launches five kernels per evaluation.
cudaError_t ensure_dyn_smem_impl(const void *func, size_t bytes);
template <typename F> inline cudaError_t ensure_dyn_smem(F func, size_t bytes) { return ensure_dyn_smem_impl(reinterpret_cast<const void *>(func), bytes); }

// ------------------------------------------------ reference-faithful arithmetic
// The reference evaluates every arithmetic op separately (Python floats, eager
// torch, numpy), so the geometry kernels must not contract a*b+c into an FMA:
// all maths goes through these round-to-nearest intrinsics, which nvcc never fuses.
template <typename T> struct Ar;
template <> struct Ar<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float rcp(float a) { return __frcp_rn(a); }  // == 1.0f / a, IEEE
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
};
template <> struct Ar<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double rcp(double a) { return __drcp_rn(a); }  // correctly rounded == 1.0 / a
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __device__ __forceinline__ double max(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double min(double a, double b) { return fmin(a, b); }
};

template <typename T> struct V2 { T x, y; };

// ---------------------------------------------------------------------------
// Sutherland-Hodgman clip of one convex quad by another, restating
// utils/box_intersection.pyx:27-70 (== utils/box_util.py:404-440): clip edges in
// order (c[3]->c[0], c[0]->c[1], ...), strict `>` inside test (pyx:23-24),
// intersection formula pyx:13-19 with n3 = 1/(..) then multiply, stop when the
// polygon becomes empty.  Vertex lists live in a per-thread shared-memory
// scratch (two ping-pong rows of MAXV vertices, element i of thread t at
// buf[i*STRIDE + t]: conflict-free), the subject quad of pass 0 and the output
// of pass 3 stay in registers.  A convex quad clipped by 4 half-planes has at
// most 8 vertices; MAXV=8 saturates beyond that (non-convex input only).
// The last pass streams its vertices into `Sink` (shoelace accumulator).
// ---------------------------------------------------------------------------
constexpr int SH_MAXV = 8;

template <typename T> struct ClipEdge {
    T c1x, c1y, ex, ey, dcx, dcy, n1;
    __device__ __forceinline__ ClipEdge(T ax, T ay, T bx, T by)
    {
        using A = Ar<T>;
        c1x = ax; c1y = ay;
        ex = A::sub(bx, ax); ey = A::sub(by, ay);
        dcx = A::sub(ax, bx); dcy = A::sub(ay, by);
        n1 = A::sub(A::mul(ax, by), A::mul(ay, bx));
    }
    __device__ __forceinline__ bool inside(T px, T py) const
    {
        using A = Ar<T>;
        return A::mul(ex, A::sub(py, c1y)) > A::mul(ey, A::sub(px, c1x));
    }
    __device__ __forceinline__ V2<T> isect(T sx, T sy, T px, T py) const
    {
        using A = Ar<T>;
        const T dpx = A::sub(sx, px), dpy = A::sub(sy, py);
        const T n2 = A::sub(A::mul(sx, py), A::mul(sy, px));
        const T n3 = A::rcp(A::sub(A::mul(dcx, dpy), A::mul(dcy, dpx)));
        V2<T> r;
        r.x = A::mul(A::sub(A::mul(n1, dpx), A::mul(n2, dcx)), n3);
        r.y = A::mul(A::sub(A::mul(n1, dpy), A::mul(n2, dcy)), n3);
        return r;
    }
};

// Shoelace areas of the clipped polygon left in shared scratch (vertex i of this
// thread at poly[i*STRIDE]).  All follow the reference expression
// 0.5*|dot(x, roll(y,1)) - dot(y, roll(x,1))| term by term from i = 0 (roll(.,1)[0]
// is the LAST vertex), because the fp32 sums are ill-conditioned in absolute
// coordinates (|x*y| ~ 10 against areas ~ 1e-2..1) and the order is visible at 1e-6.
template <int STRIDE>
__device__ __forceinline__ float area_f32(const V2<float> *poly, int n)
{   // fp32 torch path: utils/box_util.py:591-598
    if (n <= 0) return 0.f;
    float d1 = 0.f, d2 = 0.f;
    V2<float> pv = poly[(n - 1) * STRIDE];
    for (int i = 0; i < n; ++i) {
        const V2<float> v = poly[i * STRIDE];
        d1 = __fadd_rn(d1, __fmul_rn(v.x, pv.y));
        d2 = __fadd_rn(d2, __fmul_rn(v.y, pv.x));
        pv = v;
    }
    return __fmul_rn(fabsf(__fsub_rn(d1, d2)), 0.5f);
}
template <int STRIDE>
__device__ __forceinline__ float area_cython(const V2<double> *poly, int n)
{   // pyx:196-198: polygon cast to fp32, fp32 products accumulated in double (np.dot), fp32 result
    if (n <= 0) return 0.f;
    double d1 = 0.0, d2 = 0.0;
    const V2<double> l = poly[(n - 1) * STRIDE];
    float px = (float)l.x, py = (float)l.y;
    for (int i = 0; i < n; ++i) {
        const V2<double> v = poly[i * STRIDE];
        const float x = (float)v.x, y = (float)v.y;
        d1 = __dadd_rn(d1, (double)__fmul_rn(x, py));
        d2 = __dadd_rn(d2, (double)__fmul_rn(y, px));
        px = x; py = y;
    }
    return __fmul_rn(0.5f, fabsf(__fsub_rn((float)d1, (float)d2)));
}
template <int STRIDE>
__device__ __forceinline__ double area_f64(const V2<double> *poly, int n)
{   // poly_area (box_util.py:84-86); stands in for Qhull's hull area (:96); <3 vertices -> 0
    if (n < 3) return 0.0;
    double d1 = 0.0, d2 = 0.0;
    V2<double> pv = poly[(n - 1) * STRIDE];
    for (int i = 0; i < n; ++i) {
        const V2<double> v = poly[i * STRIDE];
        d1 = __dadd_rn(d1, __dmul_rn(v.x, pv.y));
        d2 = __dadd_rn(d2, __dmul_rn(v.y, pv.x));
        pv = v;
    }
    return __dmul_rn(0.5, fabs(__dsub_rn(d1, d2)));
}

// ---------------------------------------------------------------------------
// Cooperative clip: EIGHT lanes per pair, lane i of the group holds vertex i of the current polygon in registers.
// One pass = one inside test and at most one intersection per lane, all lanes in parallel, then a compaction through
// an 8-entry shared row (output order = the serial loop's: for each vertex, [intersection], [vertex]; entries past
// SH_MAXV dropped exactly like the serial `m < SH_MAXV` guards).  Per-element arithmetic is the serial code's, so the
// polygon is bit-identical; the dependent chain per pass is one vertex long instead of n, which is what matters when
// only a handful of pairs per CTA reach the clipper (the fp64 serial clip is a 4-5 us chain, profiles/r1_notes.md).
// All 32 lanes of the warp must call these together (full-mask shuffles of width 8); n is uniform per group.
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int coop_pass(const ClipEdge<T> &ce, T &vx, T &vy, int n, int gl, int gshift, V2<T> *gbuf)
{
    const unsigned full = 0xffffffffu;
    const bool live = gl < n;
    const bool e_in = live && ce.inside(vx, vy);
    const int src = max((gl == 0 ? n : gl) - 1, 0);   // predecessor in the cyclic order
    const T sx = __shfl_sync(full, vx, src, 8), sy = __shfl_sync(full, vy, src, 8);
    const unsigned inmask = (__ballot_sync(full, e_in) >> gshift) & 0xffu;
    const bool s_in = (inmask >> src) & 1u;
    const bool cross = live && (e_in != s_in);
    const unsigned crmask = (__ballot_sync(full, cross) >> gshift) & 0xffu;
    const unsigned lower = (1u << gl) - 1u;
    const int pos = __popc(crmask & lower) + __popc(inmask & lower);
    if (cross && pos < SH_MAXV) gbuf[pos] = ce.isect(sx, sy, vx, vy);
    const int pe = pos + (cross ? 1 : 0);
    if (e_in && pe < SH_MAXV) { V2<T> v; v.x = vx; v.y = vy; gbuf[pe] = v; }
    const int m = min(__popc(crmask) + __popc(inmask), SH_MAXV);
    __syncwarp();
    if (gl < m) { const V2<T> v = gbuf[gl]; vx = v.x; vy = v.y; }
    __syncwarp();
    return m;
}

// subject vertex gl (< 4) in (vx, vy); clip quad cl[8]; returns n with the polygon left in the lanes' registers
template <typename T>
__device__ __forceinline__ int coop_clip_quads(const T *cl, T &vx, T &vy, int n, int gl, int gshift, V2<T> *gbuf)
{
    n = coop_pass(ClipEdge<T>(cl[6], cl[7], cl[0], cl[1]), vx, vy, n, gl, gshift, gbuf);
    n = coop_pass(ClipEdge<T>(cl[0], cl[1], cl[2], cl[3]), vx, vy, n, gl, gshift, gbuf);
    n = coop_pass(ClipEdge<T>(cl[2], cl[3], cl[4], cl[5]), vx, vy, n, gl, gshift, gbuf);
    n = coop_pass(ClipEdge<T>(cl[4], cl[5], cl[6], cl[7]), vx, vy, n, gl, gshift, gbuf);
    return n;
}

// the two reference shoelace flavours on a register-resident polygon: products in parallel, sums in the reference's order
__device__ __forceinline__ float coop_area_f32(float vx, float vy, int n, int gl)
{
    const unsigned full = 0xffffffffu;
    const int src = max((gl == 0 ? n : gl) - 1, 0);
    const float px = __shfl_sync(full, vx, src, 8), py = __shfl_sync(full, vy, src, 8);
    const float t1 = __fmul_rn(vx, py), t2 = __fmul_rn(vy, px);
    float d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int i = 0; i < SH_MAXV; ++i) {
        const float a = __shfl_sync(full, t1, i, 8), b = __shfl_sync(full, t2, i, 8);
        if (i < n) { d1 = __fadd_rn(d1, a); d2 = __fadd_rn(d2, b); }
    }
    return n <= 0 ? 0.f : __fmul_rn(fabsf(__fsub_rn(d1, d2)), 0.5f);
}
__device__ __forceinline__ float coop_area_cython(double vx, double vy, int n, int gl)
{
    const unsigned full = 0xffffffffu;
    const int src = max((gl == 0 ? n : gl) - 1, 0);
    const float x = (float)vx, y = (float)vy;
    const float px = __shfl_sync(full, x, src, 8), py = __shfl_sync(full, y, src, 8);
    const float t1 = __fmul_rn(x, py), t2 = __fmul_rn(y, px);   // fp32 products, widened exactly
    double d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int i = 0; i < SH_MAXV; ++i) {
        const float a = __shfl_sync(full, t1, i, 8), b = __shfl_sync(full, t2, i, 8);
        if (i < n) { d1 = __dadd_rn(d1, (double)a); d2 = __dadd_rn(d2, (double)b); }
    }
    return n <= 0 ? 0.f : __fmul_rn(0.5f, fabsf(__fsub_rn((float)d1, (float)d2)));
}

// One pass over a vertex list held in shared scratch, output to shared scratch.
template <typename T, int STRIDE>
__device__ __forceinline__ int sh_pass_mem(const ClipEdge<T> &ce, const V2<T> *in, int n, V2<T> *out)
{
    V2<T> s = in[(n - 1) * STRIDE];
    bool s_in = ce.inside(s.x, s.y);
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const V2<T> e = in[i * STRIDE];
        const bool e_in = ce.inside(e.x, e.y);
        if (e_in != s_in && m < SH_MAXV) out[(m++) * STRIDE] = ce.isect(s.x, s.y, e.x, e.y);
        if (e_in && m < SH_MAXV) out[(m++) * STRIDE] = e;
        s = e; s_in = e_in;
    }
    return m;
}

// Full clip: subject quad s[4], clip quad c[4] (x,y interleaved, already of type T).
// bufA/bufB: this thread's scratch rows.  Returns the vertex count; the polygon is left in bufB.
template <typename T, int STRIDE>
__device__ __forceinline__ int sh_clip_quads(const T *s, const T *c, V2<T> *bufA, V2<T> *bufB)
{
    int n;
    {   // pass 0: subject in registers -> bufA
        const ClipEdge<T> ce(c[6], c[7], c[0], c[1]);
        T sx = s[6], sy = s[7];
        bool s_in = ce.inside(sx, sy);
        int m = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const T px = s[2 * i], py = s[2 * i + 1];
            const bool e_in = ce.inside(px, py);
            if (e_in != s_in) bufA[(m++) * STRIDE] = ce.isect(sx, sy, px, py);
            if (e_in) { V2<T> v; v.x = px; v.y = py; bufA[(m++) * STRIDE] = v; }
            sx = px; sy = py; s_in = e_in;
        }
        n = m;
    }
    if (n == 0) return 0;
    {
        const ClipEdge<T> ce(c[0], c[1], c[2], c[3]);
        n = sh_pass_mem<T, STRIDE>(ce, bufA, n, bufB);
    }
    if (n == 0) return 0;
    {
        const ClipEdge<T> ce(c[2], c[3], c[4], c[5]);
        n = sh_pass_mem<T, STRIDE>(ce, bufB, n, bufA);
    }
    if (n == 0) return 0;
    {
        const ClipEdge<T> ce(c[4], c[5], c[6], c[7]);
        n = sh_pass_mem<T, STRIDE>(ce, bufA, n, bufB);
    }
    return n;
}

// 16-byte read-only load
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

}  // namespace ovdet
