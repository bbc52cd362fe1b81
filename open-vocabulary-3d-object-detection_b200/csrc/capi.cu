// capi.cu -- library-level pieces of the C ABI: error reporting, version,
// device probe and the host-staging scratch used by the *_host entry points.
#include <stdarg.h>

#include <mutex>

#include "common.cuh"

namespace ovdet {

cudaError_t ensure_dyn_smem_impl(const void *func, size_t bytes)
{
    struct Ent { const void *f; int dev; size_t bytes; };
    static Ent tab[256];
    static int n = 0;
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> g(mu);
    for (int i = 0; i < n; ++i)
        if (tab[i].f == func && tab[i].dev == dev) {
            if (bytes <= tab[i].bytes) return cudaSuccess;
            e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e == cudaSuccess) tab[i].bytes = bytes;
            return e;
        }
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && n < 256) tab[n++] = Ent{func, dev, bytes};
    return e;
}


static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return OVDET_ERR_CUDA;
}

int HostStaging::ensure(size_t bytes)
{
    if (!stream) {
        OVDET_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        OVDET_CUDA_TRY(cudaStreamCreateWithFlags(&stream_d2h, cudaStreamNonBlocking));
        for (auto &e : ev) OVDET_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (bytes <= cap) return OVDET_OK;
    if (dev) { OVDET_CUDA_TRY(cudaStreamSynchronize(stream)); OVDET_CUDA_TRY(cudaFree(dev)); dev = nullptr; cap = 0; }
    size_t want = bytes + bytes / 2;
    OVDET_CUDA_TRY(cudaMalloc(&dev, want));
    cap = want;
    return OVDET_OK;
}

int DeviceScratch::acquire(size_t bytes, cudaStream_t stream, void **out)
{
    if (!last_use) OVDET_CUDA_TRY(cudaEventCreateWithFlags(&last_use, cudaEventDisableTiming));
    if (bytes > cap) {
        // growing frees the old buffer: a CUDA graph captured earlier has its address baked in, and a capture in progress
        // cannot synchronise on the event -- refuse rather than hand out memory a replay would find freed
        cudaStreamCaptureStatus cs0 = cudaStreamCaptureStatusNone;
        OVDET_CUDA_TRY(cudaStreamIsCapturing(stream, &cs0));
        if (cs0 != cudaStreamCaptureStatusNone) {
            set_error("the library scratch buffer would have to grow (%zu > %zu bytes) inside a stream capture; run the call once eagerly at this size before capturing", bytes, cap);
            return OVDET_ERR_UNSUPPORTED;
        }
        if (pinned_by_graph) {
            set_error("the library scratch buffer is referenced by a captured CUDA graph and cannot grow (%zu > %zu bytes); run the larger problem before capturing", bytes, cap);
            return OVDET_ERR_UNSUPPORTED;
        }
        if (dev) { OVDET_CUDA_TRY(cudaEventSynchronize(last_use)); OVDET_CUDA_TRY(cudaFree(dev)); dev = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 2 + 4096;
        OVDET_CUDA_TRY(cudaMalloc(&dev, want));
        cap = want;
    } else {
        // inside a stream capture the captured stream's own order is all there is (an outside event cannot be waited on)
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        OVDET_CUDA_TRY(cudaStreamIsCapturing(stream, &cs));
        if (cs == cudaStreamCaptureStatusNone) OVDET_CUDA_TRY(cudaStreamWaitEvent(stream, last_use, 0));
        else pinned_by_graph = true;   // the graph keeps the address: from now on the buffer never moves
    }
    *out = dev;
    return OVDET_OK;
}

int DeviceScratch::release(cudaStream_t stream)
{
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    OVDET_CUDA_TRY(cudaStreamIsCapturing(stream, &cs));
    if (cs == cudaStreamCaptureStatusNone) OVDET_CUDA_TRY(cudaEventRecord(last_use, stream));
    return OVDET_OK;
}

// Library state is per (host thread, CUDA device): a buffer allocated on device A must never be handed to a kernel
// running on device B (a process that drives several GPUs calls the same entry points under different current devices).
static int current_device_slot()
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; }
    return (d >= 0 && d < OVDET_MAX_DEVICES) ? d : 0;
}

DeviceScratch &device_scratch()
{
    static thread_local DeviceScratch ds[OVDET_MAX_DEVICES];
    return ds[current_device_slot()];
}

int HostStaging::ensure_pinned(size_t bytes)
{
    if (bytes <= pinned_cap) return OVDET_OK;
    if (pinned) { if (stream) OVDET_CUDA_TRY(cudaStreamSynchronize(stream)); OVDET_CUDA_TRY(cudaFreeHost(pinned)); pinned = nullptr; pinned_cap = 0; }
    const size_t want = bytes + bytes / 2 + 4096;
    OVDET_CUDA_TRY(cudaHostAlloc(&pinned, want, cudaHostAllocDefault));
    pinned_cap = want;
    return OVDET_OK;
}

HostStaging &host_staging()
{
    static thread_local HostStaging hs[OVDET_MAX_DEVICES];
    return hs[current_device_slot()];
}

}  // namespace ovdet

extern "C" int ovdet_version(void) { return 1; }
extern "C" const char *ovdet_last_error(void) { return ovdet::g_err; }
extern "C" int ovdet_stream_synchronize(void *stream)
{
    OVDET_CUDA_TRY(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
    return OVDET_OK;
}
extern "C" int ovdet_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- symmetric buffers (CUDA IPC): one allocation per rank, mapped into every peer process, so that kernels can
// exchange with plain stores over NVLink (csrc/ap_compact.cu, ovdet_apx_reduce)
extern "C" int ovdet_symm_alloc(size_t bytes, void **dev_ptr)
{
    OVDET_REQUIRE(dev_ptr && bytes > 0, "null pointer / zero size");
    void *p = nullptr;
    OVDET_CUDA_TRY(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(p); return ovdet::cuda_fail(e, "cudaMemset(symmetric buffer)"); }
    *dev_ptr = p;
    return OVDET_OK;
}

extern "C" int ovdet_symm_free(void *dev_ptr)
{
    if (dev_ptr) OVDET_CUDA_TRY(cudaFree(dev_ptr));
    return OVDET_OK;
}

extern "C" int ovdet_symm_export(void *dev_ptr, void *handle_out)
{
    OVDET_REQUIRE(dev_ptr && handle_out, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == OVDET_SYMM_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    OVDET_CUDA_TRY(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle_out, &h, sizeof(h));
    return OVDET_OK;
}

extern "C" int ovdet_symm_open(const void *handle, void **peer_ptr)
{
    OVDET_REQUIRE(handle && peer_ptr, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    OVDET_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_ptr = p;
    return OVDET_OK;
}

extern "C" int ovdet_symm_close(void *peer_ptr)
{
    if (peer_ptr) OVDET_CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
    return OVDET_OK;
}
