// core.cu — process-wide state of libowrx_b200.so: error string, launch counter, device selection.
#include "common.cuh"

namespace owrx {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int select_device(int device, int* sm_count)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(OWRX_E_CUDA, "no CUDA device available (%s); libowrx_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(OWRX_E_INVALID, "device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    OWRX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(OWRX_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                    prop.minor);
    OWRX_CUDA(cudaSetDevice(device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return OWRX_OK;
}

bool host_is_pinned(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace owrx

extern "C" {

const char* owrx_last_error(void) { return owrx::g_err; }
const char* owrx_version(void) { return "owrx_b200 0.1.0 (sm_100a)"; }
uint64_t owrx_launch_count(void) { return owrx::g_launches.load(); }

int owrx_pinned_alloc(size_t bytes, void** out)
{
    if (!out || !bytes) return owrx::fail(OWRX_E_INVALID, "bad argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { *out = nullptr; cudaGetLastError(); return owrx::fail(OWRX_E_NOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e)); }
    return OWRX_OK;
}

void owrx_pinned_free(void* p) { if (p) cudaFreeHost(p); }

int owrx_host_is_pinned(const void* p) { return owrx::host_is_pinned(p) ? 1 : 0; }

int owrx_device_count(int* n)
{
    if (!n) return owrx::fail(OWRX_E_INVALID, "n is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *n = 0; return owrx::fail(OWRX_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *n = c;
    return OWRX_OK;
}

}
