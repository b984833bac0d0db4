// iq_hop.cu — the one exchange step of the multi-GPU path: the ingest GPU's wideband IQ block reaches every
// GPU that holds a shard of the client channels.  In the reference every DspManager attaches its own reader to the
// source ring (owrx/dsp.py:835-837, owrx/source/__init__.py:307-330); across GPUs that ring is replicated once per block.
//
// B200 / NVSwitch: the block is written ONCE through a multicast address (NVLS) — `multimem.st` stores leave the ingest GPU
// over NVLink a single time and the switch replicates them into the same offset of every member's buffer, so the hop costs
// one block of NVLink egress whatever the number of GPUs and no SM time on the receivers (an NCCL ring broadcast forwards
// the block GPU to GPU with copy kernels on every rank, beside the DSP kernels).  The multicast mapping itself is set up
// by the host (torch.distributed symmetric memory in bench.py); this file is the data path.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace owrx {
namespace {

__device__ __forceinline__ void multimem_st_v4(float4* mc, float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// grid-stride copy src (local HBM) -> multicast address; 4 independent 16-byte loads in flight per thread
__global__ void __launch_bounds__(512)
iq_multicast_kernel(const float4* __restrict__ src, float4* __restrict__ mc_dst, size_t n16)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        const float4 a = __ldg(src + i), b = __ldg(src + i + stride), c = __ldg(src + i + 2 * stride), d = __ldg(src + i + 3 * stride);
        multimem_st_v4(mc_dst + i, a);
        multimem_st_v4(mc_dst + i + stride, b);
        multimem_st_v4(mc_dst + i + 2 * stride, c);
        multimem_st_v4(mc_dst + i + 3 * stride, d);
    }
    for (; i < n16; i += stride) multimem_st_v4(mc_dst + i, __ldg(src + i));
    // make the block visible system-wide before whatever the host orders after this kernel (the "landed" barrier)
    __threadfence_system();
}

}  // namespace
}  // namespace owrx

extern "C" int owrx_iq_multicast_store(const void* src_dev, void* multicast_dst, size_t n_bytes, void* stream)
{
    using namespace owrx;
    if (!src_dev || !multicast_dst) return fail(OWRX_E_INVALID, "NULL argument");
    if ((n_bytes & 15) || ((uintptr_t)src_dev & 15) || ((uintptr_t)multicast_dst & 15))
        return fail(OWRX_E_INVALID, "multicast store needs 16-byte aligned pointers and size");
    if (!n_bytes) return OWRX_OK;
    // a few dozen CTAs saturate the NVLink port; more would only take SM time from the DSP kernels running beside the hop
    static const int ctas = getenv("OWRX_HOP_CTAS") ? std::max(1, atoi(getenv("OWRX_HOP_CTAS"))) : 32;
    const size_t n16 = n_bytes / 16;
    const unsigned grid = (unsigned)std::min<size_t>((size_t)ctas, (n16 + 511) / 512);
    iq_multicast_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src_dev),
                                                               reinterpret_cast<float4*>(multicast_dst), n16);
    OWRX_LAUNCH_CHECK();
    return OWRX_OK;
}
