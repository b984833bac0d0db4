// ingress.cuh — IQ ingress formats.  The reference converts non-float sources on the CPU before its ring:
// Chain([Convert(Format.COMPLEX_SHORT, Format.COMPLEX_FLOAT), Gain(Format.COMPLEX_FLOAT, 5.0)]) (owrx/source/fifi_sdr.py:27-28,
// wired by owrx/source/direct.py:59-71).  Here the raw samples cross PCIe as they arrive (2 or 4 bytes per complex sample
// instead of 8) and are converted on the GPU while the next chunk is in flight.
//   csdr Convert short -> float:  y = (float)x / 32767        (SHRT_MAX)
//   csdr Convert uchar -> float:  y = (float)x / 127.5 - 1
//   Gain(g):                      y *= g                      (skipped when g == 1)
#pragma once
#include "common.cuh"

namespace owrx {

inline size_t iq_format_bytes(int fmt)
{
    return fmt == OWRX_IQ_CS16 ? 4 : (fmt == OWRX_IQ_CU8 ? 2 : (fmt == OWRX_IQ_CF32 ? 8 : 0));
}

template <int FMT>
__global__ void __launch_bounds__(256) iq_convert_kernel(const void* __restrict__ raw, float2* __restrict__ out, size_t n, float gain)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float re, im;
    if (FMT == OWRX_IQ_CS16) {
        const short2 v = reinterpret_cast<const short2*>(raw)[i];
        re = __fdiv_rn((float)v.x, 32767.0f);
        im = __fdiv_rn((float)v.y, 32767.0f);
    } else {
        const uchar2 v = reinterpret_cast<const uchar2*>(raw)[i];
        re = __fdiv_rn((float)v.x, 127.5f) - 1.0f;
        im = __fdiv_rn((float)v.y, 127.5f) - 1.0f;
    }
    if (gain != 1.0f) { re *= gain; im *= gain; }
    out[i] = make_float2(re, im);
}

// raw (device, `fmt`) -> complex float32 on stream st
inline int iq_convert_launch(int fmt, const void* d_raw, float2* d_out, size_t n, float gain, cudaStream_t st)
{
    if (!n) return OWRX_OK;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (fmt == OWRX_IQ_CS16) iq_convert_kernel<OWRX_IQ_CS16><<<grid, 256, 0, st>>>(d_raw, d_out, n, gain);
    else if (fmt == OWRX_IQ_CU8) iq_convert_kernel<OWRX_IQ_CU8><<<grid, 256, 0, st>>>(d_raw, d_out, n, gain);
    else return fail(OWRX_E_INVALID, "unknown IQ format %d", fmt);
    OWRX_LAUNCH_CHECK();
    return OWRX_OK;
}

}  // namespace owrx
