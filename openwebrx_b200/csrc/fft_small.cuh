// fft_small.cuh — register-resident small DFTs (radix 2/4/8/16) shared by the waterfall FFT (waterfall.cu) and the
// fast-convolution channeliser (fastconv.cu).
#pragma once
#include "common.cuh"

namespace owrx {

// ------------------------------------------------------------------------------------------------
// register-resident small DFTs.  After dftR(v) register q holds X[slot<R>(q)].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dft2(float2& a, float2& b)
{
    float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d)
{
    float2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = cmul_mi(csub(b, d));
    a = cadd(t0, t2);
    c = csub(t0, t2);
    b = cadd(t1, t3);
    d = csub(t1, t3);
}

template <int R> __device__ __forceinline__ int slot(int q);
template <> __device__ __forceinline__ int slot<1>(int q) { return q; }
template <> __device__ __forceinline__ int slot<2>(int q) { return q; }
template <> __device__ __forceinline__ int slot<4>(int q) { return q; }
template <> __device__ __forceinline__ int slot<8>(int q) { return (q >> 1) + 4 * (q & 1); }
template <> __device__ __forceinline__ int slot<16>(int q) { return (q >> 2) + 4 * (q & 3); }

#define OWRX_SQRT1_2 0.70710678118654752440f
#define OWRX_COS_PI_8 0.92387953251128675613f
#define OWRX_SIN_PI_8 0.38268343236508977173f

template <int R> __device__ __forceinline__ void dft(float2* v);
template <> __device__ __forceinline__ void dft<1>(float2*) {}
template <> __device__ __forceinline__ void dft<2>(float2* v) { dft2(v[0], v[1]); }
template <> __device__ __forceinline__ void dft<4>(float2* v) { dft4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ void dft<8>(float2* v)
{
    // n = 2 n1 + n2, m = m1 + 4 m2
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
    // v[2 m1 + 1] *= W8^m1
    v[3] = make_float2((v[3].x + v[3].y) * OWRX_SQRT1_2, (v[3].y - v[3].x) * OWRX_SQRT1_2);
    v[5] = cmul_mi(v[5]);
    v[7] = make_float2((v[7].y - v[7].x) * OWRX_SQRT1_2, -(v[7].x + v[7].y) * OWRX_SQRT1_2);
    dft2(v[0], v[1]);
    dft2(v[2], v[3]);
    dft2(v[4], v[5]);
    dft2(v[6], v[7]);
}
template <> __device__ __forceinline__ void dft<16>(float2* v)
{
    // n = 4 n1 + n2, m = m1 + 4 m2;  W16^{nm} = W4^{n1 m1} W16^{n2 m1} W4^{n2 m2}
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    const float2 w1 = make_float2(OWRX_COS_PI_8, -OWRX_SIN_PI_8);
    const float2 w3 = make_float2(OWRX_SIN_PI_8, -OWRX_COS_PI_8);
    // m1 = 1: exponents 1,2,3
    v[5] = cmul(v[5], w1);
    v[6] = make_float2((v[6].x + v[6].y) * OWRX_SQRT1_2, (v[6].y - v[6].x) * OWRX_SQRT1_2);
    v[7] = cmul(v[7], w3);
    // m1 = 2: exponents 2,4,6
    v[9] = make_float2((v[9].x + v[9].y) * OWRX_SQRT1_2, (v[9].y - v[9].x) * OWRX_SQRT1_2);
    v[10] = cmul_mi(v[10]);
    v[11] = make_float2((v[11].y - v[11].x) * OWRX_SQRT1_2, -(v[11].x + v[11].y) * OWRX_SQRT1_2);
    // m1 = 3: exponents 3,6,9
    v[13] = cmul(v[13], w3);
    v[14] = make_float2((v[14].y - v[14].x) * OWRX_SQRT1_2, -(v[14].x + v[14].y) * OWRX_SQRT1_2);
    v[15] = cmul(v[15], make_float2(-OWRX_COS_PI_8, OWRX_SIN_PI_8));
#pragma unroll
    for (int m1 = 0; m1 < 4; m1++) dft4(v[4 * m1], v[4 * m1 + 1], v[4 * m1 + 2], v[4 * m1 + 3]);
}

}  // namespace owrx
