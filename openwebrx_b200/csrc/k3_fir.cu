// k3_fir.cu — K3: batched NCO mix + polyphase FIR decimation (Shift + FirDecimate of every client channel).
// Own translation unit: built with `-Xptxas -O1` (see openwebrx_b200/_build.py) — at the default level the
// ptxas scheduler software-pipelines the FFMA2 loop and inserts ~24 register copies per sample on the
// already saturated FMA pipe; -O1 keeps all 56 accumulator pairs in place (88 instead of 108 instructions).
#define OWRX_K3_ONLY
#include "selector_kernels.cuh"

namespace owrx {

// Blackwell packed FP32: fma.rn.f32x2 -> SASS FFMA2 (two FMAs per issue slot; a scalar operand is
// broadcast by the hardware).  Accumulators of adjacent polyphase branches (p, p+1) share a 64-bit
// register pair, the tap pair comes straight out of the LDS.128, and the rotated sample component is
// the broadcast scalar: no register-bank conflict between the two fresh operands (both are aligned
// even/odd pairs) and half the issue slots of the scalar FFMA stream.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(f32x2& c, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b)); }

// Work decomposition: INPUT-stationary.  The launch covers n_blocks = n_k + 27 input blocks of D samples;
// CTA (range r, tap segment ts, r-split rs, channel group cg) streams blocks [r*JB, (r+1)*JB) and keeps,
// per channel, 28 rolling accumulators: accumulator p of block j belongs to output k = j - p.  Every FMA
// feeds some output (no warm-up waste): outputs whose 28 blocks straddle a range boundary get one
// partial sum from each of the two ranges (`partial` from the earlier one, `side` from the later one)
// and fir_reduce_kernel adds them.
__global__ void __launch_bounds__(K3_NW * 32, 1) fir_decimate_kernel(K3Params p)
{
    extern __shared__ float4 k3_smem[];
    float* hs = reinterpret_cast<float*>(k3_smem);                // [RB][28]
    float2* xs = reinterpret_cast<float2*>(hs + p.RB * K3_PP);    // [2][RB]
    float* red = reinterpret_cast<float*>(xs + 2 * p.RB);         // [2][NW][128]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bx = blockIdx.x;
    const int rs = bx % p.nrs; bx /= p.nrs;
    const int ts = bx % p.nseg;
    const int range = bx / p.nseg;
    const int cg = blockIdx.y;
    const int part = ts * p.nrs + rs;
    const int n_ranges = (p.n_blocks + p.JB - 1) / p.JB;

    const int r_lo = rs * p.RB;
    const int rcount = min(p.RB, p.D - r_lo);
    const int jb0 = range * p.JB;
    const int jb1 = min(p.n_blocks, jb0 + p.JB);
    if (jb0 >= jb1 || rcount <= 0) return;

    {   // tap slice of this CTA: rcount rows of 28
        const float4* src = reinterpret_cast<const float4*>(p.taps + ((size_t)ts * p.D + r_lo) * K3_PP);
        float4* dst = reinterpret_cast<float4*>(hs);
        for (int i = tid; i < rcount * (K3_PP / 4); i += K3_NW * 32) dst[i] = __ldg(src + i);
    }

    const int slot0 = cg * K3_CG + lane * K3_CN;
    const double rate0 = p.ch_rate[slot0], rate1 = p.ch_rate[slot0 + 1];
    const double ph0 = p.ch_phase[slot0], ph1 = p.ch_phase[slot0 + 1];
    const float2 w0 = p.ch_w[slot0], w1 = p.ch_w[slot0 + 1];

    const int SL = (rcount + K3_NW - 1) / K3_NW;
    const int i0 = min(rcount, warp * SL), i1 = min(rcount, i0 + SL);

    // pair j holds branches (2j, 2j+1)
    f32x2 a0r[K3_PP / 2], a0i[K3_PP / 2], a1r[K3_PP / 2], a1i[K3_PP / 2];
#pragma unroll
    for (int q = 0; q < K3_PP / 2; q++) { a0r[q] = 0ull; a0i[q] = 0ull; a1r[q] = 0ull; a1i[q] = 0ull; }

    auto load_tile = [&](int jl, int buf) {
        const long long s0 = (long long)(jl + ts * K3_PP) * p.D + r_lo;
        float2* dst = xs + buf * p.RB;
        for (int i = tid; i < rcount; i += K3_NW * 32) {
            const long long s = s0 + i;
            if (s < p.n_lim) cp_async8(dst + i, p.iq + s);
            else dst[i] = make_float2(0.f, 0.f);
        }
        cp_async_commit();
    };
    // sum the NW warp partials that completed after block jl_done -> output k = jl_done - 27
    auto flush = [&](int jl_done, int buf) {
        const int k = jl_done - (K3_PP - 1);
        if (k < 0 || k >= p.n_k || tid >= 128) return;
        const float* r = red + buf * (K3_NW * 128) + tid;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < K3_NW; w++) s += r[w * 128];
        float* out;
        if (k >= jb0) out = reinterpret_cast<float*>(p.partial + ((size_t)part * p.n_k + k) * p.slots + cg * K3_CG);
        else out = reinterpret_cast<float*>(p.side + (((size_t)part * n_ranges + range) * (K3_PP - 1) + (k - (jb0 - (K3_PP - 1)))) * p.slots + cg * K3_CG);
        out[tid] = s;
    };
    // emit the oldest branch (p = 27) into the cross-warp reduction buffer, then age every branch by one block
#define K3_ROLL(A)                                                                                      \
    {                                                                                                   \
        float plo, phi, clo, chi;                                                                       \
        _Pragma("unroll") for (int q = K3_PP / 2 - 1; q > 0; q--) {                                     \
            upk2(A[q], clo, chi); upk2(A[q - 1], plo, phi);                                             \
            A[q] = pk2(phi, clo);                                                                       \
        }                                                                                               \
        upk2(A[0], clo, chi);                                                                           \
        A[0] = pk2(0.f, clo);                                                                           \
    }
    auto emit_and_roll = [&](int buf) {
        float lo, e0, e1, e2, e3;
        upk2(a0r[K3_PP / 2 - 1], lo, e0); upk2(a0i[K3_PP / 2 - 1], lo, e1);
        upk2(a1r[K3_PP / 2 - 1], lo, e2); upk2(a1i[K3_PP / 2 - 1], lo, e3);
        reinterpret_cast<float4*>(red + buf * (K3_NW * 128) + warp * 128)[lane] = make_float4(e0, e1, e2, e3);
        K3_ROLL(a0r) K3_ROLL(a0i) K3_ROLL(a1r) K3_ROLL(a1i)
    };

    load_tile(jb0, 0);
    for (int jl = jb0; jl < jb1; jl++) {
        const int buf = (jl - jb0) & 1;
        cp_async_wait_all();
        __syncthreads();
        if (jl + 1 < jb1) load_tile(jl + 1, buf ^ 1);
        if (jl > jb0) flush(jl - 1, buf ^ 1);

        if (i0 < i1) {
            // re-seed both NCOs at the first sample of this warp's slice (double phase -> float sincos)
            const long long n_rel = (long long)(jl + ts * K3_PP) * p.D + r_lo + i0;
            double t0 = ph0 + rate0 * (double)(n_rel + 1), t1 = ph1 + rate1 * (double)(n_rel + 1);
            t0 -= floor(t0); t1 -= floor(t1);
            float2 q0, q1;
            sincospif(2.0f * (float)t0, &q0.y, &q0.x);
            sincospif(2.0f * (float)t1, &q1.y, &q1.x);
            const float2* xb = xs + buf * p.RB;
            const unsigned hs_base = (unsigned)__cvta_generic_to_shared(hs);
#pragma unroll 1
            for (int i = i0; i < i1; i++) {
                // taps as two-float pairs straight from shared memory (volatile: keeps ptxas from rotating the
                // loop, which costs ~40 register copies per sample)
                const unsigned haddr = hs_base + (unsigned)i * (K3_PP * 4);
                f32x2 hp[K3_PP / 2];
#pragma unroll
                for (int g = 0; g < K3_PP / 4; g++)
                    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(hp[2 * g]), "=l"(hp[2 * g + 1]) : "r"(haddr + g * 16));
                const float2 x = xb[i];
                const float2 z0 = cmul(x, q0), z1 = cmul(x, q1);
                q0 = cmul(q0, w0); q1 = cmul(q1, w1);
                const f32x2 z0x = pk2(z0.x, z0.x), z0y = pk2(z0.y, z0.y), z1x = pk2(z1.x, z1.x), z1y = pk2(z1.y, z1.y);
#pragma unroll
                for (int j = 0; j < K3_PP / 2; j++) ffma2(a0r[j], z0x, hp[j]);
#pragma unroll
                for (int j = 0; j < K3_PP / 2; j++) ffma2(a0i[j], z0y, hp[j]);
#pragma unroll
                for (int j = 0; j < K3_PP / 2; j++) ffma2(a1r[j], z1x, hp[j]);
#pragma unroll
                for (int j = 0; j < K3_PP / 2; j++) ffma2(a1i[j], z1y, hp[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < K3_PP / 2; j++) asm volatile("" : "+l"(a0r[j]), "+l"(a0i[j]), "+l"(a1r[j]), "+l"(a1i[j]));
        emit_and_roll(buf);
#pragma unroll
        for (int j = 0; j < K3_PP / 2; j++) asm volatile("" : "+l"(a0r[j]), "+l"(a0i[j]), "+l"(a1r[j]), "+l"(a1i[j]));
    }
    // drain: the 27 younger branches are the range-tail partial sums of outputs jb1-27 .. jb1-1
    int buf = (jb1 - jb0) & 1;
    __syncthreads();
    flush(jb1 - 1, buf ^ 1);
    for (int d = 1; d < K3_PP; d++) {
        emit_and_roll(buf);
        __syncthreads();
        flush(jb1 - 1 + d, buf);
        buf ^= 1;
    }
#undef K3_ROLL
}


int launch_fir_decimate(const K3Params& p, dim3 grid, size_t smem, cudaStream_t st)
{
    OWRX_CUDA(cudaFuncSetAttribute(fir_decimate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fir_decimate_kernel<<<grid, K3_NW * 32, smem, st>>>(p);
    OWRX_LAUNCH_CHECK();
    return OWRX_OK;
}

}  // namespace owrx
