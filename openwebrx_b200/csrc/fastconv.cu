// fastconv.cu — K3F kernels: polyphase fast-convolution channeliser (see fastconv.cuh for the algebra).
// Replaces pycsdr Shift + FirDecimate (reference call sites csdr/chain/selector.py:29,57,95,140) for all client
// channels of a decimator group at once.  Hand-written for sm_100a; no cuFFT / cuBLAS.
#include "fastconv.cuh"
#include "fft_small.cuh"

namespace owrx {

namespace {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ------------------------------------------------------------------------------------------------
// Tab[q][r][slot] = sum_{s<P} h[D s + r] e^{j 2 pi rate (D s + r)} e^{+j 2 pi q s / M}
// Phases in double (two sincospi per entry, then a 27-step double recurrence), rounded once to float.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fc_table_kernel(const float* __restrict__ h, int T, int D, int Dp, int P, int slots, const int* __restrict__ slot_list,
                const double* __restrict__ rate_list, float2* __restrict__ tab)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)FC_M * D) return;
    const int q = (int)(idx / D), r = (int)(idx % D);
    const double rate = rate_list[blockIdx.y];
    const int slot = slot_list[blockIdx.y];
    double a0 = rate * (double)r;
    a0 -= floor(a0);
    double aw = rate * (double)D;
    aw -= floor(aw);
    aw += (double)q / FC_M;
    double c0, s0, cw, sw;
    sincospi(2.0 * a0, &s0, &c0);
    sincospi(2.0 * aw, &sw, &cw);
    double ar = 0.0, ai = 0.0;
    for (int s = 0; s < P; s++) {
        const int t = D * s + r;
        if (t < T) {
            const double hv = (double)__ldg(h + t);
            ar += hv * c0;
            ai += hv * s0;
        }
        const double nc = c0 * cw - s0 * sw, ns = c0 * sw + s0 * cw;
        c0 = nc; s0 = ns;
    }
    tab[((size_t)q * Dp + r) * slots + slot] = make_float2((float)ar, (float)ai);
}

// ------------------------------------------------------------------------------------------------
// 256-point FFT of 32 interleaved sequences per CTA (16 threads x 16 points each per sequence, radix 16 x 16).
// Thread (g = tid >> 5, j = tid & 31): sequence j, pass-1 residue n2 = g, pass-2 output residue k1 = g.
// A warp = one g, 32 adjacent sequences -> every global access of a warp is one 256-byte row.
// ------------------------------------------------------------------------------------------------
constexpr int FC_SEQ = 32;
constexpr int FC_STR = FC_M + 1;            // odd sequence stride (in float2): conflict-free across the 32 lanes

__device__ __forceinline__ void fc_fill_twiddles(float2* tw)
{
    // tw[r * 16 + k] = e^{-2 pi i r k / 256}
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int e = ((i >> 4) * (i & 15)) & 255;
        float s, c;
        sincospif(-(float)e / 128.0f, &s, &c);
        tw[i] = make_float2(c, s);
    }
}

// in: v[n1] = x[16 n1 + g];  out: v[q] = X[g + 16 slot<16>(q)]
__device__ __forceinline__ void fc_fft256(float2* v, float2* seq, const float2* tw, int g)
{
    dft<16>(v);
#pragma unroll
    for (int q = 0; q < 16; q++) seq[g * 16 + slot<16>(q)] = v[q];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = seq[r * 16 + g];
#pragma unroll
    for (int r = 1; r < 16; r++) v[r] = cmul(v[r], tw[r * 16 + g]);
    dft<16>(v);
}

__global__ void __launch_bounds__(16 * FC_SEQ)
fc_forward_kernel(const float2* __restrict__ iq, long long n_lim, int D, int Dp, int Kb, int B, float2* __restrict__ F)
{
    extern __shared__ float2 fc_smem[];
    float2* tw = fc_smem;                       // [256]
    float2* seqs = fc_smem + 256;               // [32][257]
    const int tid = threadIdx.x, j = tid & 31, g = tid >> 5;
    const int b = blockIdx.y;
    const int r = blockIdx.x * FC_SEQ + j;
    fc_fill_twiddles(tw);
    float2 v[16];
    const long long s0 = (long long)b * Kb * D + r;
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) {
        const long long s = s0 + (long long)(16 * n1 + g) * D;
        v[n1] = (r < D && s < n_lim) ? __ldg(iq + s) : make_float2(0.f, 0.f);
    }
    __syncthreads();                            // twiddles visible
    fc_fft256(v, seqs + j * FC_STR, tw, g);
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int bin = g + 16 * slot<16>(q);
        F[((size_t)bin * B + b) * Dp + r] = v[q];
    }
}

// ------------------------------------------------------------------------------------------------
// Z[q][b][c] = sum_{r<Dp} F[q][b][r] * Tab[q][r][c]     (complex x complex, FP32 FMA pipe)
// CTA = (bin q, tile of BT = 8 NW blocks, 64 channel slots).  Warp w owns rows 8w..8w+7, lane l columns 2l, 2l+1:
// per pair of branches a thread issues 8 broadcast LDS.128 (two F values of each of its rows) + 2 LDS.128 (its two
// table columns of both branches) for 128 FMAs.  Operand chunks of 32 branches stream through a 3-stage cp.async
// pipeline.
// ------------------------------------------------------------------------------------------------
constexpr int FC_ST = 3;

template <int NW>
__global__ void __launch_bounds__(NW * 32)
fc_contract_kernel(const float2* __restrict__ F, const float2* __restrict__ tab, float2* __restrict__ Z, int B, int Dp, int slots, int nbt)
{
    constexpr int BT = 8 * NW;
    extern __shared__ float4 fc_smem4[];
    float2* Fs = reinterpret_cast<float2*>(fc_smem4);                 // [ST][BT][KC]
    float2* Ts = Fs + FC_ST * BT * FC_KC;                             // [ST][KC][64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x / nbt, bt = blockIdx.x % nbt;
    const int cg = blockIdx.y;
    const int b0 = bt * BT;
    const int rows = min(BT, B - b0);
    const int nchunks = Dp / FC_KC;

    const float2* Fq = F + ((size_t)q * B + b0) * Dp;
    const float2* Tq = tab + (size_t)q * Dp * slots + (size_t)cg * FC_CG;

    auto load_chunk = [&](int chunk, int stage) {
        float2* fs = Fs + stage * BT * FC_KC;
        float2* ts = Ts + stage * FC_KC * FC_CG;
        // F: BT rows x 16 pieces of 16 B
        for (int i = tid; i < BT * (FC_KC / 2); i += NW * 32) {
            const int row = i / (FC_KC / 2), pc = i % (FC_KC / 2);
            const int srow = min(row, rows - 1);
            cp_async16(fs + row * FC_KC + pc * 2, Fq + (size_t)srow * Dp + chunk * FC_KC + pc * 2);
        }
        // Tab: KC rows x 32 pieces
        for (int i = tid; i < FC_KC * (FC_CG / 2); i += NW * 32) {
            const int kr = i / (FC_CG / 2), pc = i % (FC_CG / 2);
            cp_async16(ts + kr * FC_CG + pc * 2, Tq + (size_t)(chunk * FC_KC + kr) * slots + pc * 2);
        }
    };

    float2 acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i][0] = make_float2(0.f, 0.f); acc[i][1] = make_float2(0.f, 0.f); }

#pragma unroll
    for (int s = 0; s < FC_ST - 1; s++) {
        if (s < nchunks) load_chunk(s, s);
        cp_commit();
    }
    for (int ch = 0; ch < nchunks; ch++) {
        cp_wait<FC_ST - 2>();
        __syncthreads();
        {
            const int nx = ch + FC_ST - 1;
            if (nx < nchunks) load_chunk(nx, nx % FC_ST);
            cp_commit();
        }
        const int stage = ch % FC_ST;
        const float4* fs = reinterpret_cast<const float4*>(Fs + stage * BT * FC_KC + warp * 8 * FC_KC);
        const float4* ts = reinterpret_cast<const float4*>(Ts + stage * FC_KC * FC_CG) + lane;
#pragma unroll 4
        for (int kk = 0; kk < FC_KC / 2; kk++) {
            const float4 t0 = ts[(2 * kk) * (FC_CG / 2)];
            const float4 t1 = ts[(2 * kk + 1) * (FC_CG / 2)];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float4 f = fs[i * (FC_KC / 2) + kk];
                acc[i][0].x = fmaf(f.x, t0.x, acc[i][0].x); acc[i][0].x = fmaf(-f.y, t0.y, acc[i][0].x);
                acc[i][0].y = fmaf(f.x, t0.y, acc[i][0].y); acc[i][0].y = fmaf(f.y, t0.x, acc[i][0].y);
                acc[i][1].x = fmaf(f.x, t0.z, acc[i][1].x); acc[i][1].x = fmaf(-f.y, t0.w, acc[i][1].x);
                acc[i][1].y = fmaf(f.x, t0.w, acc[i][1].y); acc[i][1].y = fmaf(f.y, t0.z, acc[i][1].y);
                acc[i][0].x = fmaf(f.z, t1.x, acc[i][0].x); acc[i][0].x = fmaf(-f.w, t1.y, acc[i][0].x);
                acc[i][0].y = fmaf(f.z, t1.y, acc[i][0].y); acc[i][0].y = fmaf(f.w, t1.x, acc[i][0].y);
                acc[i][1].x = fmaf(f.z, t1.z, acc[i][1].x); acc[i][1].x = fmaf(-f.w, t1.w, acc[i][1].x);
                acc[i][1].y = fmaf(f.z, t1.w, acc[i][1].y); acc[i][1].y = fmaf(f.w, t1.z, acc[i][1].y);
            }
        }
    }
    float4* Zq = reinterpret_cast<float4*>(Z + ((size_t)q * B + b0) * slots + (size_t)cg * FC_CG) + lane;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int row = warp * 8 + i;
        if (row < rows) Zq[(size_t)row * (slots / 2)] = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
    }
}

// ------------------------------------------------------------------------------------------------
// z[m] = (1/M) IFFT_M over q of Z[q][b][c]; valid m < Kb; post-rotation e^{j 2 pi (ph + rate (kD + 1))}; store s1.
// IFFT(x) = conj(FFT(conj(x))).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(16 * FC_SEQ)
fc_inverse_kernel(const float2* __restrict__ Z, int B, int slots, int D, int Kb, const double* __restrict__ ch_rate,
                  const double* __restrict__ ch_phase, long long k0, long long n_k, float2* __restrict__ out)
{
    extern __shared__ float2 fc_smem[];
    float2* tw = fc_smem;
    float2* seqs = fc_smem + 256;
    const int tid = threadIdx.x, j = tid & 31, g = tid >> 5;
    const int b = blockIdx.y;
    const int c = blockIdx.x * FC_SEQ + j;
    fc_fill_twiddles(tw);
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) {
        const float2 z = __ldg(Z + ((size_t)(16 * n1 + g) * B + b) * slots + c);
        v[n1] = make_float2(z.x, -z.y);
    }
    __syncthreads();
    fc_fft256(v, seqs + j * FC_STR, tw, g);
    const double rate = ch_rate[c], ph = ch_phase[c];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const int m = g + 16 * slot<16>(q);
        const long long k = k0 + (long long)b * Kb + m;
        if (m < Kb && k < n_k) {
            double t = ph + rate * (double)(k * D + 1);
            t -= floor(t);
            float sn, cs;
            sincospif(2.0f * (float)t, &sn, &cs);
            const float2 z = make_float2(v[q].x * (1.0f / FC_M), -v[q].y * (1.0f / FC_M));
            out[(size_t)k * slots + c] = cmul(z, make_float2(cs, sn));
        }
    }
}

constexpr size_t kFftSmem = (256 + FC_SEQ * FC_STR) * sizeof(float2);

template <int NW>
int launch_contract_nw(const FcShape& sh, const float2* F, const float2* tab, int B, float2* Z, cudaStream_t st)
{
    constexpr int BT = 8 * NW;
    const size_t smem = (size_t)FC_ST * (BT * FC_KC + FC_KC * FC_CG) * sizeof(float2);
    OWRX_CUDA(cudaFuncSetAttribute(fc_contract_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nbt = (B + BT - 1) / BT;
    fc_contract_kernel<NW><<<dim3((unsigned)(FC_M * nbt), (unsigned)(sh.slots / FC_CG)), NW * 32, smem, st>>>(F, tab, Z, B, sh.Dp, sh.slots, nbt);
    OWRX_LAUNCH_CHECK();
    return OWRX_OK;
}

}  // namespace

int fc_launch_table(const FcShape& sh, const float* d_h, const int* d_slot_list, const double* d_rate_list, int n, float2* d_tab,
                    cudaStream_t st)
{
    if (n <= 0) return OWRX_OK;
    const long long total = (long long)FC_M * sh.D;
    fc_table_kernel<<<dim3((unsigned)((total + 255) / 256), (unsigned)n), 256, 0, st>>>(d_h, sh.T, sh.D, sh.Dp, sh.P, sh.slots, d_slot_list,
                                                                                       d_rate_list, d_tab);
    OWRX_LAUNCH_CHECK();
    return OWRX_OK;
}

int fc_launch_forward(const FcShape& sh, const float2* iq, long long n_lim, int B, float2* d_F, cudaStream_t st)
{
    OWRX_CUDA(cudaFuncSetAttribute(fc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFftSmem));
    fc_forward_kernel<<<dim3((unsigned)(sh.Dp / FC_SEQ), (unsigned)B), 16 * FC_SEQ, kFftSmem, st>>>(iq, n_lim, sh.D, sh.Dp, sh.Kb, B, d_F);
    OWRX_LAUNCH_CHECK();
    return OWRX_OK;
}

int fc_launch_contract(const FcShape& sh, const float2* d_F, const float2* d_tab, int B, float2* d_Z, int sm_count, cudaStream_t st)
{
    (void)sm_count;
    static const int force_nw = getenv("OWRX_FC_NW") ? atoi(getenv("OWRX_FC_NW")) : 0;
    int nw = force_nw;
    if (!nw) {
        // smallest tile that covers B in as few tiles as the largest one does
        const int cands[] = {1, 2, 4, 6, 8, 11};
        const int best_tiles = (B + 87) / 88;
        nw = 11;
        for (int c : cands) if ((B + 8 * c - 1) / (8 * c) == best_tiles) { nw = c; break; }
    }
    switch (nw) {
    case 1: return launch_contract_nw<1>(sh, d_F, d_tab, B, d_Z, st);
    case 2: return launch_contract_nw<2>(sh, d_F, d_tab, B, d_Z, st);
    case 4: return launch_contract_nw<4>(sh, d_F, d_tab, B, d_Z, st);
    case 6: return launch_contract_nw<6>(sh, d_F, d_tab, B, d_Z, st);
    case 8: return launch_contract_nw<8>(sh, d_F, d_tab, B, d_Z, st);
    default: return launch_contract_nw<11>(sh, d_F, d_tab, B, d_Z, st);
    }
}

int fc_launch_inverse(const FcShape& sh, const float2* d_Z, int B, const double* d_rate, const double* d_phase, long long k0, long long n_k,
                      float2* out, cudaStream_t st)
{
    OWRX_CUDA(cudaFuncSetAttribute(fc_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFftSmem));
    fc_inverse_kernel<<<dim3((unsigned)(sh.slots / FC_SEQ), (unsigned)B), 16 * FC_SEQ, kFftSmem, st>>>(d_Z, B, sh.slots, sh.D, sh.Kb, d_rate,
                                                                                                    d_phase, k0, n_k, out);
    OWRX_LAUNCH_CHECK();
    return OWRX_OK;
}

}  // namespace owrx
