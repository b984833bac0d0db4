// adpcm.cuh — the IMA-ADPCM quantiser step shared by FftAdpcm (waterfall.cu) and the client audio tail (selector_kernels.cuh).
//
// The codec is strictly sample-serial per stream (predictor and step index feed back): one warp, one stream per lane, and the
// speed is what ONE warp can issue per sample.  Measured on B200 (profiles/r2_adpcm_notes.md): the compare/subtract cascade with
// the successor steps selected from a 16-byte candidate entry was 43 instructions and 114 cycles per sample; a variant with
// seven independent threshold compares and IADD3 trees (8 dependent levels instead of 16) needed 50 instructions and ran
// SLOWER (142 cycles) — a lone warp is bound by ALU-pipe issue (~2.8 cycles per instruction), not by the dependency chain,
// so neither a shorter chain bought with more instructions nor fewer instructions bought with a load on the chain wins.
// Values of the step table are pinned by the browser decoder, reference htdocs/lib/AudioEngine.js:426-438.
#pragma once

#include "common.cuh"

namespace owrx {

__constant__ int16_t c_ima_step[89] = {
    7, 8, 9, 10, 11, 12, 13, 14, 16, 17, 19, 21, 23, 25, 28, 31, 34, 37, 41, 45,
    50, 55, 60, 66, 73, 80, 88, 97, 107, 118, 130, 143, 157, 173, 190, 209, 230, 253, 279, 307,
    337, 371, 408, 449, 494, 544, 598, 658, 724, 796, 876, 963, 1060, 1166, 1282, 1411, 1552, 1707, 1878, 2066,
    2272, 2499, 2749, 3024, 3327, 3660, 4026, 4428, 4871, 5358, 5894, 6484, 7132, 7845, 8630, 9493, 10442, 11487, 12635, 13899,
    15289, 16818, 18500, 20350, 22385, 24623, 27086, 29794, 32767};

struct ImaState {
    int pred, index, st;                           // st = step of the current index, carried in a register
};

// cand[i] = { step[max(i-1,0)] | step[min(i+2,88)] << 16,  step[min(i+4,88)] | step[min(i+6,88)] << 16,  step[min(i+8,88)], - }:
// the steps of the five possible successor indices in ONE 16-byte shared-memory entry per index, fetched while the
// quantiser's compare chain resolves, so no table lookup sits on the sample-to-sample dependency chain.
// (A single word succ[index * 8 + code] = next index << 16 | next step saves ten instructions but puts a shared-memory load
// between the code and the next sample's first compare: 126 instead of 114 cycles per sample, measured.)
// Fills `cand` (IMA_TABLE_ENTRIES entries of shared memory) cooperatively; the caller synchronises afterwards.
constexpr int IMA_TABLE_ENTRIES = 89;
__device__ __forceinline__ void ima_build_table(uint4* cand, int tid, int nthreads)
{
    for (int i = tid; i < 89; i += nthreads) {
        const unsigned c0 = c_ima_step[max(i - 1, 0)], c1 = c_ima_step[min(i + 2, 88)], c2 = c_ima_step[min(i + 4, 88)];
        const unsigned c3 = c_ima_step[min(i + 6, 88)], c4 = c_ima_step[min(i + 8, 88)];
        cand[i] = make_uint4(c0 | (c1 << 16), c2 | (c3 << 16), c4, 0u);
    }
}

__device__ __forceinline__ ImaState ima_state(int index, int pred)
{
    ImaState s;
    s.pred = pred; s.index = index; s.st = c_ima_step[index];
    return s;
}

__device__ __forceinline__ int ima_index(const ImaState& s) { return s.index; }

// one sample -> one 4-bit code
__device__ __forceinline__ unsigned ima_encode(int sample, ImaState& s, const uint4* cand)
{
    const uint4 c = cand[s.index];
    const int st = s.st;
    int diff = sample - s.pred;
    const int neg = diff < 0;
    diff = abs(diff);
    int code = 0;
    int d = st >> 3;
    if (diff >= st) { code = 4; diff -= st; d += st; }
    const int s1 = st >> 1;
    if (diff >= s1) { code |= 2; diff -= s1; d += s1; }
    const int s2 = st >> 2;
    if (diff >= s2) { code |= 1; d += s2; }
    s.pred = max(-32768, min(32767, neg ? s.pred - d : s.pred + d));
    // successor: index-1 for code < 4, else index + 2*(code-3); the matching step is field f of the entry
    const int f = max(code - 3, 0);
    s.index = max(0, min(88, s.index + (code < 4 ? -1 : 2 * code - 6)));
    const unsigned w = f < 2 ? c.x : (f < 4 ? c.y : c.z);
    s.st = (int)((w >> ((f & 1) << 4)) & 0xffffu);
    return (unsigned)(code | (neg << 3));
}

}  // namespace owrx
