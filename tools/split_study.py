#!/usr/bin/env python
"""How many operand bits does the tensor-core contraction of K3F need?  numpy only (no GPU): one overlap-save block of the
C2 shape (10 MS/s -> 12 kHz, D = 833, T = 22223, 256-point branch FFTs, the 64-carrier plan of the parity tests), the
contraction Z[q] = sum_r F[q][r] Tab[q][r] evaluated with the operands split into 16-bit terms, accumulated in float64 so that
only the operand rounding shows, against the float32-operand result; relative RMS of the block's valid outputs on the three
weakest channels (66 dB below full scale) and the strongest.

  bf16 x 1          hh                                   what a plain bf16 GEMM would give
  bf16 x 2 (3)      hh + hm + mh
  bf16 x 3 (6)      + hl + lh + mm                       round 1's form (fastconv_tc.cu, tc_levels = 3)
  fp16 x 2 (3)      hh + hm + mh on block-scaled fp16    round 2's form for D >= 2048 (tc_levels = 2)
  f32 accumulate    float32 operands, float32 accumulation (numpy pairwise): the noise floor any FP32 evaluation has

Result (profiles/r2_operand_split_study.md): 16 operand bits leave 4e-5 of the channel level — inside the 1e-4 budget with a factor
2.5 to spare, not enough; 22 bits (fp16 x 2) leave 6-9e-7, the level of float32 accumulation noise itself (0.6-1e-6); 24 bits
(bf16 x 3) 6e-8.  On the GPU the two-term form measures BETTER than the three-term form against float64: it
issues half as many FP32 accumulations in TMEM."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwebrx_b200.synth import carrier_plan, make_iq                   # noqa: E402


def bf16(x):
    x = np.asarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16                      # round to nearest even on the top 16 bits
    return r.astype(np.uint32).view(np.float32)


def split(x, levels, rnd):
    out, res = [], np.asarray(x, np.float32).copy()
    for _ in range(levels):
        h = rnd(res).astype(np.float32)
        out.append(h.astype(np.float64))
        res = (res - h).astype(np.float32)
    return out


def pow2_scale(bound, target_log2):
    _, e = np.frexp(bound)                                                # bound < 2^e
    return float(np.ldexp(1.0, target_log2 - int(e)))


def main():
    fs, D, M, T = 10e6, 833, 256, 22223
    P = -(-T // D)
    Kb = M - P + 1
    n = np.arange(T) - (T - 1) / 2
    fc = 0.5 / D
    h = 2 * fc * np.sinc(2 * fc * n) * np.hamming(T)
    h = (h / h.sum()).astype(np.float32).astype(np.float64)
    cars = carrier_plan(64, fs, seed=3)
    x = make_iq(M * D + D, fs, cars)
    F32 = np.fft.fft(x[:M * D].reshape(M, D).astype(np.complex128), axis=0).astype(np.complex64)      # F[q][r]
    hp = np.zeros(P * D)
    hp[:T] = h
    order = np.argsort([c["amp"] for c in cars])
    sF = pow2_scale(M * np.sqrt(2.0) * max(np.abs(x.real).max(), np.abs(x.imag).max()), 15)
    sT = pow2_scale(np.abs(hp.reshape(P, D)).sum(0).max(), 14)
    f16 = lambda v: np.asarray(v, np.float32).astype(np.float16)                                    # noqa: E731
    print("| channel | level | bf16 x 1 | bf16 x 2 (3 products) | bf16 x 3 (6) | fp16 x 2 (3), block-scaled | f32 accumulate |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for ci in list(order[:3]) + [order[-1]]:
        rate = -cars[ci]["offset"] / fs
        g = (hp * np.exp(2j * np.pi * rate * np.arange(P * D))).reshape(P, D)
        s, q = np.arange(P)[None, :, None], np.arange(M)[:, None, None]
        T32 = (g[None] * np.exp(2j * np.pi * q * s / M)).sum(1).astype(np.complex64)                 # Tab[q][r]

        def contract(levels, maxw, rnd, kf=1.0, kt=1.0):
            Fr, Fi = split(F32.real * np.float32(kf), levels, rnd), split(F32.imag * np.float32(kf), levels, rnd)
            Tr, Ti = split(T32.real * np.float32(kt), levels, rnd), split(T32.imag * np.float32(kt), levels, rnd)
            Z = np.zeros(M, np.complex128)
            for i in range(levels):
                for j in range(levels):
                    if i + j <= maxw:
                        Z += ((Fr[i] * Tr[j] - Fi[i] * Ti[j]) + 1j * (Fr[i] * Ti[j] + Fi[i] * Tr[j])).sum(1)
            return Z / (kf * kt)

        ref = np.fft.ifft((F32.astype(np.complex128) * T32.astype(np.complex128)).sum(1))[:Kb]
        lvl = np.sqrt(np.mean(np.abs(ref) ** 2))
        err = lambda Z: np.sqrt(np.mean(np.abs(np.fft.ifft(Z)[:Kb] - ref) ** 2)) / lvl               # noqa: E731
        row = [err(contract(1, 0, bf16)), err(contract(2, 1, bf16)), err(contract(3, 2, bf16)), err(contract(2, 1, f16, sF, sT)),
               err((F32 * T32).sum(1, dtype=np.complex64).astype(np.complex128))]
        print("| %d (%.1f dBFS) | %.2e | %s |" % (ci, 20 * np.log10(cars[ci]["amp"]), lvl, " | ".join("%.1e" % v for v in row)))


if __name__ == "__main__":
    main()
