/*
 * csdr_oracle.c — CPU restatement of the csdr arithmetic on OpenWebRX+'s DSP hot path.
 * TEST INFRASTRUCTURE ONLY (see csdr_oracle.h).  PARITY UNPINNED against a pycsdr binary; every
 * function cites the reference call site it serves and the SURVEY.md Appendix-A clause it follows.
 *
 * Build: make -C oracle   (gcc -O3 -march=x86-64-v3; no external libraries)
 */
#include "csdr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------
 * Filter and window design — SURVEY A.1.  Used by FirDecimate (csdr/chain/selector.py:29,57),
 * Bandpass (csdr/chain/selector.py:115-117,165-166) and the FractionalDecimator prefilter
 * (csdr/chain/analog.py:66,91).  Designed in double, rounded once to float32.
 * ---------------------------------------------------------------------------------------------- */
int oc_filter_len(double transition)
{
    int len = (int)(4.0 / transition);
    if ((len & 1) == 0) len += 1;
    return len;
}

static double design_window(double r) /* r in [-1,1]; Hamming */
{
    return 0.54 - 0.46 * cos(2.0 * M_PI * (0.5 + r / 2.0));
}

static void lowpass_double(double* h, int len, double fc)
{
    int middle = len / 2;
    double sum;
    h[middle] = 2.0 * M_PI * fc * design_window(0.0);
    for (int i = 1; i <= middle; i++) {
        double v = sin(2.0 * M_PI * fc * i) / i * design_window((double)i / middle);
        h[middle + i] = v;
        h[middle - i] = v;
    }
    sum = 0.0;
    for (int i = 0; i < len; i++) sum += h[i];
    for (int i = 0; i < len; i++) h[i] /= sum;
}

void oc_firdes_lowpass(float* taps, int len, double cutoff_rate)
{
    double* h = (double*)malloc(sizeof(double) * (size_t)len);
    lowpass_double(h, len, cutoff_rate);
    for (int i = 0; i < len; i++) taps[i] = (float)h[i];
    free(h);
}

void oc_firdes_bandpass(oc_cf32* taps, int len, double lo, double hi)
{
    double* h = (double*)malloc(sizeof(double) * (size_t)len);
    double fc = (hi - lo) / 2.0, centre = (hi + lo) / 2.0;
    lowpass_double(h, len, fc);
    for (int i = 0; i < len; i++) {
        double ph = 2.0 * M_PI * centre * i;
        taps[i].re = (float)(h[i] * cos(ph));
        taps[i].im = (float)(h[i] * sin(ph));
    }
    free(h);
}

/* Fft default window — SURVEY A.2; Fft(size=, every_n_samples=) at csdr/chain/fft.py:34 */
void oc_fft_window_hamming(float* w, int n)
{
    for (int i = 0; i < n; i++)
        w[i] = (float)(0.54 - 0.46 * cos(2.0 * M_PI * i / (double)(n - 1)));
}

/* NfmDeemphasis(sampleRate) — csdr/chain/analog.py:43,52.  SPEC-DEFINED (upstream tap tables are
 * not recoverable): linear-phase FIR by frequency sampling of A(f) = 1 (f<=400 Hz), 400/f
 * (400<f<=4000), 0 above; Hamming window; gain normalised to 1 at 400 Hz. */
int oc_nfm_deemphasis_len(int sample_rate) { return sample_rate >= 24000 ? 199 : 79; }

void oc_nfm_deemphasis_taps(float* taps, int len, int sample_rate)
{
    const int M = 8192;
    int mid = len / 2;
    double fs = (double)sample_rate;
    double* h = (double*)malloc(sizeof(double) * (size_t)len);
    for (int n = 0; n < len; n++) {
        double acc = 0.0;
        for (int k = 0; k <= M / 2; k++) {
            double f = k * fs / M, a;
            if (f <= 400.0) a = 1.0; else if (f <= 4000.0) a = 400.0 / f; else a = 0.0;
            double c = (k == 0 || k == M / 2) ? 0.5 : 1.0;
            acc += c * a * cos(2.0 * M_PI * k * (double)(n - mid) / M);
        }
        h[n] = acc * 2.0 / M * (0.54 - 0.46 * cos(2.0 * M_PI * n / (double)(len - 1)));
    }
    double g = 0.0;
    for (int n = 0; n < len; n++) g += h[n] * cos(2.0 * M_PI * 400.0 / fs * (n - mid));
    for (int n = 0; n < len; n++) taps[n] = (float)(h[n] / g);
    free(h);
}

/* ------------------------------------------------------------------------------------------------
 * FFT — SURVEY A.2.  Unnormalised forward DFT, float32, iterative radix-2 DIT with a
 * thread-local plan cache (FFTW3f in upstream; an FFTW build would be ~1.5-2x faster).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int n; float* tw_re; float* tw_im; uint32_t* rev; } fft_plan;
static __thread fft_plan g_plans[8];
static __thread int g_nplans = 0;

static fft_plan* get_plan(int n)
{
    for (int i = 0; i < g_nplans; i++) if (g_plans[i].n == n) return &g_plans[i];
    fft_plan* p = &g_plans[g_nplans < 8 ? g_nplans++ : 7];
    p->n = n;
    p->tw_re = (float*)malloc(sizeof(float) * (size_t)(n / 2 + 1));
    p->tw_im = (float*)malloc(sizeof(float) * (size_t)(n / 2 + 1));
    p->rev = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    for (int k = 0; k < n / 2; k++) {
        double a = -2.0 * M_PI * k / n;
        p->tw_re[k] = (float)cos(a);
        p->tw_im[k] = (float)sin(a);
    }
    int bits = 0;
    while ((1 << bits) < n) bits++;
    for (int i = 0; i < n; i++) {
        uint32_t r = 0;
        for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1u << (bits - 1 - b);
        p->rev[i] = r;
    }
    return p;
}

void oc_fft_forward(const oc_cf32* in, oc_cf32* out, int n)
{
    fft_plan* p = get_plan(n);
    for (int i = 0; i < n; i++) out[p->rev[i]] = in[i];
    for (int half = 1; half < n; half <<= 1) {
        int step = n / (2 * half);
        for (int base = 0; base < n; base += 2 * half) {
            for (int k = 0; k < half; k++) {
                float wr = p->tw_re[k * step], wi = p->tw_im[k * step];
                oc_cf32 a = out[base + k], b = out[base + k + half];
                float tr = b.re * wr - b.im * wi;
                float ti = b.re * wi + b.im * wr;
                out[base + k].re = a.re + tr;        out[base + k].im = a.im + ti;
                out[base + k + half].re = a.re - tr; out[base + k + half].im = a.im - ti;
            }
        }
    }
}

void oc_fft_frame(const oc_cf32* x, const float* window, oc_cf32* out, int n)
{
    oc_cf32* tmp = (oc_cf32*)malloc(sizeof(oc_cf32) * (size_t)n);
    for (int i = 0; i < n; i++) { tmp[i].re = x[i].re * window[i]; tmp[i].im = x[i].im * window[i]; }
    oc_fft_forward(tmp, out, n);
    free(tmp);
}

/* LogPower(add_db=-70) — csdr/chain/fft.py:20; SURVEY A.3 */
void oc_log_power(const oc_cf32* X, float* out, int n, float add_db)
{
    for (int i = 0; i < n; i++)
        out[i] = 10.0f * log10f(X[i].re * X[i].re + X[i].im * X[i].im) + add_db;
}

/* LogAveragePower(add_db=-70, fft_size=, avg_number=) — csdr/chain/fft.py:22; SURVEY A.3 */
void oc_log_average_power(const oc_cf32* frames, int avg, float* out, int n, float add_db)
{
    float corr = add_db - 10.0f * log10f((float)avg);
    for (int i = 0; i < n; i++) {
        float s = 0.0f;
        for (int j = 0; j < avg; j++) {
            const oc_cf32 v = frames[(size_t)j * n + i];
            s += v.re * v.re + v.im * v.im;
        }
        out[i] = 10.0f * log10f(s) + corr;
    }
}

/* FftSwap(fft_size=) — csdr/chain/fft.py:36; SURVEY A.4 */
void oc_fft_swap(const float* in, float* out, int n)
{
    int h = n / 2;
    memcpy(out, in + h, sizeof(float) * (size_t)h);
    memcpy(out + h, in, sizeof(float) * (size_t)h);
}

/* IMA-ADPCM tables — identical to htdocs/lib/AudioEngine.js:426-438 */
static const int8_t IMA_INDEX[16] = { -1, -1, -1, -1, 2, 4, 6, 8, -1, -1, -1, -1, 2, 4, 6, 8 };
static const int16_t IMA_STEP[89] = {
    7, 8, 9, 10, 11, 12, 13, 14, 16, 17, 19, 21, 23, 25, 28, 31, 34, 37, 41, 45,
    50, 55, 60, 66, 73, 80, 88, 97, 107, 118, 130, 143, 157, 173, 190, 209, 230, 253, 279, 307,
    337, 371, 408, 449, 494, 544, 598, 658, 724, 796, 876, 963, 1060, 1166, 1282, 1411, 1552, 1707, 1878, 2066,
    2272, 2499, 2749, 3024, 3327, 3660, 4026, 4428, 4871, 5358, 5894, 6484, 7132, 7845, 8630, 9493, 10442, 11487, 12635, 13899,
    15289, 16818, 18500, 20350, 22385, 24623, 27086, 29794, 32767 };

static inline uint8_t ima_encode_sample(int16_t sample, int* index, int* prev)
{
    int diff = (int)sample - *prev;
    int step = IMA_STEP[*index];
    int code = 0;
    if (diff < 0) { code = 8; diff = -diff; }
    if (diff >= step) { code |= 4; diff -= step; }
    step >>= 1;
    if (diff >= step) { code |= 2; diff -= step; }
    step >>= 1;
    if (diff >= step) { code |= 1; }
    /* decoder-mirrored state update */
    int st = IMA_STEP[*index];
    int d = st >> 3;
    if (code & 1) d += st >> 2;
    if (code & 2) d += st >> 1;
    if (code & 4) d += st;
    if (code & 8) d = -d;
    int p = *prev + d;
    if (p > 32767) p = 32767; else if (p < -32768) p = -32768;
    *prev = p;
    int ix = *index + IMA_INDEX[code];
    if (ix < 0) ix = 0; else if (ix > 88) ix = 88;
    *index = ix;
    return (uint8_t)code;
}

/* low nibble first (AudioEngine.js:440-447) */
void oc_ima_adpcm_encode(const int16_t* s, int n, uint8_t* out, int* index, int* predictor)
{
    for (int i = 0; i + 1 < n; i += 2) {
        uint8_t lo = ima_encode_sample(s[i], index, predictor);
        uint8_t hi = ima_encode_sample(s[i + 1], index, predictor);
        out[i / 2] = (uint8_t)(lo | (hi << 4));
    }
}

/* Standard IMA decoder (the mirror of the encoder's state update).  The browser's decodeNibble
 * (AudioEngine.js:493-509) differs only in using the step looked up after the previous nibble,
 * starting from step=0; tests/test_adpcm_js.py transliterates that decoder separately. */
void oc_ima_adpcm_decode(const uint8_t* in, int nbytes, int16_t* out, int* index, int* predictor)
{
    for (int i = 0; i < nbytes; i++) {
        for (int half = 0; half < 2; half++) {
            int code = half ? (in[i] >> 4) & 15 : in[i] & 15;
            int st = IMA_STEP[*index];
            int d = st >> 3;
            if (code & 1) d += st >> 2;
            if (code & 2) d += st >> 1;
            if (code & 4) d += st;
            if (code & 8) d = -d;
            int p = *predictor + d;
            if (p > 32767) p = 32767; else if (p < -32768) p = -32768;
            *predictor = p;
            int ix = *index + IMA_INDEX[code];
            if (ix < 0) ix = 0; else if (ix > 88) ix = 88;
            *index = ix;
            out[2 * i + half] = (int16_t)p;
        }
    }
}

/* FftAdpcm quantiser — csdr/chain/fft.py:44,89; pad count = COMPRESS_FFT_PAD_N
 * (htdocs/openwebrx.js:845), scale = 100 (htdocs/openwebrx.js:1128). */
void oc_fft_adpcm_quantise(const float* db, int16_t* s, int n)
{
    for (int i = 0; i < n; i++) {
        float v = db[i] * 100.0f;
        int16_t q;
        if (!(v > -32768.0f)) q = -32768;           /* -inf / NaN clamp (UB upstream) */
        else if (v > 32767.0f) q = 32767;
        else q = (int16_t)v;                         /* C cast: truncate toward zero */
        s[10 + i] = q;
    }
    for (int i = 0; i < 10; i++) s[i] = s[10];
}

void oc_fft_adpcm(const float* db, uint8_t* out, int n)
{
    int16_t* s = (int16_t*)malloc(sizeof(int16_t) * (size_t)(n + 10));
    int index = 0, pred = 0;
    oc_fft_adpcm_quantise(db, s, n);
    oc_ima_adpcm_encode(s, n + 10, out, &index, &pred);
    free(s);
}

/* FftChain parameter math — csdr/chain/fft.py:75-85 */
void oc_fftchain_params(double samp_rate, int fft_size, double voverlap, double fps, int* avg, int* every_n)
{
    int a = 0;
    if (voverlap > 0) a = (int)nearbyint(1.0 * samp_rate / fft_size / fps / (1.0 - voverlap));
    *avg = a;
    if (a == 0) *every_n = (int)(samp_rate / fps);
    else *every_n = (int)(samp_rate / fps / a);
}

/* Whole FftChain: Fft -> LogPower|LogAveragePower -> FftSwap -> [FftAdpcm]; csdr/chain/fft.py:25-49 */
/* Spectral-subtraction noise filter on the averaged power of one waterfall line (BASELINE config 4).  SPEC-DEFINED: the
 * reference has no waterfall noise filter (its NoiseFilter is audio-only, csdr/chain/clientaudio.py:13-14, algorithm
 * not in the tree) — this is the stage the survey prescribes, excluded from reference-parity claims.
 * Per bin, across lines: noise floor by minimum tracking with upward drift, subtracted in the linear domain with a
 * spectral floor:   N = first ? P : min(P, N * (1 + growth));   P' = max(P - alpha * N, beta * P). */
void oc_wf_noise_filter(float* pw, float* noise, int n, int first, float alpha, float beta, float growth)
{
    for (int i = 0; i < n; i++) {
        const float p = pw[i];
        const float nf = first ? p : fminf(p, noise[i] * (1.0f + growth));
        noise[i] = nf;
        pw[i] = fmaxf(p - alpha * nf, beta * p);
    }
}

static size_t fftchain_core(const oc_cf32* iq, size_t n_samples, int n, int every_n, int avg, float add_db,
                            int compression, uint8_t* out, size_t out_cap, int16_t* s16_out, float* db_out,
                            int nf_on, float nf_alpha, float nf_beta, float nf_growth);

size_t oc_fftchain_run(const oc_cf32* iq, size_t n_samples, int n, int every_n, int avg, float add_db,
                       int compression, uint8_t* out, size_t out_cap, int16_t* s16_out, float* db_out)
{
    return fftchain_core(iq, n_samples, n, every_n, avg, add_db, compression, out, out_cap, s16_out, db_out, 0, 0.f, 0.f, 0.f);
}

size_t oc_fftchain_run_nf(const oc_cf32* iq, size_t n_samples, int n, int every_n, int avg, float add_db,
                          int compression, uint8_t* out, size_t out_cap, int16_t* s16_out, float* db_out,
                          float nf_alpha, float nf_beta, float nf_growth)
{
    return fftchain_core(iq, n_samples, n, every_n, avg, add_db, compression, out, out_cap, s16_out, db_out, 1, nf_alpha, nf_beta,
                         nf_growth);
}

static size_t fftchain_core(const oc_cf32* iq, size_t n_samples, int n, int every_n, int avg, float add_db,
                            int compression, uint8_t* out, size_t out_cap, int16_t* s16_out, float* db_out,
                            int nf_on, float nf_alpha, float nf_beta, float nf_growth)
{
    float* noise = (float*)malloc(sizeof(float) * (size_t)n);
    size_t line_bytes = compression ? (size_t)(n + 10) / 2 : (size_t)n * 4;
    size_t frames_per_line = avg > 0 ? (size_t)avg : 1;
    size_t nframes = 0, nlines = 0;
    if (n_samples >= (size_t)n && every_n > 0) nframes = (n_samples - (size_t)n) / (size_t)every_n + 1;
    size_t total_lines = nframes / frames_per_line;
    float* window = (float*)malloc(sizeof(float) * (size_t)n);
    oc_cf32* X = (oc_cf32*)malloc(sizeof(oc_cf32) * (size_t)n);
    float* pw = (float*)malloc(sizeof(float) * (size_t)n);
    float* line = (float*)malloc(sizeof(float) * (size_t)n);
    float* swapped = (float*)malloc(sizeof(float) * (size_t)n);
    int16_t* s = (int16_t*)malloc(sizeof(int16_t) * (size_t)(n + 10));
    oc_fft_window_hamming(window, n);
    for (size_t l = 0; l < total_lines; l++) {
        if ((nlines + 1) * line_bytes > out_cap && out) break;
        if (avg > 0) {
            float corr = add_db - 10.0f * log10f((float)avg);
            memset(pw, 0, sizeof(float) * (size_t)n);
            for (int j = 0; j < avg; j++) {
                size_t f = l * (size_t)avg + (size_t)j;
                oc_fft_frame(iq + f * (size_t)every_n, window, X, n);
                for (int i = 0; i < n; i++) pw[i] += X[i].re * X[i].re + X[i].im * X[i].im;
            }
            if (nf_on) oc_wf_noise_filter(pw, noise, n, l == 0, nf_alpha, nf_beta, nf_growth);
            for (int i = 0; i < n; i++) line[i] = 10.0f * log10f(pw[i]) + corr;
        } else {
            oc_fft_frame(iq + l * (size_t)every_n, window, X, n);
            if (nf_on) {
                for (int i = 0; i < n; i++) pw[i] = X[i].re * X[i].re + X[i].im * X[i].im;
                oc_wf_noise_filter(pw, noise, n, l == 0, nf_alpha, nf_beta, nf_growth);
                for (int i = 0; i < n; i++) line[i] = 10.0f * log10f(pw[i]) + add_db;
            } else {
                oc_log_power(X, line, n, add_db);
            }
        }
        oc_fft_swap(line, swapped, n);
        if (db_out) memcpy(db_out + l * (size_t)n, swapped, sizeof(float) * (size_t)n);
        if (compression) {
            int index = 0, pred = 0;
            oc_fft_adpcm_quantise(swapped, s, n);
            if (s16_out) memcpy(s16_out + l * (size_t)(n + 10), s, sizeof(int16_t) * (size_t)(n + 10));
            if (out) oc_ima_adpcm_encode(s, n + 10, out + l * line_bytes, &index, &pred);
        } else if (out) {
            memcpy(out + l * line_bytes, swapped, line_bytes);
        }
        nlines++;
    }
    free(window); free(X); free(pw); free(line); free(swapped); free(s); free(noise);
    return nlines;
}

/* ------------------------------------------------------------------------------------------------
 * Selector stages
 * ---------------------------------------------------------------------------------------------- */

/* Shift(rate) — csdr/chain/selector.py:95,140; SURVEY A.6.  Phase is a pure function of the
 * absolute sample index, so results do not depend on how the stream is partitioned. */
void oc_shift(const oc_cf32* x, oc_cf32* y, size_t n, double rate, double phase0_turns, uint64_t n0, int fast)
{
    if (!fast) {
        for (size_t i = 0; i < n; i++) {
            double t = phase0_turns + rate * (double)(n0 + i + 1);
            t -= floor(t);
            double a = 2.0 * M_PI * t;
            float c = (float)cos(a), s = (float)sin(a);
            float re = x[i].re * c - x[i].im * s;
            float im = x[i].re * s + x[i].im * c;
            y[i].re = re; y[i].im = im;
        }
        return;
    }
    double wa = 2.0 * M_PI * (rate - floor(rate));
    float wc = (float)cos(wa), ws = (float)sin(wa);
    for (size_t b = 0; b < n; b += 256) {
        double t = phase0_turns + rate * (double)(n0 + b + 1);
        t -= floor(t);
        float c = (float)cos(2.0 * M_PI * t), s = (float)sin(2.0 * M_PI * t);
        size_t e = b + 256 < n ? b + 256 : n;
        for (size_t i = b; i < e; i++) {
            float re = x[i].re * c - x[i].im * s;
            float im = x[i].re * s + x[i].im * c;
            y[i].re = re; y[i].im = im;
            float nc = c * wc - s * ws, ns = c * ws + s * wc;
            c = nc; s = ns;
        }
    }
}

/* FirDecimate(decimation, transition, cutoff) — csdr/chain/selector.py:29,57; SURVEY A.7.
 * Accumulates in 16 interleaved float32 partial sums (even lanes re, odd lanes im). */
size_t oc_fir_decimate(const oc_cf32* x, size_t n, const float* taps, int T, int D, oc_cf32* y)
{
    if (n < (size_t)T) return 0;
    size_t n_out = (n - (size_t)T) / (size_t)D + 1;
    size_t T2 = (size_t)T * 2;
    float* hh = (float*)malloc(sizeof(float) * (T2 + 16));
    for (int t = 0; t < T; t++) { hh[2 * t] = taps[t]; hh[2 * t + 1] = taps[t]; }
    for (size_t k = 0; k < n_out; k++) {
        const float* xf = (const float*)(x + k * (size_t)D);
        float acc[16];
        for (int j = 0; j < 16; j++) acc[j] = 0.0f;
        size_t t = 0;
        for (; t + 16 <= T2; t += 16)
            for (int j = 0; j < 16; j++) acc[j] += xf[t + j] * hh[t + j];
        for (; t < T2; t++) acc[t & 15] += xf[t] * hh[t];
        float re = 0.0f, im = 0.0f;
        for (int j = 0; j < 16; j += 2) { re += acc[j]; im += acc[j + 1]; }
        y[k].re = re; y[k].im = im;
    }
    free(hh);
    return n_out;
}

/* FractionalDecimator — csdr/chain/selector.py:33,60 (COMPLEX_FLOAT), csdr/chain/analog.py:66,91
 * (FLOAT, prefilter=True); SURVEY A.8.  12-point Lagrange on integer nodes ih-5..ih+6,
 * where_m = 5 + m*rate kept in double (partition-invariant). */
static void lagrange12(float d, float* c)
{
    /* evaluation point xe = -d relative to node 5 (= ih); nodes x_i = i-5 */
    static float den[12];
    static int init = 0;
    if (!init) {
        for (int i = 0; i < 12; i++) {
            float p = 1.0f;
            for (int j = 0; j < 12; j++) if (j != i) p *= (float)(i - j);
            den[i] = p;
        }
        init = 1;
    }
    float xe = -d;
    for (int i = 0; i < 12; i++) {
        float p = 1.0f;
        for (int j = 0; j < 12; j++) if (j != i) p *= xe - (float)(j - 5);
        c[i] = p / den[i];
    }
}

size_t oc_fractional_decimator_cf(const oc_cf32* x, size_t n, double rate, oc_cf32* y, size_t cap)
{
    size_t m = 0;
    float c[12];
    for (;; m++) {
        double where = 5.0 + (double)m * rate;
        double ihd = ceil(where);
        size_t ih = (size_t)ihd;
        if (ih + 6 >= n || m >= cap) break;
        lagrange12((float)(ihd - where), c);
        float re = 0.0f, im = 0.0f;
        for (int i = 0; i < 12; i++) { re += c[i] * x[ih - 5 + i].re; im += c[i] * x[ih - 5 + i].im; }
        y[m].re = re; y[m].im = im;
    }
    return m;
}

size_t oc_fractional_decimator_f(const float* x, size_t n, double rate, const float* pre, int Tpre, float* y, size_t cap)
{
    size_t m = 0;
    float c[12];
    size_t extra = Tpre > 0 ? (size_t)(Tpre - 1) : 0;
    for (;; m++) {
        double where = 5.0 + (double)m * rate;
        double ihd = ceil(where);
        size_t ih = (size_t)ihd;
        if (ih + 6 + extra >= n || m >= cap) break;
        lagrange12((float)(ihd - where), c);
        float acc = 0.0f;
        for (int i = 0; i < 12; i++) {
            size_t idx = ih - 5 + (size_t)i;
            float v;
            if (Tpre > 0) {
                v = 0.0f;
                for (int t = 0; t < Tpre; t++) v += x[idx + (size_t)t] * pre[t];
            } else v = x[idx];
            acc += c[i] * v;
        }
        y[m] = acc;
    }
    return m;
}

/* Bandpass(transition=, use_fft=True) + setBandpass(lo,hi) — csdr/chain/selector.py:115-117,159-166;
 * mathematically a causal linear convolution with zero initial history. */
void oc_bandpass(const oc_cf32* x, size_t n, const oc_cf32* taps, int T, oc_cf32* y)
{
    for (size_t i = 0; i < n; i++) {
        float re = 0.0f, im = 0.0f;
        int tmax = (size_t)T <= i + 1 ? T : (int)(i + 1);
        for (int t = 0; t < tmax; t++) {
            const oc_cf32 v = x[i - (size_t)t], h = taps[t];
            re += h.re * v.re - h.im * v.im;
            im += h.re * v.im + h.im * v.re;
        }
        y[i].re = re; y[i].im = im;
    }
}

/* Squelch(Format.COMPLEX_FLOAT, length=, decimation=5, hangLength=, flushLength=, reportInterval=)
 * — csdr/chain/selector.py:119-130; SURVEY A.10.  SPEC-DEFINED: whole blocks only; a closed block
 * is emitted as zeros (stream stays continuous); hang counted in whole blocks. */
size_t oc_squelch(const oc_cf32* x, size_t n, int length, int decimation, int hang_length, float level,
                  int report_interval, oc_cf32* y, float* power_out, size_t power_cap, size_t* n_power)
{
    size_t nblocks = n / (size_t)length, np = 0;
    int hang_blocks = hang_length / length, hang = 0;
    for (size_t b = 0; b < nblocks; b++) {
        const oc_cf32* xb = x + b * (size_t)length;
        float p = 0.0f; int cnt = 0;
        for (int i = 0; i < length; i += decimation) { p += xb[i].re * xb[i].re + xb[i].im * xb[i].im; cnt++; }
        p /= (float)cnt;
        int open = 0;
        if (p >= level) { open = 1; hang = hang_blocks; }
        else if (hang > 0) { open = 1; hang--; }
        if (y) {
            if (open) memcpy(y + b * (size_t)length, xb, sizeof(oc_cf32) * (size_t)length);
            else memset(y + b * (size_t)length, 0, sizeof(oc_cf32) * (size_t)length);
        }
        if (report_interval > 0 && (b % (size_t)report_interval) == 0 && power_out && np < power_cap) power_out[np++] = p;
    }
    if (n_power) *n_power = np;
    return nblocks * (size_t)length;
}

/* ------------------------------------------------------------------------------------------------
 * Demodulators — csdr/chain/analog.py; SURVEY A.11
 * ---------------------------------------------------------------------------------------------- */
void oc_am_demod(const oc_cf32* x, size_t n, float* y)      /* AmDemod — analog.py:17 */
{
    for (size_t i = 0; i < n; i++) y[i] = sqrtf(x[i].re * x[i].re + x[i].im * x[i].im);
}

void oc_fm_demod(const oc_cf32* x, size_t n, float* y, oc_cf32* last)   /* FmDemod — analog.py:41,64 */
{
    const float K = 0.340447550238101026565118445432744920253753662109375f;
    oc_cf32 p = *last;
    for (size_t i = 0; i < n; i++) {
        float I = x[i].re, Q = x[i].im;
        float num = I * (Q - p.im) - Q * (I - p.re);
        float den = I * I + Q * Q;
        y[i] = den != 0.0f ? K * num / den : 0.0f;
        p = x[i];
    }
    *last = p;
}

void oc_limit(float* x, size_t n)                            /* Limit — analog.py:42,60 */
{
    for (size_t i = 0; i < n; i++) x[i] = x[i] > 1.0f ? 1.0f : (x[i] < -1.0f ? -1.0f : x[i]);
}

void oc_real_part(const oc_cf32* x, size_t n, float* y)     /* RealPart — analog.py:124 */
{
    for (size_t i = 0; i < n; i++) y[i] = x[i].re;
}

/* DcBlock — analog.py:18.  Upstream removes a block-wise ramped mean over whatever block the ring
 * buffer hands it; the oracle fixes the block at `block` samples — the Squelch block length
 * int(outputRate/16), which is the granularity Squelch writes downstream — and processes whole
 * blocks only, n must be a multiple of block. */
void oc_dc_block(const float* x, size_t n, int block, float* y, float* last_dc)
{
    float last = *last_dc;
    for (size_t b = 0; b + (size_t)block <= n; b += (size_t)block) {
        float avg = 0.0f;
        for (int i = 0; i < block; i++) avg += x[b + (size_t)i];
        avg /= (float)block;
        float diff = avg - last;
        for (int i = 0; i < block; i++)
            y[b + (size_t)i] = x[b + (size_t)i] - (last + diff * ((float)i / (float)block));
        last = avg;
    }
    *last_dc = last;
}

void oc_fir_f(const float* x, size_t n, const float* taps, int T, float* y)  /* NfmDeemphasis FIR */
{
    for (size_t i = 0; i < n; i++) {
        float acc = 0.0f;
        int tmax = (size_t)T <= i + 1 ? T : (int)(i + 1);
        for (int t = 0; t < tmax; t++) acc += taps[t] * x[i - (size_t)t];
        y[i] = acc;
    }
}

/* WfmDeemphasis(sampleRate, tau) — analog.py:67,85,92 */
void oc_wfm_deemphasis(const float* x, size_t n, int sample_rate, double tau, float* y, float* state)
{
    double dt = 1.0 / (double)sample_rate;
    float alpha = (float)(dt / (tau + dt));
    float om = 1.0f - alpha;
    float s = *state;
    for (size_t i = 0; i < n; i++) { s = alpha * x[i] + om * s; y[i] = s; }
    *state = s;
}

/* Agc(Format.FLOAT) + setProfile/setInitialGain/setMaxGain — analog.py:13-15,37-39,121-122.
 * SPEC-DEFINED (upstream constants unknown; luarvique changed them, CHANGELOG:71,107-108). */
void oc_agc_init(oc_agc* a, int profile, float initial_gain, float max_gain)
{
    a->reference = 0.8f;
    a->attack = 0.1f;
    a->decay = profile == 1 ? 0.001f : 0.0001f;
    a->hang_time = profile == 1 ? 200 : 600;
    a->hang_counter = 0;
    a->max_gain = max_gain > 0.0f ? max_gain : 65535.0f;
    a->gain = initial_gain > 0.0f ? initial_gain : 1.0f;
}

void oc_agc_process(oc_agc* a, const float* x, size_t n, float* y)
{
    float gain = a->gain; int hang = a->hang_counter;
    for (size_t i = 0; i < n; i++) {
        float v = x[i];
        if (v != 0.0f) {
            float err = fabsf(v) * gain / a->reference;
            if (err > 1.0f) { gain *= 1.0f - a->attack; hang = a->hang_time; }
            else if (hang > 0) hang--;
            else gain *= 1.0f + a->decay;
        }
        if (gain > a->max_gain) gain = a->max_gain;
        if (gain < 0.0f) gain = 0.0f;
        float o = v * gain;
        y[i] = o > 1.0f ? 1.0f : (o < -1.0f ? -1.0f : o);
    }
    a->gain = gain; a->hang_counter = hang;
}

/* Convert(Format.COMPLEX_SHORT, Format.COMPLEX_FLOAT) [+ Gain(Format.COMPLEX_FLOAT, g)] — the source-side format conversion,
 * owrx/source/fifi_sdr.py:27-28 wired by owrx/source/direct.py:59-71.  csdr: y = (float)x / SHRT_MAX.  n = number of REAL values. */
void oc_convert_s16_f(const int16_t* x, size_t n, float gain, float* y)
{
    for (size_t i = 0; i < n; i++) {
        float v = (float)x[i] / 32767.0f;
        y[i] = gain != 1.0f ? v * gain : v;
    }
}

/* Convert(Format.COMPLEX_CHAR / uint8 offset binary, ...): csdr  y = (float)x / (UCHAR_MAX / 2.0) - 1.0 */
void oc_convert_u8_f(const uint8_t* x, size_t n, float gain, float* y)
{
    for (size_t i = 0; i < n; i++) {
        float v = (float)x[i] / 127.5f - 1.0f;
        y[i] = gain != 1.0f ? v * gain : v;
    }
}

/* Convert(Format.FLOAT, Format.SHORT) — csdr/chain/clientaudio.py:12; SURVEY A.12 */
void oc_convert_f_s16(const float* x, size_t n, int16_t* y)
{
    for (size_t i = 0; i < n; i++) {
        float v = x[i] * 32767.0f;
        y[i] = v > 32767.0f ? 32767 : (v < -32768.0f ? -32768 : (int16_t)v);
    }
}

/* AdpcmEncoder(sync=True) — csdr/chain/clientaudio.py:34; framing pinned by
 * htdocs/lib/AudioEngine.js:449-491: "SYNC" + int16 LE stepIndex + int16 LE predictor, then 1001
 * data bytes (syncCounter=1000, post-decrement test).  n is rounded down to even. */
size_t oc_adpcm_sync_encode(const int16_t* s, size_t n, uint8_t* out, size_t cap)
{
    int index = 0, pred = 0;
    size_t o = 0, since_sync = 1001;
    for (size_t i = 0; i + 1 < n; i += 2) {
        if (since_sync == 1001) {
            if (o + 8 > cap) break;
            out[o++] = 'S'; out[o++] = 'Y'; out[o++] = 'N'; out[o++] = 'C';
            out[o++] = (uint8_t)(index & 0xff); out[o++] = (uint8_t)((index >> 8) & 0xff);
            out[o++] = (uint8_t)(pred & 0xff);  out[o++] = (uint8_t)((pred >> 8) & 0xff);
            since_sync = 0;
        }
        if (o + 1 > cap) break;
        uint8_t lo = ima_encode_sample(s[i], &index, &pred);
        uint8_t hi = ima_encode_sample(s[i + 1], &index, &pred);
        out[o++] = (uint8_t)(lo | (hi << 4));
        since_sync++;
    }
    return o;
}

/* Decimator parameter math — csdr/chain/selector.py:21-26,37-51 */
void oc_decimator_params(double input_rate, double output_rate, int* D, double* frac, double* transition, double* cutoff)
{
    if (output_rate > input_rate) output_rate = input_rate;
    double d = input_rate / output_rate;
    int di = (int)d;
    *D = di;
    *frac = (input_rate / di) / output_rate;
    *transition = 0.15 * (output_rate / input_rate);
    *cutoff = 0.5 * di / (input_rate / output_rate);
}

/* Whole client chain over a finite record: Selector (csdr/chain/selector.py:89-113) then the
 * demodulator chain (csdr/chain/analog.py), one full pass and one buffer per stage like the
 * reference's Chain._connect (csdr/chain/__init__.py:21-25). */
int oc_client_chain_run(const oc_chain_cfg* cfg, const oc_cf32* iq, size_t n,
                        oc_cf32* if_out, size_t if_cap, float* demod_out, size_t demod_cap,
                        float* audio_out, size_t audio_cap, oc_chain_counts* counts)
{
    int D, T; double frac, transition, cutoff;
    oc_chain_counts cnt = {0, 0, 0};
    oc_decimator_params(cfg->input_rate, cfg->output_rate, &D, &frac, &transition, &cutoff);
    T = oc_filter_len(transition);
    float* taps = (float*)malloc(sizeof(float) * (size_t)T);
    oc_firdes_lowpass(taps, T, cutoff / D);

    /* Shift */
    oc_cf32* shifted = (oc_cf32*)malloc(sizeof(oc_cf32) * (n + 1));
    oc_shift(iq, shifted, n, -cfg->offset_hz / cfg->input_rate, 0.0, 0, cfg->fast_shift);
    /* FirDecimate */
    size_t n1cap = n / (size_t)D + 2;
    oc_cf32* s1 = (oc_cf32*)malloc(sizeof(oc_cf32) * n1cap);
    size_t n1 = oc_fir_decimate(shifted, n, taps, T, D, s1);
    free(shifted); free(taps);
    /* FractionalDecimator */
    oc_cf32* s2 = s1; size_t n2 = n1;
    if (frac != 1.0) {
        s2 = (oc_cf32*)malloc(sizeof(oc_cf32) * (n1 + 2));
        n2 = oc_fractional_decimator_cf(s1, n1, frac, s2, n1 + 2);
        free(s1);
    }
    /* Bandpass */
    oc_cf32* s3 = s2;
    if (cfg->bp_lo_hz < cfg->bp_hi_hz) {
        int Tb = oc_filter_len(320.0 / cfg->output_rate);
        oc_cf32* bt = (oc_cf32*)malloc(sizeof(oc_cf32) * (size_t)Tb);
        oc_firdes_bandpass(bt, Tb, cfg->bp_lo_hz / cfg->output_rate, cfg->bp_hi_hz / cfg->output_rate);
        s3 = (oc_cf32*)malloc(sizeof(oc_cf32) * (n2 + 1));
        oc_bandpass(s2, n2, bt, Tb, s3);
        free(bt); free(s2);
    }
    cnt.n_if = n2;
    if (if_out) memcpy(if_out, s3, sizeof(oc_cf32) * (n2 < if_cap ? n2 : if_cap));

    /* Squelch (selector.py:119-130) at its default level passes whole blocks of int(outputRate/16)
     * samples; the trailing partial block is withheld.  DcBlock then sees exactly those blocks. */
    int sq_len = (int)(cfg->output_rate / 16.0);
    if (sq_len < 1) sq_len = 1;
    n2 -= n2 % (size_t)sq_len;

    /* demodulator */
    float* d0 = (float*)malloc(sizeof(float) * (n2 + 1));
    float* d1 = (float*)malloc(sizeof(float) * (n2 + 1));
    size_t nd = 0; int have_agc = 0; oc_agc agc;
    switch (cfg->demod) {
    case OC_DEMOD_NFM: {
        oc_cf32 last = {0.0f, 0.0f};
        oc_fm_demod(s3, n2, d0, &last);
        oc_limit(d0, n2);
        int Td = oc_nfm_deemphasis_len((int)cfg->output_rate);
        float* dt = (float*)malloc(sizeof(float) * (size_t)Td);
        oc_nfm_deemphasis_taps(dt, Td, (int)cfg->output_rate);
        oc_fir_f(d0, n2, dt, Td, d1);
        free(dt);
        nd = n2; have_agc = 1;
        oc_agc_init(&agc, cfg->agc_profile, 1.0f, 3.0f);         /* analog.py:37-39 */
        break; }
    case OC_DEMOD_AM: {
        oc_am_demod(s3, n2, d0);
        float last_dc = 0.0f;
        nd = n2;
        oc_dc_block(d0, nd, sq_len, d1, &last_dc);
        have_agc = 1;
        oc_agc_init(&agc, cfg->agc_profile, 200.0f, 65535.0f);   /* analog.py:13-15 */
        break; }
    case OC_DEMOD_SSB: {
        oc_real_part(s3, n2, d1);
        nd = n2; have_agc = 1;
        oc_agc_init(&agc, cfg->agc_profile, 1.0f, 65535.0f);     /* analog.py:121-122 */
        break; }
    case OC_DEMOD_WFM: {
        oc_cf32 last = {0.0f, 0.0f};
        oc_fm_demod(s3, n2, d0, &last);
        oc_limit(d0, n2);
        double r = cfg->output_rate / cfg->audio_rate;            /* analog.py:66: 250000/sampleRate */
        int Tp = oc_filter_len(0.03);
        float* pt = (float*)malloc(sizeof(float) * (size_t)Tp);
        oc_firdes_lowpass(pt, Tp, 0.5 / (r - 0.03));
        float* d2 = (float*)malloc(sizeof(float) * (n2 + 1));
        size_t n3 = oc_fractional_decimator_f(d0, n2, r, pt, Tp, d2, n2 + 1);
        float st = 0.0f;
        oc_wfm_deemphasis(d2, n3, (int)cfg->audio_rate, cfg->wfm_tau, d1, &st);
        free(pt); free(d2);
        nd = n3; have_agc = 0;
        break; }
    default:
        nd = 0;
    }
    cnt.n_demod = nd;
    if (demod_out) memcpy(demod_out, d1, sizeof(float) * (nd < demod_cap ? nd : demod_cap));
    if (have_agc) {
        oc_agc_process(&agc, d1, nd, d0);
        if (audio_out) memcpy(audio_out, d0, sizeof(float) * (nd < audio_cap ? nd : audio_cap));
    } else if (audio_out) {
        memcpy(audio_out, d1, sizeof(float) * (nd < audio_cap ? nd : audio_cap));
    }
    cnt.n_audio = nd;
    free(d0); free(d1); free(s3);
    if (counts) *counts = cnt;
    return 0;
}
