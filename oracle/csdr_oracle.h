/*
 * csdr_oracle.h — CPU restatement of the pycsdr/libcsdr arithmetic that OpenWebRX+ wires
 * together on its DSP hot path (FftChain + Selector + analog demodulators).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (openwebrx_b200/, pycsdr/) never links, imports or calls anything in oracle/.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in luarvique/csdr + luarvique/pycsdr
 * (branch master, no commit pin; debian/control:22 requires python3-csdr >= 0.18.36), which are
 * NOT under /root/reference and cannot be built here (FFTW3/libsamplerate absent, no network).
 * The reference's own tests (test/property) never touch DSP.  This file therefore restates the
 * published csdr algorithms and is anchored on
 *   - the reference's call sites (cited per function),
 *   - the in-tree browser decoder htdocs/lib/AudioEngine.js:410-509 (IMA-ADPCM tables, nibble
 *     order, SYNC framing) and htdocs/openwebrx.js:845,1117-1131 (10-sample pad, /100 scaling),
 *   - numpy / scipy cross-checks in tests/.
 * Stages whose upstream constants are not recoverable (Agc, NfmDeemphasis taps, Squelch hang/flush,
 * DcBlock partitioning) are SPEC-DEFINED here and labelled so.
 */
#ifndef CSDR_ORACLE_H
#define CSDR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } oc_cf32;

/* ---------- filter / window design ---------- */
int  oc_filter_len(double transition);                                  /* odd(int(4/transition)) */
void oc_firdes_lowpass(float* taps, int len, double cutoff_rate);       /* Hamming windowed sinc, sum = 1 */
void oc_firdes_bandpass(oc_cf32* taps, int len, double lo, double hi);  /* LPF((hi-lo)/2) modulated to (hi+lo)/2 */
void oc_fft_window_hamming(float* w, int n);                            /* 0.54-0.46cos(2 pi n/(N-1)) */
int  oc_nfm_deemphasis_len(int sample_rate);                            /* spec-defined */
void oc_nfm_deemphasis_taps(float* taps, int len, int sample_rate);     /* spec-defined */

/* ---------- FftChain stages: csdr/chain/fft.py:18-22,34-45 ---------- */
void oc_fft_forward(const oc_cf32* in, oc_cf32* out, int n);            /* unnormalised forward DFT, n = 2^k */
void oc_fft_frame(const oc_cf32* x, const float* window, oc_cf32* out, int n);   /* Fft: window then DFT */
void oc_log_power(const oc_cf32* X, float* out, int n, float add_db);            /* LogPower */
void oc_log_average_power(const oc_cf32* frames, int avg, float* out, int n, float add_db); /* LogAveragePower */
void oc_fft_swap(const float* in, float* out, int n);                   /* FftSwap */
void oc_fft_adpcm_quantise(const float* db, int16_t* s, int n);         /* s[10+i]=(int16)(db*100), s[0..9]=s[10] */
void oc_ima_adpcm_encode(const int16_t* s, int n, uint8_t* out, int* index, int* predictor); /* n even */
void oc_ima_adpcm_decode(const uint8_t* in, int nbytes, int16_t* out, int* index, int* predictor);
void oc_fft_adpcm(const float* db, uint8_t* out, int n);                /* FftAdpcm: (n+10)/2 bytes, reset state */

/* whole FftChain over a finite IQ record (csdr/chain/fft.py:25-49): returns number of lines.
   compression: 0 -> float32 lines (4n bytes each), 1 -> adpcm ((n+10)/2 bytes each).
   avg == 0 selects LogPower (one line per frame).  If s16_out != NULL it receives the quantised
   int16 lines (n+10 each) that fed the ADPCM encoder. */
size_t oc_fftchain_run(const oc_cf32* iq, size_t n_samples, int n, int every_n, int avg, float add_db,
                       int compression, uint8_t* out, size_t out_cap, int16_t* s16_out, float* db_out);

/* spec-defined waterfall noise filter (BASELINE config 4; no reference counterpart, see csdr_oracle.c): in-place on the
   averaged linear power of one line, `noise` = per-bin floor estimate carried across lines */
void oc_wf_noise_filter(float* pw, float* noise, int n, int first, float alpha, float beta, float growth);
size_t oc_fftchain_run_nf(const oc_cf32* iq, size_t n_samples, int n, int every_n, int avg, float add_db,
                          int compression, uint8_t* out, size_t out_cap, int16_t* s16_out, float* db_out,
                          float nf_alpha, float nf_beta, float nf_growth);

/* ---------- Selector stages: csdr/chain/selector.py:29,33,95,115-130 ---------- */
/* Shift: y[i] = x[i] * exp(j 2 pi frac(phase0 + rate*(n0+i+1))); returns nothing, pure function of
   the absolute sample index.  fast != 0 uses a float32 rotation recurrence re-seeded every 256
   samples (CPU-baseline speed mode). */
void oc_shift(const oc_cf32* x, oc_cf32* y, size_t n, double rate, double phase0_turns, uint64_t n0, int fast);
/* FirDecimate: y[k] = sum_t x[kD+t] h[t], k < n_out where n_out = floor((n-T)/D)+1 (n >= T) */
size_t oc_fir_decimate(const oc_cf32* x, size_t n, const float* taps, int T, int D, oc_cf32* y);
/* FractionalDecimator over a finite record; where_m = 5 + m*rate (double). prefilter taps may be NULL. */
size_t oc_fractional_decimator_cf(const oc_cf32* x, size_t n, double rate, oc_cf32* y, size_t cap);
size_t oc_fractional_decimator_f(const float* x, size_t n, double rate, const float* pre, int Tpre, float* y, size_t cap);
/* Bandpass: causal convolution with zero history */
void oc_bandpass(const oc_cf32* x, size_t n, const oc_cf32* taps, int T, oc_cf32* y);
/* Squelch (spec-defined): block power over every `decimation`-th sample; gate + hang; power reports */
size_t oc_squelch(const oc_cf32* x, size_t n, int length, int decimation, int hang_length, float level,
                  int report_interval, oc_cf32* y, float* power_out, size_t power_cap, size_t* n_power);

/* ---------- demodulators: csdr/chain/analog.py:11-127 ---------- */
void oc_am_demod(const oc_cf32* x, size_t n, float* y);
void oc_fm_demod(const oc_cf32* x, size_t n, float* y, oc_cf32* last);   /* last carried; init {0,0} */
void oc_limit(float* x, size_t n);
void oc_real_part(const oc_cf32* x, size_t n, float* y);
void oc_dc_block(const float* x, size_t n, int block, float* y, float* last_dc); /* whole blocks only; returns via y */
void oc_fir_f(const float* x, size_t n, const float* taps, int T, float* y);     /* causal, zero history */
void oc_wfm_deemphasis(const float* x, size_t n, int sample_rate, double tau, float* y, float* state);
typedef struct { float reference, attack, decay, max_gain, gain; int hang_time, hang_counter; } oc_agc;
void oc_agc_init(oc_agc* a, int profile /*0 slow,1 fast*/, float initial_gain, float max_gain);
void oc_agc_process(oc_agc* a, const float* x, size_t n, float* y);
void oc_convert_s16_f(const int16_t* x, size_t n, float gain, float* y);   /* Convert(COMPLEX_SHORT -> COMPLEX_FLOAT) + Gain */
void oc_convert_u8_f(const uint8_t* x, size_t n, float gain, float* y);     /* Convert(uint8 offset binary -> float) + Gain */
void oc_convert_f_s16(const float* x, size_t n, int16_t* y);
/* AdpcmEncoder(sync=True): csdr/chain/clientaudio.py:34; framing htdocs/lib/AudioEngine.js:449-491 */
size_t oc_adpcm_sync_encode(const int16_t* s, size_t n, uint8_t* out, size_t cap);

/* ---------- whole client chain over a finite IQ record (owrx/dsp.py:39-72 wiring) ---------- */
enum { OC_DEMOD_NFM = 0, OC_DEMOD_AM = 1, OC_DEMOD_SSB = 2, OC_DEMOD_WFM = 3, OC_DEMOD_NONE = 4 };
typedef struct {
    double input_rate;      /* wideband sample rate */
    double output_rate;     /* selector output (IF) rate: 12000, or 250000 for WFM */
    double offset_hz;       /* frequencyOffset */
    double bp_lo_hz, bp_hi_hz;  /* bandpass cutoffs in Hz; lo >= hi disables */
    int    demod;           /* OC_DEMOD_* */
    double audio_rate;      /* WFM only: hd_output_rate (48000) */
    double wfm_tau;         /* WFM only */
    int    agc_profile;     /* 0 slow, 1 fast */
    int    fast_shift;
} oc_chain_cfg;
typedef struct {
    size_t n_if;            /* IF samples written (selector output, post band-pass)  */
    size_t n_demod;         /* pre-AGC demodulator output samples */
    size_t n_audio;         /* post-AGC samples */
} oc_chain_counts;
/* Runs Shift -> FirDecimate -> [FractionalDecimator] -> [Bandpass] -> demod stages.
   Any output pointer may be NULL. Capacities are in samples. */
int oc_client_chain_run(const oc_chain_cfg* cfg, const oc_cf32* iq, size_t n,
                        oc_cf32* if_out, size_t if_cap, float* demod_out, size_t demod_cap,
                        float* audio_out, size_t audio_cap, oc_chain_counts* counts);

/* Decimator parameter math (csdr/chain/selector.py:21-26,37-51) */
void oc_decimator_params(double input_rate, double output_rate, int* D, double* frac, double* transition, double* cutoff);
/* FftChain parameter math (csdr/chain/fft.py:75-85) */
void oc_fftchain_params(double samp_rate, int fft_size, double voverlap, double fps, int* avg, int* every_n);


#ifdef __cplusplus
}
#endif
#endif
