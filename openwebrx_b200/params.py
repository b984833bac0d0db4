"""Host-side parameter math of the hot path, restated from the reference's chain classes.

Pure Python, no GPU: used by bench.py / tests to derive the same shapes the reference's own
FftChain / Decimator / Selector objects would hand to pycsdr.
"""


def fftchain_params(samp_rate, fft_size, fft_voverlap_factor, fft_fps):
    """FftChain._updateParameters — reference csdr/chain/fft.py:75-85.
    Returns (fft_averages, every_n_samples)."""
    avg = 0
    if fft_voverlap_factor > 0:
        avg = int(round(1.0 * samp_rate / fft_size / fft_fps / (1.0 - fft_voverlap_factor)))
    if avg == 0:
        every_n = int(samp_rate / fft_fps)
    else:
        every_n = int(samp_rate / fft_fps / avg)
    return avg, every_n


def decimator_params(input_rate, output_rate):
    """Decimator.__init__/_getDecimation — reference csdr/chain/selector.py:21-26,37-51.
    Returns (decimation, fraction, transition, cutoff)."""
    if output_rate > input_rate:
        output_rate = input_rate
    d = input_rate / output_rate
    d_int = int(d)
    d_float = float(input_rate / d_int) / output_rate
    transition = 0.15 * (output_rate / float(input_rate))
    cutoff = 0.5 * d_int / (input_rate / output_rate)
    return d_int, d_float, transition, cutoff


def filter_length(transition):
    """csdr filter length rule: odd(int(4 / transition))."""
    n = int(4.0 / transition)
    return n + 1 if n % 2 == 0 else n


def bandpass_params(output_rate, low_cut, high_cut):
    """Selector._buildBandpass + setBandpass — reference csdr/chain/selector.py:115-117,159-166.
    Returns (transition, lo_rate, hi_rate)."""
    return 320.0 / output_rate, low_cut / output_rate, high_cut / output_rate


def squelch_params(output_rate, measurements_per_sec=16, readings_per_sec=4):
    """Selector._buildSquelch — reference csdr/chain/selector.py:119-130.
    Returns dict(length, decimation, hangLength, flushLength, reportInterval)."""
    block = int(output_rate / measurements_per_sec)
    return dict(length=block, decimation=5, hangLength=2 * block, flushLength=5 * block,
                reportInterval=int(measurements_per_sec / readings_per_sec))


def shift_rate(frequency_offset, input_rate):
    """Selector._updateShift — reference csdr/chain/selector.py:138-140."""
    return -frequency_offset / input_rate
