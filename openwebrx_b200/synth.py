"""Synthetic wideband IQ for tests and bench.py: white noise floor plus K
AM / NFM / USB (or WFM) carriers; channel c tunes to carrier c mod K.  numpy only."""
import numpy as np

SEED = 20260101


def carrier_plan(k, fs, seed=SEED, wfm=False, span=0.45):
    """Deterministic carrier table: offsets (integer Hz, as clients tune in Hz: owrx/dsp.py:457),
    amplitudes 10^(U(-50,-6)/20), kind cycling am/nfm/usb (or all wfm)."""
    rng = np.random.default_rng(seed)
    offs = np.round(rng.uniform(-span * fs, span * fs, k)).astype(np.int64)
    amps = 10.0 ** (rng.uniform(-50.0, -6.0, k) / 20.0)
    kinds = ["wfm"] * k if wfm else [("am", "nfm", "usb")[i % 3] for i in range(k)]
    total = amps.sum()
    if total > 0.9:
        amps = amps * (0.9 / total)
    return [dict(offset=int(o), amp=float(a), kind=kd) for o, a, kd in zip(offs, amps, kinds)]


def make_iq(n, fs, carriers, seed=SEED, noise=1e-3, t0=0):
    """complex64 array of n samples starting at absolute sample index t0."""
    rng = np.random.default_rng(seed + 1 + (t0 % 1000003))
    x = (noise * rng.standard_normal(n) + 1j * noise * rng.standard_normal(n)).astype(np.complex64)
    chunk = 1 << 20
    for c in carriers:
        f, a, kind = c["offset"], c["amp"], c["kind"]
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            t = (np.arange(s, e, dtype=np.float64) + t0) / fs
            if kind == "am":
                sig = a * (1.0 + 0.5 * np.cos(2 * np.pi * 1000.0 * t)) / 1.5 * np.exp(2j * np.pi * f * t)
            elif kind == "nfm":
                sig = a * np.exp(1j * (2 * np.pi * f * t + 2.5 * np.sin(2 * np.pi * 1000.0 * t)))
            elif kind == "wfm":
                sig = a * np.exp(1j * (2 * np.pi * f * t + 75.0 * np.sin(2 * np.pi * 1000.0 * t)))
            elif kind == "amfm":
                # one carrier that carries audio for an envelope detector (1 kHz, m = 0.5) AND a discriminator (700 Hz, +-2.5 kHz)
                sig = a * (1.0 + 0.5 * np.cos(2 * np.pi * 1000.0 * t)) / 1.5 * np.exp(1j * (2 * np.pi * f * t + (2500.0 / 700.0) * np.sin(2 * np.pi * 700.0 * t)))
            elif kind == "usb":
                sig = a * 0.5 * (np.exp(2j * np.pi * (f + 700.0) * t) + np.exp(2j * np.pi * (f + 1900.0) * t))
            else:
                sig = a * np.exp(2j * np.pi * f * t)
            x[s:e] += sig.astype(np.complex64)
    return x


# default band-passes per mode: reference owrx/modes.py:124-128
BANDPASS = {"nfm": (-5999, 5999), "am": (-4700, 4700), "usb": (150, 3000), "lsb": (-3000, -150), "wfm": (-124000, 124000)}
