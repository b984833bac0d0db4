"""openwebrx_b200 — B200-native (sm_100a) DSP hot path of OpenWebRX+.

Host-side mirror of the reference's hot-path objects on top of libowrx_b200.so:
  Waterfall    <- FftChain            (reference csdr/chain/fft.py:25-96)
  ChannelBank  <- N x (Selector + analog demodulator chain)
                                       (reference csdr/chain/selector.py:89-214, csdr/chain/analog.py:11-127)
The drop-in `pycsdr` package at the repository root binds the same library so that the reference's
unmodified csdr.chain classes run on top of it.
"""
from .params import fftchain_params, decimator_params, bandpass_params, squelch_params  # noqa: F401
from .waterfall import Waterfall  # noqa: F401
from .bank import ChannelBank  # noqa: F401

__all__ = ["Waterfall", "ChannelBank", "fftchain_params", "decimator_params", "bandpass_params", "squelch_params"]
