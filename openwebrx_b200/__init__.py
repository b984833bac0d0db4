"""openwebrx_b200 — B200-native (sm_100a) DSP hot path of OpenWebRX+.

Host-side mirror of the reference's hot-path objects on top of libowrx_b200.so:
  Waterfall    <- FftChain            (reference csdr/chain/fft.py:25-96)
  ChannelBank  <- N x (Selector + analog demodulator chain)
                                       (reference csdr/chain/selector.py:89-214, csdr/chain/analog.py:11-127)
The drop-in `pycsdr` package at the repository root binds the same library so that the reference's
unmodified csdr.chain classes run on top of it.
"""
from .params import fftchain_params, decimator_params, bandpass_params, squelch_params  # noqa: F401

__all__ = ["Waterfall", "ChannelBank", "fftchain_params", "decimator_params", "bandpass_params", "squelch_params"]


def __getattr__(name):
    # The CUDA library is loaded when a class that computes is first asked for — not when the pure-Python helpers
    # (params, synth, sharding) are imported: bench.py's CPU arm (--impl reference) must not map libowrx_b200.so.
    # There is still no CPU fallback: asking for Waterfall / ChannelBank without the built library raises ImportError.
    if name == "Waterfall":
        from .waterfall import Waterfall
        return Waterfall
    if name == "ChannelBank":
        from .bank import ChannelBank
        return ChannelBank
    if name == "_native":
        import importlib
        return importlib.import_module("._native", __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
