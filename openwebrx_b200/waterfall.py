"""Waterfall — host-side mirror of the reference's FftChain (csdr/chain/fft.py:25-96) on libowrx_b200."""
import ctypes as C

import numpy as np

from . import _native as N
from .params import fftchain_params


def _ptr(obj):
    """device pointer of a torch tensor / raw int."""
    if obj is None:
        return None
    if hasattr(obj, "data_ptr"):
        return obj.data_ptr()
    return int(obj)


class Waterfall:
    """Fft -> LogAveragePower|LogPower -> FftSwap -> [FftAdpcm] as one GPU plan.

    Constructor arguments are FftChain's (csdr/chain/fft.py:26): samp_rate, fft_size,
    fft_v_overlap_factor, fft_fps, fft_compression ("adpcm" | "none").
    """

    def __init__(self, samp_rate, fft_size, fft_v_overlap_factor, fft_fps, fft_compression="adpcm", device=0,
                 add_db=-70.0):
        self.sampleRate = samp_rate
        self.size = fft_size
        self.vOverlapFactor = fft_v_overlap_factor
        self.fps = fft_fps
        self.compression = fft_compression
        self.fftAverages, self.blockSize = fftchain_params(samp_rate, fft_size, fft_v_overlap_factor, fft_fps)
        h = C.c_void_p()
        N.check(N.lib.owrx_wf_create(device, fft_size, self.blockSize, self.fftAverages, add_db,
                                     self._comp(fft_compression), C.byref(h)))
        self._h = h

    @staticmethod
    def _comp(name):
        if name == "adpcm":
            return N.COMPRESSION_ADPCM
        if name == "none":
            return N.COMPRESSION_NONE
        raise ValueError("unknown fft_compression %r" % (name,))

    # --- FftChain setters (csdr/chain/fft.py:57-96)
    def _update(self):
        avg, every_n = fftchain_params(self.sampleRate, self.size, self.vOverlapFactor, self.fps)
        if avg != self.fftAverages:
            self.fftAverages = avg
            N.check(N.lib.owrx_wf_set_avg_number(self._h, avg))
        if every_n != self.blockSize:
            self.blockSize = every_n
            N.check(N.lib.owrx_wf_set_every_n_samples(self._h, every_n))

    def setVOverlapFactor(self, v):
        self.vOverlapFactor = v
        self._update()

    def setFps(self, fps):
        self.fps = fps
        self._update()

    def setSampleRate(self, samp_rate):
        self.sampleRate = samp_rate
        self._update()

    def setCompression(self, compression):
        N.check(N.lib.owrx_wf_set_compression(self._h, self._comp(compression)))
        self.compression = compression

    def set_noise_filter(self, enable=True, alpha=0.9, beta=0.05, growth=0.02):
        """Spectral-subtraction noise filter on the averaged line power (BASELINE config 4).  An extension: the reference's
        FftChain has no such stage; spec in include/owrx_b200.h, oracle in oracle.fftchain_run(noise_filter=)."""
        N.check(N.lib.owrx_wf_set_noise_filter(self._h, 1 if enable else 0, float(alpha), float(beta), float(growth)))

    # --- data path
    @property
    def line_bytes(self):
        return N.lib.owrx_wf_line_bytes(self._h)

    def lines_for(self, n_samples):
        return N.lib.owrx_wf_lines_for(self._h, n_samples)

    def feed(self, iq):
        """iq: complex64 numpy array (host). Returns a list of completed lines (bytes objects)."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        N.check(N.lib.owrx_wf_feed(self._h, iq.ctypes.data_as(C.c_void_p), iq.size))
        return self.read()

    def feed_raw(self, raw, fmt, gain=1.0):
        """RAW source samples from host memory: fmt "cs16" (int16, interleaved I, Q) or "cu8" (uint8, offset binary); the
        reference's CPU-side Convert (+ Gain) (owrx/source/fifi_sdr.py:27-28) runs on the GPU.  Returns completed lines."""
        raw = np.ascontiguousarray(raw, dtype=np.int16 if fmt == "cs16" else np.uint8)
        assert raw.size % 2 == 0
        N.check(N.lib.owrx_wf_feed_fmt(self._h, raw.ctypes.data_as(C.c_void_p), raw.size // 2, N.IQ_FORMATS[fmt], float(gain)))
        return self.read()

    def read_messages(self):
        """completed lines as websocket messages: b"\x01" + line each (write_spectrum_data, owrx/connection.py:473-475)"""
        out = []
        buf = np.empty(self.line_bytes + 1, np.uint8)
        while True:
            n = C.c_size_t()
            N.check(N.lib.owrx_wf_read_message(self._h, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(n)))
            if n.value == 0:
                return out
            out.append(buf[:n.value].tobytes())

    def read(self):
        lb = self.line_bytes
        out = []
        buf = np.empty(lb * 64, np.uint8)
        while True:
            n = C.c_size_t()
            N.check(N.lib.owrx_wf_read(self._h, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(n)))
            if n.value == 0:
                break
            for k in range(n.value // lb):
                out.append(buf[k * lb:(k + 1) * lb].tobytes())
        return out

    def process_device(self, iq_dev, n_samples, out_dev, out_cap_bytes, db_dev=None, s16_dev=None, stream=None):
        """Device-resident batch: pointers are device addresses (ints or torch tensors). Returns n_lines."""
        n = C.c_size_t()
        N.check(N.lib.owrx_wf_process_device(self._h, _ptr(iq_dev), n_samples, _ptr(out_dev), out_cap_bytes, _ptr(db_dev),
                                             _ptr(s16_dev), C.byref(n), _ptr(stream)))
        return n.value

    def set_pipelined(self, enable=True):
        N.check(N.lib.owrx_wf_set_pipelined(self._h, 1 if enable else 0))

    def join(self, stream=None):
        N.check(N.lib.owrx_wf_join(self._h, _ptr(stream)))

    def close(self):
        if getattr(self, "_h", None):
            N.lib.owrx_wf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fft_adpcm_encode_device(s16_dev, fft_size, n_lines, out_dev, device=0, stream=None):
    N.check(N.lib.owrx_fft_adpcm_encode_device(device, _ptr(s16_dev), fft_size, n_lines, _ptr(out_dev), _ptr(stream)))
