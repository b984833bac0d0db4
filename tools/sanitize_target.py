#!/usr/bin/env python
"""A small run of every kernel family and every stream-overlap mode, sized for compute-sanitizer (10-100x slower than native):

  compute-sanitizer --tool memcheck  python tools/sanitize_target.py      (one tool per gpurun call, B200_PROFILING.md)
  compute-sanitizer --tool racecheck python tools/sanitize_target.py

Covers: smoke() (waterfall + 3 channels through the host API); the three-stream pipelined device path
(owrx_bank_set_pipelined) over several blocks with a retune, a band-pass change and an added client in between; the host path
in deferred-drain mode with chunked uploads; all three Shift + FirDecimate forms; the WFM chain with the partitioned-FFT
band-pass; the four-step 8192-point waterfall with the side-stream ADPCM encoder; raw int16 ingress; the client audio tail."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OWRX_FEED_CHUNK_LOG2", "16")          # several upload chunks per (small) feed

import torch                                                   # noqa: E402
import __graft_entry__                                         # noqa: E402
from openwebrx_b200 import ChannelBank, Waterfall, _native as N, fftchain_params   # noqa: E402
from openwebrx_b200.synth import BANDPASS, carrier_plan, make_iq                   # noqa: E402


def main():
    __graft_entry__.smoke()
    dev = torch.device("cuda", 0)
    fs = 2.4e6
    cars = carrier_plan(6, fs, seed=11)
    n = 5333 + 200 * 900
    iq = make_iq(3 * n, fs, cars, seed=11)
    st = torch.cuda.Stream(device=dev)
    for mode in ("direct", "fastconv", "fastconv_tc"):
        bank = ChannelBank(fs, device=0, outputs=N.OUT_AUDIO | N.OUT_DEMOD | N.OUT_IF | N.OUT_POWER)
        bank.set_fir_mode(mode)
        chans = [bank.add_channel(12000, demod=c["kind"], offset=c["offset"], bandpass=BANDPASS[c["kind"]]) for c in cars]
        chans[0].setAudioFormat("adpcm")
        # ---- pipelined device path: block i+1's FIR beside block i's low-rate stages beside block i-1's Agc
        bank.set_pipelined(True)
        d_iq = torch.from_numpy(iq.view(np.float32).copy()).to(dev)
        for b in range(3):
            bank.process_device(d_iq[2 * b * n:2 * (b + 1) * n], n, stream=st.cuda_stream)
            if b == 0:
                chans[1].setFrequencyOffset(cars[1]["offset"] + 100)
                chans[2].setBandpass(-3000, 3000)
            if b == 1:
                extra = bank.add_channel(12000, demod="am", offset=1000, bandpass=BANDPASS["am"])
        bank.join(st.cuda_stream)
        st.synchronize()
        bank.drain()
        got = [len(c.read_audio()) for c in chans]
        assert min(got) > 0, got
        extra.remove()
        bank.set_pipelined(False)
        # ---- host path, streaming (deferred drain), ragged feeds
        bank.set_deferred_drain(True)
        for part in np.array_split(iq, 5):
            bank.feed(part)
        bank.flush()
        assert all(len(c.read_audio()) > 0 for c in chans[1:])
        assert len(chans[0].read_bytes()) > 0
        bank.close()
    # ---- WFM (partitioned-FFT band-pass, prefilter, Lagrange, de-emphasis), raw int16 ingress
    fs2 = 2.0e6
    cw = carrier_plan(2, fs2, seed=12, wfm=True, span=0.3)
    x = make_iq(213 + 8 * (15625 * 2 + 50), fs2, cw, seed=12)
    bank = ChannelBank(fs2, device=0)
    ch = [bank.add_channel(250000, demod="wfm", offset=c["offset"], bandpass=BANDPASS["wfm"], audio_rate=48000.0, tau=50e-6) for c in cw]
    raw = (np.clip(x.view(np.float32), -1, 1) * 32767).astype(np.int16)
    bank.feed_raw(raw, "cs16")
    assert all(len(c.read_audio()) > 1000 for c in ch)
    bank.close()
    # ---- waterfall: four-step 8192 points, side-stream ADPCM, streaming host feed
    nfft = 8192
    wf = Waterfall(fs, nfft, 0.3, 30, "adpcm", device=0)
    avg, every_n = fftchain_params(fs, nfft, 0.3, 30)
    y = make_iq(every_n * avg * 4 + nfft, fs, cars, seed=13)
    d_y = torch.from_numpy(y.view(np.float32).copy()).to(dev)
    out = torch.zeros(4 * wf.line_bytes, dtype=torch.uint8, device=dev)
    wf.set_pipelined(True)
    for _ in range(3):
        assert wf.process_device(d_y, len(y), out, out.numel(), stream=st.cuda_stream) == 4
    wf.join(st.cuda_stream)
    st.synchronize()
    wf.set_pipelined(False)
    lines = []
    for part in np.array_split(y, 7):
        lines += wf.feed(part)
    assert len(lines) == 4
    wf.close()
    torch.cuda.synchronize()
    print("sanitize target ok: %d kernel launches" % N.lib.owrx_launch_count())


if __name__ == "__main__":
    main()
