#!/usr/bin/env python
"""Device-resident throughput of the other BASELINE configs (not the headline bench line): C3 per-GPU share
(61.44 MS/s, 128 of the 1024 channels), C5 (WFM x128 from 20 MS/s), C4 (65536-pt waterfall).  Prints one JSON
object per config; used for DESIGN.md's table."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwebrx_b200 import ChannelBank, Waterfall, fftchain_params   # noqa: E402
from openwebrx_b200.synth import BANDPASS, carrier_plan                # noqa: E402


def time_bank(fs, out_rate, n_ch, kinds, block, steps=5, **kw):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    iq = 1e-3 * torch.randn(block, 2, device=dev, generator=g)
    t = torch.arange(block, device=dev, dtype=torch.float32)
    iq[:, 0] += 0.2 * torch.cos(0.3 * t); iq[:, 1] += 0.2 * torch.sin(0.3 * t)
    del t
    cars = carrier_plan(n_ch, fs, seed=3, wfm=(kinds == "wfm"))
    bank = ChannelBank(fs)
    for c in cars:
        kind = c["kind"]
        bank.add_channel(out_rate, demod=kind, offset=c["offset"], bandpass=BANDPASS[kind], **kw)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    bank.set_pipelined(True)
    bank.profile(True)
    for _ in range(3):
        bank.process_device(iq, block, stream=st.cuda_stream)
    bank.join(st.cuda_stream); st.synchronize()
    bank.profile_read(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        bank.process_device(iq, block, stream=st.cuda_stream)
    bank.join(st.cuda_stream)
    e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1) / steps
    k3_ms, k3_n = bank.profile_read()
    return dict(ms_per_block=ms, k3_ms=k3_ms / max(k3_n, 1), channel_MSps=n_ch * block / (ms * 1e-3) / 1e6,
                realtime_factor=block / fs / (ms * 1e-3))


def time_wf(fs, n, fps, ov, lines, steps=5):
    dev = torch.device("cuda", 0)
    avg, every_n = fftchain_params(fs, n, ov, fps)
    ns = every_n * avg * lines + n
    g = torch.Generator(device=dev); g.manual_seed(2)
    iq = 1e-3 * torch.randn(ns, 2, device=dev, generator=g)
    wf = Waterfall(fs, n, ov, fps, "adpcm")
    wf.set_pipelined(True)
    out = torch.empty(lines * wf.line_bytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(); torch.cuda.synchronize()
    for _ in range(2):
        wf.process_device(iq, ns, out, out.numel(), stream=st.cuda_stream)
    wf.join(st.cuda_stream); st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        got = wf.process_device(iq, ns, out, out.numel(), stream=st.cuda_stream)
    wf.join(st.cuda_stream)
    e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1) / steps
    unique = 8.0 * ((avg - 1) * every_n + n) + wf.line_bytes
    return dict(ms_per_batch=ms, lines_per_s=got / (ms * 1e-3), realtime_factor=got / (ms * 1e-3) / fps,
                hbm_GBps=unique * got / (ms * 1e-3) / 1e9, avg=avg, every_n=every_n)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c5", "c4"]
    if "c3" in which:
        print(json.dumps({"C3 per-GPU share: 61.44 MS/s, 128 ch x 12 kHz (D=5120, T=136533)": time_bank(61.44e6, 12000, 128, "mix", 1 << 24)}))
    if "c5" in which:
        print(json.dumps({"C5: 20 MS/s, 128 WFM ch (250 kHz IF -> 48 kHz)": time_bank(20e6, 250000, 128, "wfm", 1 << 22, audio_rate=48000.0, tau=50e-6)}))
    if "c4" in which:
        print(json.dumps({"C4: 61.44 MS/s 65536-pt 30 fps waterfall": time_wf(61.44e6, 65536, 30, 0.3, int(os.environ.get("C4_LINES", "128")))}))
