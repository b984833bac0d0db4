#!/usr/bin/env python
"""Per-kernel device timing of Shift + FirDecimate in both evaluations (direct K3 / fast-convolution K3F) on the
BASELINE shapes, and the difference between their outputs on the same block.  One JSON object per shape."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwebrx_b200 import ChannelBank                                  # noqa: E402
from openwebrx_b200 import _native as N                                 # noqa: E402
from openwebrx_b200.synth import BANDPASS, carrier_plan                 # noqa: E402


def run(fs, out_rate, n_ch, block, mode, steps=20, wfm=False, **kw):
    steps = int(os.environ.get("FC_PROBE_STEPS", steps))                # short runs under ncu
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    iq = 1e-3 * torch.randn(block, 2, device=dev, generator=g)
    t = torch.arange(block, device=dev, dtype=torch.float32)
    iq[:, 0] += 0.2 * torch.cos(0.3 * t); iq[:, 1] += 0.2 * torch.sin(0.3 * t)
    del t
    cars = carrier_plan(n_ch, fs, seed=3, wfm=wfm)
    bank = ChannelBank(fs, outputs=N.OUT_IF | N.OUT_AUDIO)
    bank.set_fir_mode(mode)
    chans = [bank.add_channel(out_rate, demod=c["kind"], offset=c["offset"], bandpass=BANDPASS[c["kind"]], **kw) for c in cars]
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    bank.set_pipelined(True)
    for _ in range(2):
        bank.process_device(iq, block, stream=st.cuda_stream)
    bank.join(st.cuda_stream); st.synchronize()
    bank.profile(True)
    bank.profile_read_ex(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        bank.process_device(iq, block, stream=st.cuda_stream)
    bank.join(st.cuda_stream)
    e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1) / steps
    prof = {k: (v[0] / steps, v[1] // steps) for k, v in bank.profile_read_ex().items() if v[1]}   # ms per block, brackets per block
    bank.profile(False)
    bank.drain()
    outs = [chans[i].read_if() for i in (0, n_ch // 2, n_ch - 1)]
    res = dict(mode=mode, ms_per_block=round(ms, 4), channel_MSps=round(n_ch * block / (ms * 1e-3) / 1e6, 1),
               stages_ms_per_block={k: round(v[0], 4) for k, v in prof.items()}, brackets_per_block={k: v[1] for k, v in prof.items()})
    bank.close()
    return res, outs


def main():
    shapes = {
        "C2": dict(fs=10e6, out_rate=12000, n_ch=64, block=1 << 24),
        "C3/8": dict(fs=61.44e6, out_rate=12000, n_ch=128, block=1 << 24),
        "C3/8L": dict(fs=61.44e6, out_rate=12000, n_ch=128, block=1 << 25, steps=8),
        "C3": dict(fs=61.44e6, out_rate=12000, n_ch=1024, block=1 << 25, steps=3),
        "C5": dict(fs=20e6, out_rate=250000, n_ch=128, block=1 << 23, wfm=True, audio_rate=48000.0, tau=50e-6),
    }
    pick = [a for a in sys.argv[1:] if a in shapes] or list(shapes)
    only = [a for a in sys.argv[1:] if a in ("direct", "fastconv", "fastconv_tc")]
    for name in pick:
        if only:
            print(json.dumps({name: run(mode=only[0], **shapes[name])[0]}), flush=True)
            continue
        def rel(oa, ob):
            diff = []
            for x, y in zip(oa, ob):
                n = min(len(x), len(y))
                den = float(np.sqrt(np.mean(np.abs(x[:n]) ** 2)))
                diff.append(float(np.sqrt(np.mean(np.abs(x[:n] - y[:n]) ** 2))) / den if den else 0.0)
            return diff

        a, oa = run(mode="direct", **shapes[name])
        b, ob = run(mode="fastconv", **shapes[name])
        c, oc = run(mode="fastconv_tc", **shapes[name])
        print(json.dumps({name: dict(direct=a, fastconv=b, fastconv_tc=c, if_rel_rms_direct_vs_fastconv=rel(oa, ob),
                                     if_rel_rms_direct_vs_tc=rel(oa, oc), if_rel_rms_fastconv_vs_tc=rel(ob, oc))}), flush=True)


if __name__ == "__main__":
    main()
