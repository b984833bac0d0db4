#!/usr/bin/env python
"""Real-time sustain check: one GPU's share of the north-star load, paced at the true sample rate.

61.44 MS/s complex-float IQ arrives in 1/30 s blocks from pinned host memory; every block goes through the public
host API of BOTH halves of the hot path: 128 concurrent 12 kHz client channels (NFM/AM/USB, ADPCM-less float audio
read back per channel) and the 65536-point, 30 fps waterfall (ADPCM lines read back).  Reports how long a block
takes against its 33.3 ms budget.  One JSON object."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwebrx_b200 import ChannelBank, Waterfall                       # noqa: E402
from openwebrx_b200.synth import BANDPASS, carrier_plan                 # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--fs", type=float, default=61.44e6)
    ap.add_argument("--fps", type=int, default=30)
    args = ap.parse_args()
    fs, fps = args.fs, args.fps
    block = int(fs / fps)
    n_blocks = int(args.seconds * fps)
    # a ring of 8 distinct pinned blocks of noise + a few carriers (the content does not change the cost)
    ring = []
    g = torch.Generator(); g.manual_seed(1)
    t = torch.arange(block, dtype=torch.float32)
    for k in range(8):
        x = 1e-3 * torch.randn(block, 2, generator=g)
        x[:, 0] += 0.1 * torch.cos(0.01 * (k + 1) * t); x[:, 1] += 0.1 * torch.sin(0.01 * (k + 1) * t)
        ring.append(x.pin_memory())
    cars = carrier_plan(args.channels, fs, seed=7)
    bank = ChannelBank(fs)
    chans = [bank.add_channel(12000, demod=c["kind"], offset=c["offset"], bandpass=BANDPASS[c["kind"]]) for c in cars]
    wf = Waterfall(fs, 65536, 0.3, fps, "adpcm")
    audio = np.empty(1 << 14, np.float32)

    def one(k):
        x = ring[k % len(ring)]
        bank.feed_ptr(x.data_ptr(), block)
        got = sum(c.read_audio_into(audio) for c in chans)
        lines = wf.feed(x.numpy().view(np.complex64).reshape(-1))
        return got, len(lines)

    for k in range(4):
        one(k)                                               # tables, scratch, first-use allocations
    busy, n_audio, n_lines, late = [], 0, 0, 0
    period = 1.0 / fps
    t0 = time.perf_counter()
    for k in range(n_blocks):
        deadline = t0 + k * period
        now = time.perf_counter()
        if now < deadline:
            time.sleep(deadline - now)
        elif now - deadline > period:
            late += 1
        s = time.perf_counter()
        a, l = one(k)
        busy.append(time.perf_counter() - s)
        n_audio += a; n_lines += l
    wall = time.perf_counter() - t0
    busy = np.asarray(busy) * 1e3
    print(json.dumps({
        "workload": "%d x 12 kHz channels + 65536-pt %d fps waterfall from %.2f MS/s, %d blocks of 1/%d s, host API (H2D + D2H inside)"
                    % (args.channels, fps, fs / 1e6, n_blocks, fps),
        "budget_ms_per_block": period * 1e3, "block_ms_mean": float(busy.mean()), "block_ms_p99": float(np.percentile(busy, 99)),
        "block_ms_max": float(busy.max()), "headroom_x": float(period * 1e3 / busy.mean()), "late_blocks": late,
        "audio_samples_per_channel_per_s": n_audio / args.channels / wall, "waterfall_lines_per_s": n_lines / wall, "wall_s": wall}))


if __name__ == "__main__":
    main()
