#!/usr/bin/env python
"""Stage times of the C2 device path with the three-stream pipeline on and off (each stage's own CUDA-event bracket):
off = every stage alone on the GPU, on = what the stages cost while they overlap.  One JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                             # noqa: E402
from openwebrx_b200 import ChannelBank                                   # noqa: E402
from openwebrx_b200.synth import BANDPASS                                # noqa: E402


def run(pipelined, iq, cars, steps=50):
    bank = ChannelBank(bench.FS)
    for c in cars:
        bank.add_channel(bench.OUT_RATE, demod=c["kind"], offset=c["offset"], bandpass=BANDPASS[c["kind"]])
    st = torch.cuda.Stream()
    bank.set_pipelined(pipelined)
    for _ in range(3):
        bank.process_device(iq, bench.BLOCK, stream=st.cuda_stream)
    bank.join(st.cuda_stream); st.synchronize()
    bank.profile(True); bank.profile_read_ex(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        bank.process_device(iq, bench.BLOCK, stream=st.cuda_stream)
    bank.join(st.cuda_stream)
    e1.record(st); st.synchronize()
    prof = {k: round(v[0] / max(v[1], 1), 4) for k, v in bank.profile_read_ex().items() if v[1]}
    bank.profile(False)
    bank.close()
    return dict(pipelined=pipelined, ms_per_block=round(e0.elapsed_time(e1) / steps, 4), stages_ms=prof)


if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    cars = bench.channel_plan(0, bench.CH_PER_GPU)
    iq = bench.synth_iq_torch(bench.BLOCK, bench.FS, cars, dev)
    print(json.dumps([run(False, iq, cars), run(True, iq, cars)]))
