"""Split-phase drain of the device path (owrx_bank_drain_begin / _end): block i's outputs are copied to the host while block
i + 1 is already being processed; queues must receive exactly what the synchronous owrx_bank_drain delivers."""
import numpy as np
import pytest

from openwebrx_b200 import ChannelBank
from openwebrx_b200 import _native as N
from openwebrx_b200.synth import BANDPASS, carrier_plan, make_iq

pytestmark = pytest.mark.gpu


def _blocks(fs, cars, n_blocks, n):
    # one continuous stream cut into [carry | new] blocks the way a streaming host presents them
    return make_iq(n * n_blocks + 4096, fs, cars, seed=11)


@pytest.mark.parametrize("pipelined", [False, True])
def test_split_phase_drain_equals_synchronous_drain(gpu, pipelined):
    import torch
    fs, out, n, n_blocks = 2.4e6, 12000, 1 << 19, 5
    cars = carrier_plan(6, fs, seed=9)
    iq = _blocks(fs, cars, n_blocks, n)

    def run(split):
        bank = ChannelBank(fs, outputs=N.OUT_AUDIO | N.OUT_DEMOD | N.OUT_IF)
        chans = [bank.add_channel(out, demod=c["kind"], offset=c["offset"], bandpass=BANDPASS[c["kind"]]) for c in cars]
        bank.set_pipelined(pipelined)
        st = torch.cuda.Stream()
        got = [[[], [], []] for _ in chans]

        def pop():
            for i, ch in enumerate(chans):
                got[i][0].append(ch.read_audio()); got[i][1].append(ch.read_demod()); got[i][2].append(ch.read_if())

        pos = 0
        keep = []                                   # device blocks stay alive until everything issued on `st` has run
        for b in range(n_blocks):
            blk = torch.from_numpy(iq[pos:pos + n].view(np.float32)).cuda()
            keep.append(blk)
            bank.process_device(blk, n, stream=st.cuda_stream)
            pos += bank.last_consumed()
            if split:
                bank.drain_end()                    # the previous block (no-op for the first)
                if b:
                    pop()
                bank.drain_begin()                  # this block: copied while the next one is processed
            else:
                bank.drain()
                pop()
        if split:
            bank.drain_end()
            pop()
        st.synchronize()
        bank.close()
        return [[np.concatenate(x) for x in ch] for ch in got]

    want, have = run(False), run(True)
    for w, h in zip(want, have):
        for a, b in zip(w, h):
            assert len(a) == len(b) and len(a) > 0
            assert np.array_equal(a, b)


def test_drain_begin_twice_is_refused(gpu):
    import torch
    fs = 2.4e6
    cars = carrier_plan(2, fs, seed=9)
    bank = ChannelBank(fs, outputs=N.OUT_AUDIO)
    for c in cars:
        bank.add_channel(12000, demod=c["kind"], offset=c["offset"], bandpass=BANDPASS[c["kind"]])
    iq = torch.from_numpy(make_iq(1 << 18, fs, cars, seed=3).view(np.float32)).cuda()
    bank.process_device(iq, 1 << 18)
    bank.drain_begin()
    with pytest.raises(Exception):
        bank.drain_begin()
    bank.drain_end()
    bank.drain_end()                                # no-op
    bank.close()
