"""Control path under load (csdr/chain/selector.py:132-166, owrx/dsp.py:96-148,835-839): websocket threads
call setFrequencyOffset / setBandpass / setSquelchLevel / setDemodulator and clients come and go while the DSP thread streams
blocks.  Nothing a client does may stall the device or change anybody's samples except from the next block boundary on.

Live run: a feeder thread issues pipelined device-resident blocks without waiting for them, a second thread fires the events.
Every event is stamped with the block it preceded (one Python lock around "event + stamp" and around "issue + count").
Quiesced run: one thread applies the same events before the same blocks with a device synchronisation after every call.
The drained audio of both runs must be bit-identical, and the live run may not be slower than an event-free run by more than
10 %."""
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FS = 10_000_000
OUT = 12000
BASE = 56            # base clients; with the 8 dynamic ones the group never outgrows its 64 slots (growing is the one
DYN = 8              # control operation documented to quiesce the device)
KINDS = ("nfm", "am", "usb")


def _events(n, seed, other_group=False):
    """deterministic event list; channels are addressed by logical index (0..BASE-1 fixed, BASE..BASE+DYN-1 come and go)"""
    rng = np.random.default_rng(seed)
    ev = []
    for _ in range(n):
        r = rng.random()
        tgt = int(rng.integers(0, BASE + DYN))
        if r < 0.40:
            lo = float(rng.uniform(-5900, -100)); hi = float(rng.uniform(100, 5900))
            ev.append(("bandpass", tgt, lo, hi))
        elif r < 0.65:
            ev.append(("offset", tgt, int(rng.integers(-4_000_000, 4_000_000))))
        elif r < 0.75:
            ev.append(("squelch", tgt, float(rng.uniform(-90.0, -20.0))))
        elif r < 0.85:
            ev.append(("demod", tgt, KINDS[int(rng.integers(0, 3))]))
        elif r < 0.90:
            ev.append(("format", tgt, ("f32", "s16", "adpcm")[int(rng.integers(0, 3))]))
        else:
            rate = OUT
            if other_group and rng.random() < 0.5:
                rate = 8000
            ev.append(("toggle", BASE + int(rng.integers(0, DYN)), rate, KINDS[int(rng.integers(0, 3))],
                       int(rng.integers(-4_000_000, 4_000_000))))
    return ev


class _Rig:
    def __init__(self, torch, cars):
        from openwebrx_b200 import ChannelBank
        from openwebrx_b200.synth import BANDPASS
        self.torch = torch
        self.bank = ChannelBank(FS)
        self.ch = {}
        for i, c in enumerate(cars[:BASE]):
            self.ch[i] = self.bank.add_channel(OUT, demod=c["kind"], offset=c["offset"], bandpass=BANDPASS[c["kind"]])
            # the client audio tail (Convert / AdpcmEncoder) runs from the start: format events then change who uses it, not
            # whether the block carries that stage (the timing comparison below is about stalls, not about added work)
            self.ch[i].setAudioFormat(("adpcm", "s16", "f32")[i % 3])
        self.ids_seen = []

    def apply(self, e):
        from openwebrx_b200.synth import BANDPASS
        kind, tgt = e[0], e[1]
        if kind == "toggle":
            if tgt in self.ch:
                self.ch.pop(tgt).remove()
            else:
                self.ch[tgt] = self.bank.add_channel(e[2], demod=e[3], offset=e[4], bandpass=BANDPASS[e[3]])
                self.ids_seen.append(self.ch[tgt].id)
            return
        ch = self.ch.get(tgt)
        if ch is None:
            return
        if kind == "bandpass":
            ch.setBandpass(e[2], e[3])
        elif kind == "offset":
            ch.setFrequencyOffset(e[2])
        elif kind == "squelch":
            ch.setSquelchLevel(e[2])
        elif kind == "demod":
            ch.setDemodulator(e[2])
        elif kind == "format":
            ch.setAudioFormat(e[2])

    def collect(self):
        """drain the last block: {logical index: (audio float32, tail bytes)}"""
        self.bank.drain()
        return {k: (c.read_audio(), c.read_bytes()) for k, c in sorted(self.ch.items())}


def _live(torch, cars, blocks, n, n_blocks, events, drain_every, fire=True):
    rig = _Rig(torch, cars)
    st = torch.cuda.Stream()
    rig.bank.set_pipelined(True)
    lock = threading.Lock()
    tick = threading.Condition()
    state = dict(issued=0, fired=0)
    stamps = []

    per_block = max(1, len(events) // max(n_blocks, 1))

    def fire_all():
        for k, e in enumerate(events):
            # paced: `per_block` events per block of the stream (a Python lock is not fair: an unpaced loop would starve the feeder)
            with tick:
                while state["issued"] < k // per_block:
                    tick.wait(0.05)                    # (no spinning: a busy Python thread would take the GIL from the feeder)
            with lock:
                rig.apply(e)
                stamps.append(state["issued"])
                state["fired"] += 1

    # warm-up outside the measurement (allocations, table builds)
    for i in range(3):
        rig.bank.process_device(blocks[i & 1], n, stream=st.cuda_stream)
    rig.bank.join(st.cuda_stream); st.synchronize()
    rig.collect()
    th = threading.Thread(target=fire_all) if fire else None
    out = {}
    marks = []
    t0 = time.perf_counter()
    if th:
        th.start()
    b = 0
    while b < n_blocks or (fire and state["fired"] < len(events)):
        with lock:
            rig.bank.process_device(blocks[b & 1], n, stream=st.cuda_stream)
            state["issued"] += 1
            if (b + 1) % drain_every == 0:
                out[b] = rig.collect()                     # same lock hold: no event between a block and its drain
        with tick:
            tick.notify_all()
        ev = torch.cuda.Event(); ev.record(st); marks.append(ev)
        if len(marks) > 4:
            marks[-5].synchronize()                        # run at most four blocks ahead of the device, as a paced source would
        b += 1
        if b > 20 * n_blocks:
            raise AssertionError("event thread starved")
    rig.bank.join(st.cuda_stream); st.synchronize()
    wall = time.perf_counter() - t0
    if th:
        th.join()
    ids = list(rig.ids_seen)
    n_chans = len(rig.bank.channels)
    rig.bank.close()
    return out, stamps, b, wall, ids, n_chans


def _quiesced(torch, cars, blocks, n, n_blocks, events, stamps, drain_every):
    rig = _Rig(torch, cars)
    st = torch.cuda.Stream()
    rig.bank.set_pipelined(True)
    for i in range(3):
        rig.bank.process_device(blocks[i & 1], n, stream=st.cuda_stream)
    rig.bank.join(st.cuda_stream); st.synchronize()
    rig.collect()
    out = {}
    k = 0
    for b in range(n_blocks):
        while k < len(events) and stamps[k] <= b:
            rig.apply(events[k]); k += 1
            torch.cuda.synchronize()
        rig.bank.process_device(blocks[b & 1], n, stream=st.cuda_stream)
        rig.bank.join(st.cuda_stream)
        torch.cuda.synchronize()
        if (b + 1) % drain_every == 0:
            out[b] = rig.collect()
    rig.bank.close()
    return out


def _compare(a, b):
    assert sorted(a) == sorted(b)
    checked = 0
    for blk in sorted(a):
        assert sorted(a[blk]) == sorted(b[blk]), f"block {blk}: different live channel sets"
        for k in a[blk]:
            au_a, by_a = a[blk][k]; au_b, by_b = b[blk][k]
            assert au_a.shape == au_b.shape and by_a.shape == by_b.shape, f"block {blk} channel {k}: counts differ"
            assert np.array_equal(au_a.view(np.uint32), au_b.view(np.uint32)), f"block {blk} channel {k}: audio differs"
            assert np.array_equal(by_a, by_b), f"block {blk} channel {k}: tail bytes differ"
            checked += 1
    return checked


@pytest.fixture(scope="module")
def rig_input(gpu):
    import torch
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cars = bench.channel_plan(0, 64, fs=FS)
    n = 1 << 23                                            # 0.84 s of signal per block (1.7 ms of device time with the audio tails)
    iq = bench.synth_iq_torch(n, FS, cars, dev)
    return torch, cars, [iq, iq.flip(0).contiguous()], n


def test_1000_events_over_200_blocks_bit_exact_and_no_stall(rig_input):
    torch, cars, blocks, n = rig_input
    events = _events(1000, 7)
    n_blocks, every = 200, 5
    live, stamps, issued, wall_live, ids, _ = _live(torch, cars, blocks, n, n_blocks, events, every)
    assert len(stamps) == 1000
    spread = len(set(stamps))
    assert spread >= 20, f"events landed on only {spread} distinct block boundaries: the run did not interleave"
    quiet = _quiesced(torch, cars, blocks, n, issued, events, stamps, every)
    checked = _compare(live, quiet)
    # channel ids are handles: the lowest free one is reused, the table does not grow with the number of clients ever seen
    assert ids and max(ids) < BASE + DYN, f"ids grew to {max(ids)}"
    # same run without the event thread: per-block time may not grow by more than 10 %
    base_runs, live_runs = [], [wall_live / issued]
    for _ in range(3):
        _, _, nb, wall, _, _ = _live(torch, cars, blocks, n, n_blocks, [], every, fire=False)
        base_runs.append(wall / nb)
        _, _, nb, wall, _, _ = _live(torch, cars, blocks, n, n_blocks, events, every)
        live_runs.append(wall / nb)
    ratio = min(live_runs) / min(base_runs)
    print(f"\ncontrol churn: {checked} channel-blocks bit-identical; events on {spread} block boundaries; "
          f"per-block wall {1e3 * min(live_runs):.3f} ms with 1000 events vs {1e3 * min(base_runs):.3f} ms without (x{ratio:.3f})")
    assert ratio < 1.10, f"control events slowed the stream down by x{ratio:.3f}"


def test_groups_come_and_go_without_quiescing(rig_input):
    """clients of a second class (8 kHz output: another decimation, another lock-step group) join and leave: the group is
    created, retired when its last client leaves and its memory released later, all while blocks are in flight"""
    torch, cars, blocks, n = rig_input
    events = _events(300, 11, other_group=True)
    n_blocks, every = 60, 3
    live, stamps, issued, _, ids, _ = _live(torch, cars, blocks, n, n_blocks, events, every)
    quiet = _quiesced(torch, cars, blocks, n, issued, events, stamps, every)
    checked = _compare(live, quiet)
    assert checked > 0 and max(ids) < BASE + DYN
