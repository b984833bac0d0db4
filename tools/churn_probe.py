#!/usr/bin/env python
"""Where the time goes when control events fire during a pipelined stream: per-event-kind time inside the library call,
time inside owrx_bank_process_device, and the wall clock of the run (same rig as tests/test_gpu_control_churn.py)."""
import collections
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                                                   # noqa: E402
import bench                                                   # noqa: E402
import test_gpu_control_churn as T                             # noqa: E402


def main(log2n=22, n_blocks=200, n_events=1000):
    dev = torch.device("cuda", 0); torch.cuda.set_device(0)
    cars = bench.channel_plan(0, 64, fs=T.FS)
    n = 1 << log2n
    iq = bench.synth_iq_torch(n, T.FS, cars, dev)
    blocks = [iq, iq.flip(0).contiguous()]
    events = T._events(n_events, 7)
    all_events = events
    for mode in ("none", "thread", "inline", "bandpass", "offset", "squelch", "demod", "format", "toggle"):
        events = all_events if mode in ("none", "thread", "inline") else [e if e[0] == mode else ("noop", 0) for e in all_events]
        rig = T._Rig(torch, cars)
        st = torch.cuda.Stream()
        rig.bank.set_pipelined(True)
        for i in range(3):
            rig.bank.process_device(blocks[i & 1], n, stream=st.cuda_stream)
        rig.bank.join(st.cuda_stream); st.synchronize()
        lock = threading.Lock()
        state = dict(issued=0)
        ev_t = collections.defaultdict(float); ev_n = collections.Counter()
        per_block = max(1, n_events // n_blocks)

        def fire(k):
            e = events[k]
            t0 = time.perf_counter()
            rig.apply(e)
            ev_t[e[0]] += time.perf_counter() - t0; ev_n[e[0]] += 1

        def fire_all():
            for k in range(n_events):
                while state["issued"] < k // per_block:
                    time.sleep(0)
                with lock:
                    fire(k)

        th = threading.Thread(target=fire_all) if mode == "thread" else None
        t_pd = 0.0
        marks = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rig.bank.profile(True); rig.bank.profile_read_ex(reset=True)
        e0.record(st)
        w0 = time.perf_counter()
        if th:
            th.start()
        k = 0
        for b in range(n_blocks):
            if mode not in ("none", "thread"):
                while k < n_events and k // per_block <= b:
                    fire(k); k += 1
            with lock:
                t0 = time.perf_counter()
                rig.bank.process_device(blocks[b & 1], n, stream=st.cuda_stream)
                t_pd += time.perf_counter() - t0
                state["issued"] += 1
            ev = torch.cuda.Event(); ev.record(st); marks.append(ev)
            if len(marks) > 4:
                marks[-5].synchronize()
        rig.bank.join(st.cuda_stream); e1.record(st); st.synchronize()
        wall = time.perf_counter() - w0
        if th:
            th.join()
        prof = {k: round(v[0] / max(v[1], 1), 4) for k, v in rig.bank.profile_read_ex().items() if v[1]}
        print(json.dumps(dict(mode=mode, stages_ms=prof, log2n=log2n, wall_ms_per_block=1e3 * wall / n_blocks, gpu_ms_per_block=e0.elapsed_time(e1) / n_blocks,
                              process_device_ms_per_call=1e3 * t_pd / n_blocks,
                              event_ms={k: round(1e3 * v / max(ev_n[k], 1), 4) for k, v in ev_t.items()}, event_n=dict(ev_n))))
        rig.bank.close()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 22)
