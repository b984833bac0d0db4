"""Replays a recorded pycsdr call trace (tests/golden/trace_*.json, made by tests/golden/make_trace.py from the reference's
UNMODIFIED SpectrumThread / DspManager) on the repo's pycsdr shim: the same constructors, setReader / setWriter / stop /
set* calls in the same order, on fresh objects.  Test infrastructure only."""
import json
import os
import threading
import time

import pycsdr.modules as M
from pycsdr.types import AgcProfile, Format

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


class Sink:
    """what the reference's pump threads do with an output Reader (csdr/module/__init__.py:36-53): one read() = one message"""

    def __init__(self, reader):
        self.reader = reader
        self.msgs = []
        self.lock = threading.Lock()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while True:
            d = self.reader.read()
            if d is None:
                break
            with self.lock:
                self.msgs.append(bytes(d))

    def take(self):
        with self.lock:
            out, self.msgs = self.msgs, []
        return out


class Replay:
    def __init__(self, trace):
        self.trace = trace
        self.events = trace["events"]
        self.pos = 0
        self.obj = {trace["source"]: M.Buffer(Format.COMPLEX_FLOAT)}
        self.sinks = {}
        # objects the recording never touches again are dropped at once, as the reference's own garbage is: e.g. the Reader
        # Chain.replace hands to a DummyDemodulator (owrx/dsp.py:96-103) must not stay alive as a cursor nobody advances
        self.used = set()

        def scan(v):
            if isinstance(v, dict):
                if "ref" in v:
                    self.used.add(v["ref"])
                for x in v.values():
                    scan(x)
            elif isinstance(v, list):
                for x in v:
                    scan(x)
            elif isinstance(v, str):
                self.used.add(v)
        for e in self.events:
            if e["op"] == "mark":
                scan({k: v for k, v in e.items() if k not in ("op", "name")})
            else:
                self.used.add(e["id"])
                scan(e["args"]); scan(e["kwargs"])

    # ---------------------------------------------------------------- decoding
    def _dec(self, v):
        if isinstance(v, dict):
            if "format" in v:
                return Format[v["format"]]
            if "agc" in v:
                return AgcProfile[v["agc"]]
            if "ref" in v:
                return self.obj[v["ref"]]
            raise AssertionError("the trace carries an argument the replay cannot rebuild: %r" % (v,))
        return v

    def source(self):
        return self.obj[self.trace["source"]]

    def next_mark(self):
        """executes the recorded calls up to the next mark and returns it (None at the end of the trace)"""
        while self.pos < len(self.events):
            e = self.events[self.pos]
            self.pos += 1
            if e["op"] == "mark":
                return e
            args = [self._dec(a) for a in e["args"]]
            kwargs = {k: self._dec(v) for k, v in e["kwargs"].items()}
            if e["op"] == "new":
                self.obj[e["id"]] = getattr(M, e["cls"])(*args, **kwargs)
                continue
            target = self.obj[e["id"]]
            try:
                ret = getattr(target, e["method"])(*args, **kwargs)
            except Exception as ex:
                assert e.get("raises") == type(ex).__name__, "replay of %s.%s raised %r, the recording %s" % (
                    e["id"], e["method"], ex, e.get("raises"))
                continue
            assert e.get("raises") is None, "the recording raised %s at %s.%s, the replay did not" % (e["raises"], e["id"], e["method"])
            if "ret" in e and e["ret"] in self.used:
                self.obj[e["ret"]] = ret
        return None

    # ---------------------------------------------------------------- data
    def sink(self, reader_id):
        if reader_id not in self.sinks:
            self.sinks[reader_id] = Sink(self.obj[reader_id])
        return self.sinks[reader_id]

    def feed(self, iq, chunk=150000, timeout=120.0):
        """writes complex64 samples into the source Buffer like a TcpSource would (chunked) and waits until the shim's runners —
        the source's and those of every Buffer downstream that has its own (selectorBuffer readers) — have worked them off"""
        src = self.source()
        raw = iq.tobytes()
        for o in range(0, len(raw), 8 * chunk):
            src.write(raw[o:o + 8 * chunk])
        self.settle(timeout)

    def settle(self, timeout=120.0):
        t0 = time.time()
        stable = 0
        while time.time() - t0 < timeout:
            busy = False
            for b in [o for o in self.obj.values() if isinstance(o, M.Buffer)]:
                r = b._runner
                if r is not None and r.is_alive() and r.processed < b._end:
                    busy = True
            stable = 0 if busy else stable + 1
            if stable >= 3:                 # a runner that finished may just have written into another runner's Buffer
                return
            time.sleep(0.01)
        raise AssertionError("the shim's runners did not settle within %.0f s" % timeout)

    def close(self):
        for s in self.sinks.values():
            s.reader.stop()
        for o in self.obj.values():
            if isinstance(o, M.Module):
                try:
                    o.stop()
                except Exception:
                    pass
