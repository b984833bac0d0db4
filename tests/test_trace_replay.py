"""CPU half of the reference call-trace replay (VERDICT r1 task 1b): the recorded wiring of the reference's unmodified
SpectrumThread (owrx/fft.py:13-109) and DspManager -> ClientDemodulatorChain (owrx/dsp.py:39-425,437-937) is replayed on the
pycsdr shim WITHOUT data, and at every recorded checkpoint the chains the shim would fuse are compared with what the reference
configured.  The GPU half (test_gpu_trace_replay.py) replays the same fixtures with IQ flowing."""
import pycsdr.modules as M
import trace_replay
from pycsdr.types import AgcProfile


def _heads(rp, cls, buffer):
    out = []
    for o in rp.obj.values():
        if isinstance(o, cls) and o._reader is not None and o._reader._buffer is buffer and not o._stopped and not o._reader._stopped:
            out.append(o)
    return out


def test_spectrum_trace_wiring():
    tr = trace_replay.load("trace_spectrum.json")
    rp = trace_replay.Replay(tr)
    seen = []
    try:
        while True:
            mk = rp.next_mark()
            if mk is None:
                break
            seen.append(mk["name"])
            heads = _heads(rp, M.Fft, rp.source())
            if mk["name"] == "stopped":
                assert heads == []
                continue
            assert len(heads) == 1
            chain = M._WaterfallPlan.match(M._walk(heads[0])[0])
            assert chain is not None, mk["name"]
            assert chain[0].size == mk["n"] and chain[0].every_n_samples == mk["every_n"] and chain[1].avg_number == mk["avg"]
            assert (len(chain) == 4) == (mk["compression"] == "adpcm")
            # the output Reader the reference pumps from sits on the Buffer the fused chain writes
            assert rp.obj[mk["output"]]._buffer is chain[-1]._writer
    finally:
        rp.close()
    assert seen == ["start", "fps30", "uncompressed", "size1024", "stopped"]


def test_client_trace_wiring():
    tr = trace_replay.load("trace_client.json")
    rp = trace_replay.Replay(tr)
    seen = []
    try:
        while True:
            mk = rp.next_mark()
            if mk is None:
                break
            seen.append(mk["name"])
            heads = _heads(rp, M.Shift, rp.source())
            if mk["name"] == "stopped":
                assert heads == []
                continue
            assert len(heads) == 1
            d = M._ChannelPlan.match(M._walk(heads[0])[0])
            assert d is not None, mk["name"]
            fs, out = mk["fs"], mk["out_rate"]
            assert d["demod"] == {"usb": "ssb"}.get(mk["demod"], mk["demod"])
            assert d["audio_fmt"] == "adpcm"
            assert abs(d["shift"].rate - (-mk["offset"] / fs)) < 1e-15
            assert d["fir"].decimation == int(fs / out)
            assert (d["frac"] is not None) == (fs / int(fs / out) != out)
            lo, hi = mk["bandpass"]
            assert abs(d["bandpass"].low - lo / out) < 1e-12 and abs(d["bandpass"].high - hi / out) < 1e-12
            assert abs(d["squelch"].level - 10.0 ** (mk["squelch_db"] / 10.0)) < 1e-18
            if mk["demod"] == "usb":
                assert d["agc"].profile is AgcProfile.FAST
            if mk["demod"] == "am":
                assert d["agc"].initialGain == 200
            if mk["demod"] == "nfm":
                assert d["agc"].maxGain == 3
            if mk["demod"] == "wfm":
                assert abs(d["wfm_frac"].rate - 250000.0 / mk["audio_rate"]) < 1e-12 and d["wfm_de"].tau == mk["tau"]
            # the audio Reader the reference pumps from sits on the Buffer the fused chain's tail writes
            audio = mk["readers"]["hd_audio" if mk.get("hd") else "audio"]
            assert rp.obj[audio]._buffer is d["tail"]._writer
            assert d["squelch"].powerWriter is rp.obj[mk["readers"]["smeter"]]._buffer
            if "secondary_fft" in mk:
                # both second-level heads read the selectorBuffer the client chain's Squelch writes (owrx/dsp.py:49,220-225)
                sel_buf = d["stages"][d["if_index"]]._writer
                ffts = _heads(rp, M.Fft, sel_buf)
                assert len(ffts) == 1
                c2 = M._WaterfallPlan.match(M._walk(ffts[0])[0])
                sf = mk["secondary_fft"]
                assert c2 and c2[0].size == sf["n"] and c2[0].every_n_samples == sf["every_n"] and c2[1].avg_number == sf["avg"] and len(c2) == 4
                assert c2[-1]._writer is rp.obj[mk["readers"]["secondary_fft"]]._buffer
                shifts = _heads(rp, M.Shift, sel_buf)
                assert len(shifts) == 1
                d2 = M._ChannelPlan.match(M._walk(shifts[0])[0])
                ss = mk["secondary_selector"]
                assert d2 and d2["demod"] == "none" and d2["fir"] is None and abs(d2["shift"].rate + ss["offset"] / out) < 1e-15
                assert abs(d2["bandpass"].high - ss["bandwidth"] / out) < 1e-12
                assert rp.obj[ss["output"]]._buffer is d2["tail"]._writer
    finally:
        rp.close()
    assert seen == ["nfm_start", "nfm_retuned", "nfm_squelched", "am", "usb", "usb_secondary", "wfm", "nfm_again", "stopped"]
