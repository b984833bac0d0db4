"""GPU half of the reference call-trace replay (VERDICT r1 task 1b).

tests/golden/trace_spectrum.json and trace_client.json hold the exact pycsdr call sequences of the reference's UNMODIFIED
SpectrumThread (owrx/fft.py:13-109) and DspManager -> ClientDemodulatorChain (owrx/dsp.py:39-425,437-937), recorded in the build
container (tests/golden/make_trace.py).  Here they are replayed on the shim with IQ flowing before and after every recorded
change — fps / compression / fft_size for the waterfall; retune, band-pass, squelch, NFm -> Am -> Ssb -> WFm -> NFm demodulator
swaps, the secondary FFT and a SecondarySelector on the shared selectorBuffer for the client — and what arrives at the
Readers the reference pumps from is compared with the oracle."""
import json
import os

import numpy as np
import pytest

import oracle
import trace_replay
from openwebrx_b200.synth import make_iq
from test_gpu_waterfall import _assert_db_parity, _truth_db
from test_oracle import JsImaAdpcmCodec, browser_fft_decode

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_margins.json")


def _record(case, form, **vals):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    try:
        rep = json.load(open(REPORT))
    except Exception:
        rep = {}
    rep.setdefault(case, {})[form] = vals
    json.dump(rep, open(REPORT, "w"), indent=1, sort_keys=True)
    print("[replay] %s / %s: %s" % (case, form, json.dumps(vals)))


CARRIERS = [dict(offset=250000, amp=0.20, kind="nfm"), dict(offset=-321000, amp=0.08, kind="amfm"),
            dict(offset=600000, amp=0.15, kind="usb"), dict(offset=100000, amp=0.25, kind="nfm")]


def test_spectrum_thread_trace_with_data(gpu):
    fs = 2400000
    rp = trace_replay.Replay(trace_replay.load("trace_spectrum.json"))
    hist, pos, line_start = [], 0, 0
    L = 3
    try:
        while True:
            mk = rp.next_mark()
            if mk is None or mk["name"] == "stopped":
                break
            n, avg, every_n = mk["n"], mk["avg"], mk["every_n"]
            if mk["name"] in ("start", "size1024"):
                line_start = pos                       # a new FftChain starts with the first sample that arrives after it
            end = line_start + (L * avg - 1) * every_n + n      # exactly L whole lines complete with this segment
            seg = make_iq(end - pos, fs, CARRIERS, seed=77, t0=pos)
            sink = rp.sink(mk["output"])
            hist.append(seg)
            rp.feed(seg)
            msgs = sink.take()
            all_iq = np.concatenate(hist)
            ref = oracle.fftchain_run(all_iq[line_start:end], n, every_n, avg, compression=mk["compression"])
            assert len(msgs) == L == len(ref["db"]), (mk["name"], len(msgs))
            if mk["compression"] == "adpcm":
                lb = (n + 10) // 2
                assert all(len(m) == lb for m in msgs)              # one read() = one line = one websocket message
                same = sum(m == l.tobytes() for m, l in zip(msgs, ref["lines"]))
                # what the browser shows, against the oracle's dB: the codec's own error bounds both
                e_gpu = np.sqrt(np.mean([(browser_fft_decode(np.frombuffer(m, np.uint8)) - db) ** 2 for m, db in zip(msgs, ref["db"])]))
                e_ref = np.sqrt(np.mean([(browser_fft_decode(l) - db) ** 2 for l, db in zip(ref["lines"], ref["db"])]))
                _record("SpectrumThread trace", mk["name"], lines=L, byte_identical_lines=same, shown_rms_db_err_gpu=float(e_gpu),
                        shown_rms_db_err_oracle_codec=float(e_ref))
                assert e_gpu <= 1.02 * e_ref + 0.01
            else:
                assert all(len(m) == 4 * n for m in msgs)
                db = np.stack([np.frombuffer(m, np.float32) for m in msgs])
                _assert_db_parity(db, ref["db"], _truth_db(all_iq[line_start:end], n, every_n, avg))
                _record("SpectrumThread trace", mk["name"], lines=L, max_abs_db_err=float(np.abs(db - ref["db"]).max()))
            line_start += L * avg * every_n
            pos = end
        # after SpectrumThread.stop() nothing is attached: more samples produce nothing
        rp.feed(make_iq(200000, fs, CARRIERS, seed=77, t0=pos))
        assert sink.take() == []
    finally:
        rp.close()


SPAN = 3400        # the streams may be offset by up to one squelch block of audio (750 samples; 3000 for the WFM chain) + FIR lead


def _best_lag(a, b, span=SPAN):
    """lag d maximising sum a[i + d] b[i] (|d| <= span), by FFT cross-correlation"""
    n = min(len(a), len(b)) - 2 * span
    ref = b[span:span + n].astype(np.float64)
    seg = a[:span + n + span].astype(np.float64)
    m = 1 << int(np.ceil(np.log2(len(seg) + n)))
    xc = np.fft.irfft(np.fft.rfft(seg, m) * np.conj(np.fft.rfft(ref, m)), m)[:2 * span + 1]
    return int(np.argmax(xc)) - span


def _compare_audio(got, want):
    """normalised correlation and gain ratio over the second half, after aligning the streams"""
    d = _best_lag(got, want)
    n = min(len(got), len(want)) - 2 * SPAN
    a = got[SPAN + d:SPAN + d + n].astype(np.float64)[n // 2:]
    b = want[SPAN:SPAN + n].astype(np.float64)[n // 2:]
    corr = float(np.dot(a, b) / np.sqrt(np.dot(a, a) * np.dot(b, b)))
    gain = float(np.sqrt(np.dot(a, a) / np.dot(b, b)))
    resid = float(np.sqrt(np.mean((a / gain - b) ** 2)) / np.sqrt(np.mean(b ** 2)))
    return dict(lag=d, corr=corr, gain_ratio=gain, rel_rms_after_gain=resid)


def test_dsp_manager_trace_with_data(gpu):
    fs = 2400000
    rp = trace_replay.Replay(trace_replay.load("trace_client.json"))
    KIND = {"nfm": oracle.DEMOD_NFM, "am": oracle.DEMOD_AM, "usb": oracle.DEMOD_SSB, "wfm": oracle.DEMOD_WFM}
    seg_len = 20 * 150000              # 1.25 s: a whole number of decimation steps (200) and of squelch blocks (750 x 200)
    pos = 0
    decoders = {}
    try:
        while True:
            mk = rp.next_mark()
            if mk is None or mk["name"] == "stopped":
                break
            name = mk["name"]
            audio_id = mk["readers"]["hd_audio" if mk.get("hd") else "audio"]
            a_sink, p_sink = rp.sink(audio_id), rp.sink(mk["readers"]["smeter"])
            f_sink = s_sink = if_sink = None
            if "secondary_fft" in mk:
                f_sink = rp.sink(mk["readers"]["secondary_fft"])
                s_sink = rp.sink(mk["secondary_selector"]["output"])
                # a third consumer of selectorBuffer, as a COMPLEX_FLOAT secondary demodulator would be (owrx/dsp.py:205)
                # selectorBuffer, found through the recorded wiring: the Buffer the SecondarySelector's Shift reads
                sel_buf = [o for o in rp.obj.values() if isinstance(o, trace_replay.M.Shift) and o._reader is not None
                           and o._reader._buffer is not rp.source() and not o._stopped][0]._reader._buffer
                if_sink = trace_replay.Sink(sel_buf.getReader())
            seg = make_iq(seg_len, fs, CARRIERS, seed=91, t0=pos)
            pos += seg_len
            rp.feed(seg)
            ref = oracle.client_chain_run(seg, fs, mk["out_rate"], mk["offset"], tuple(mk["bandpass"]), KIND[mk["demod"]],
                                          audio_rate=float(mk.get("audio_rate", 48000)), wfm_tau=mk.get("tau", 50e-6),
                                          agc_profile=1 if mk.get("agc") == "fast" else 0)
            # ---- audio: SYNC-framed IMA-ADPCM (AdpcmEncoder(sync=True)), decoded like the browser does (one codec per stream)
            dec = decoders.setdefault(audio_id, JsImaAdpcmCodec())
            got = dec.decodeWithSync(b"".join(a_sink.take())).astype(np.float64) / 32767.0
            want = ref["audio"]
            assert abs(len(got) - len(want)) <= 2100, (name, len(got), len(want))     # <= one SYNC period + the FIR lead
            powers = np.frombuffer(b"".join(p_sink.take()), np.float32)
            if mk["squelch_db"] > -100:
                # closed squelch: after the hang the audio is silence, and the S-meter keeps reporting the channel power
                tail = got[len(got) // 2:]
                assert np.abs(tail).max() <= 32.0 / 32767.0, (name, np.abs(tail).max())
                assert len(powers) >= 4 and np.all(powers < 10.0 ** (mk["squelch_db"] / 10.0)) and np.all(powers > 1e-4)
                _record("DspManager trace", name, audio_samples=len(got), tail_peak=float(np.abs(tail).max()), smeter_mean=float(powers.mean()))
                continue
            m = _compare_audio(got, want)
            _record("DspManager trace", name, audio_samples=len(got), oracle_samples=len(want), smeter_reports=len(powers), **m)
            assert m["corr"] >= 0.98, (name, m)
            assert 0.5 <= m["gain_ratio"] <= 2.0, (name, m)
            assert len(powers) >= 3
            if if_sink is None:
                continue
            # ---- selectorBuffer's other readers: the IF itself, the secondary FFT, the SecondarySelector
            rp.settle()
            if_got = np.frombuffer(b"".join(if_sink.take()), np.complex64)
            if_sink.reader.stop()
            assert len(if_got) >= len(ref["if_"]) - 60
            lead = len(if_got) - len(ref["if_"])                       # outputs whose FIR window straddles the segment start
            assert 0 <= lead <= 60
            a, b = if_got[lead + 400:].astype(np.complex128), ref["if_"][400:len(if_got) - lead].astype(np.complex128)
            c = np.vdot(b, a) / np.vdot(b, b)                         # the retune kept the NCO phase: a constant rotation
            e_if = float(np.sqrt(np.mean(np.abs(a - c * b) ** 2) / np.mean(np.abs(b) ** 2)))
            assert abs(abs(c) - 1.0) < 1e-3 and e_if <= 1e-4, (name, c, e_if)
            sf = mk["secondary_fft"]
            lines = f_sink.take()
            want_l = oracle.fftchain_run(if_got, sf["n"], sf["every_n"], sf["avg"])
            assert len(lines) == len(want_l["lines"]) >= 5
            same = sum(l == w.tobytes() for l, w in zip(lines, want_l["lines"]))
            e_gpu = np.sqrt(np.mean([(browser_fft_decode(np.frombuffer(l, np.uint8)) - db) ** 2 for l, db in zip(lines, want_l["db"])]))
            e_ref = np.sqrt(np.mean([(browser_fft_decode(w) - db) ** 2 for w, db in zip(want_l["lines"], want_l["db"])]))
            assert e_gpu <= 1.05 * e_ref + 0.02, (e_gpu, e_ref)
            ss = mk["secondary_selector"]
            sel_got = np.frombuffer(b"".join(s_sink.take()), np.complex64)
            cut = ss["bandwidth"] / mk["out_rate"]                      # SecondarySelector: Bandpass(-bw/rate, bw/rate, bw/rate) (selector.py:217-226)
            taps = oracle.firdes_bandpass(oracle.filter_len(cut), -cut, cut)
            sel_want = oracle.bandpass(oracle.shift(if_got, -ss["offset"] / mk["out_rate"]), taps)
            assert len(sel_got) == len(sel_want)
            e_sel = float(np.sqrt(np.mean(np.abs(sel_got - sel_want) ** 2) / np.mean(np.abs(sel_want) ** 2)))
            _record("DspManager trace", name + " (selectorBuffer readers)", if_rel_rms=e_if, if_lead=lead, secondary_fft_lines=len(lines),
                    byte_identical_lines=same, shown_rms_db_err_gpu=float(e_gpu), shown_rms_db_err_oracle_codec=float(e_ref),
                    secondary_selector_rel_rms=e_sel)
            assert e_sel <= 1e-4, e_sel
    finally:
        rp.close()
