"""IQ ingress (owrx/source/__init__.py:307-330): the reference's SdrSource.getBuffer() wires
TcpSource(port, Format) -> Buffer(COMPLEX_FLOAT), and every consumer (SpectrumThread's FftChain, each client's Selector) attaches
a Reader to that Buffer.  Here that Buffer's storage is a page-locked ring: TcpSource recv()s into it and the runner feeds the
copy engine pointers into it.  A loopback connector streams 10 MS/s complex float32 (and 2.4 MS/s complex int16 through the
source-side Convert + Gain chain of owrx/source/fifi_sdr.py:27-28) into a waterfall and two clients; every byte that reaches
the GPU must come from page-locked memory (owrx_bank_get_stats / owrx_wf_get_h2d_bytes), and the outputs must match the oracle."""
import ctypes as C
import socket
import threading
import time

import numpy as np
import pytest

import oracle
import pycsdr.modules as M
from openwebrx_b200 import _native as N
from openwebrx_b200.synth import BANDPASS, make_iq
from pycsdr.types import AgcProfile, Format
from test_gpu_pycsdr import _collect, _connect
from test_oracle import browser_fft_decode

pytestmark = pytest.mark.gpu


def _serve(payload, chunk=1 << 16, pace=None):
    """loopback 'connector' (rtl_connector & co. serve raw IQ on a TCP port): returns (port, thread)"""
    srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]

    def run():
        conn, _ = srv.accept()
        try:
            mv = memoryview(payload)
            for o in range(0, len(mv), chunk):
                conn.sendall(mv[o:o + chunk])           # chunk sizes are NOT multiples of the sample size on purpose
                if pace:
                    time.sleep(pace)
            time.sleep(0.5)
        finally:
            conn.close(); srv.close()
    th = threading.Thread(target=run, daemon=True)
    th.start()
    return port, th


def _client(src, fs, out_rate, offset, kind):
    from openwebrx_b200.params import decimator_params
    decim, frac, transition, cutoff = decimator_params(fs, out_rate)      # Decimator, csdr/chain/selector.py:21-51
    agc = M.Agc(Format.FLOAT); agc.setProfile(AgcProfile.SLOW)
    if kind == "nfm":
        agc.setMaxGain(3)
        demod = [M.FmDemod(), M.Limit(), M.NfmDeemphasis(out_rate), agc]
    else:
        agc.setInitialGain(200)
        demod = [M.AmDemod(), M.DcBlock(), agc]
    bp = M.Bandpass(transition=320.0 / out_rate, use_fft=True)
    lo, hi = BANDPASS[kind]
    bp.setBandpass(lo / out_rate, hi / out_rate)
    sh = M.Shift(0.0); sh.setRate(-offset / fs)
    ws = [sh, M.FirDecimate(decim, transition, cutoff)]
    if frac != 1.0:
        ws.append(M.FractionalDecimator(Format.COMPLEX_FLOAT, frac))
    ws += [bp, M.Squelch(Format.COMPLEX_FLOAT, length=750, decimation=5, hangLength=1500, flushLength=3750, reportInterval=4)] + demod
    _connect(ws)
    ob = M.Buffer(Format.FLOAT)
    ws[-1].setWriter(ob)
    rd = ob.getReader()
    ws[0].setReader(src.getReader())
    return ws, rd


def _waterfall(src, n, every_n, avg):
    ws = [M.Fft(size=n, every_n_samples=0), M.LogAveragePower(add_db=-70, fft_size=n, avg_number=avg), M.FftSwap(fft_size=n),
          M.FftAdpcm(fft_size=n)]
    ws[0].setEveryNSamples(every_n)
    _connect(ws)
    out = M.Buffer(Format.CHAR)
    ws[-1].setWriter(out)
    rd = out.getReader()
    ws[0].setReader(src.getReader())
    return ws, rd


def _runner_stats(src):
    r = src._runner
    st = N.BankStats()
    N.check(N.lib.owrx_bank_get_stats(r.bank, C.byref(st)))
    wf_pin = wf_page = 0
    for p in r.wf_plans.values():
        a, b = C.c_uint64(), C.c_uint64()
        N.check(N.lib.owrx_wf_get_h2d_bytes(p.handle, C.byref(a), C.byref(b)))
        wf_pin += a.value; wf_page += b.value
    return st, wf_pin, wf_page


def _rel(got, want):
    return float(np.sqrt(np.mean((got - want) ** 2)) / np.sqrt(np.mean(want ** 2)))


def test_pinned_alloc_is_page_locked(gpu):
    p = C.c_void_p()
    N.check(N.lib.owrx_pinned_alloc(1 << 20, C.byref(p)))
    try:
        assert N.lib.owrx_host_is_pinned(p) == 1 and N.lib.owrx_host_is_pinned(C.c_void_p(p.value + 12345)) == 1
        pageable = np.zeros(1024, np.float32)
        assert N.lib.owrx_host_is_pinned(pageable.ctypes.data_as(C.c_void_p)) == 0
    finally:
        N.lib.owrx_pinned_free(p)


def test_tcp_float32_10msps_lands_in_the_pinned_ring(gpu, monkeypatch):
    monkeypatch.setenv("OWRX_RING_MB", "16")                # 20 MB of stream through a 16 MB ring: it wraps
    fs, out_rate, decim = 10_000_000, 12000, 833
    cars = [dict(offset=1_250_000, amp=0.2, kind="nfm"), dict(offset=-2_100_000, amp=0.1, kind="am")]
    n_fft, every_n, avg = 4096, 11905, 4                   # ~ 9 fps-class line rate at 10 MS/s, small averaging for the test
    lines = 40
    n = every_n * avg * lines + n_fft + 3
    iq = make_iq(n, fs, cars, seed=61)
    port, th = _serve(iq.tobytes(), chunk=65521, pace=0.0004)   # prime chunk size: samples split across recv() calls; ~2x real time
    src = M.Buffer(Format.COMPLEX_FLOAT)
    wf, wf_rd = _waterfall(src, n_fft, every_n, avg)
    clients = [_client(src, fs, out_rate, c["offset"], c["kind"]) for c in cars]
    tcp = M.TcpSource(port, Format.COMPLEX_FLOAT)           # owrx/source/__init__.py:310-314
    tcp.setWriter(src)                                      # owrx/source/__init__.py:326-330
    lb = (n_fft + 10) // 2
    msgs = _collect(wf_rd, lines * lb, timeout=60)
    th.join(30)
    assert src._ring, "the source Buffer has no page-locked ring on a GPU box"
    assert len(msgs) == lines and all(len(m) == lb for m in msgs)
    ref = oracle.fftchain_run(iq, n_fft, every_n, avg)
    for m, db in zip(msgs, ref["db"]):
        assert np.median(np.abs(browser_fft_decode(np.frombuffer(m, np.uint8)) - db)) < 0.6
    kind = {"nfm": oracle.DEMOD_NFM, "am": oracle.DEMOD_AM}
    for (ws, rd), c in zip(clients, cars):
        want = oracle.client_chain_run(iq, fs, out_rate, c["offset"], BANDPASS[c["kind"]], kind[c["kind"]])["audio"]
        got = np.frombuffer(b"".join(_collect(rd, 4 * (len(want) - 8), timeout=60)), np.float32)
        k = min(len(got), len(want))
        assert k >= len(want) - 8 and _rel(got[:k], want[:k]) < 1e-2      # post-AGC (spec-defined), as in test_gpu_pycsdr
    st, wf_pin, wf_page = _runner_stats(src)
    assert st.h2d_pageable_bytes == 0 and wf_page == 0, "a pageable copy is left on the ingress path"
    # (the waterfall does not upload the samples a frame hop larger than the FFT skips)
    assert st.h2d_pinned_bytes >= (n - 4 * decim - 30000) * 8 and wf_pin >= 0.9 * n * 8
    assert getattr(tcp, "bytes_direct", 0) >= n * 8 - 8     # recv_into() the ring, not recv() + copy
    tcp.stop()
    for ws, rd in clients:
        ws[0].stop()
    wf[0].stop()


def test_tcp_cs16_2_4msps_through_source_side_convert(gpu):
    """owrx/source/fifi_sdr.py:27-28 + owrx/source/direct.py:59-71: TcpSource(COMPLEX_SHORT) -> Convert -> Gain(5.0) -> the
    COMPLEX_FLOAT source Buffer.  The raw int16 samples land in that Buffer's ring and cross PCIe as they are."""
    fs, out_rate, decim = 2_400_000, 12000, 200
    cars = [dict(offset=250_000, amp=0.2, kind="nfm"), dict(offset=-321_000, amp=0.1, kind="am")]
    n = 5333 + 200 * (750 * 4 + 20)
    iq = make_iq(n, fs, cars, seed=62)
    raw = (np.clip(iq.view(np.float32), -1, 1) * 6000).astype(np.int16)     # x5 gain restores the scale
    as_float = oracle.convert_raw_iq(raw, "cs16", 5.0).view(np.complex64)
    port, th = _serve(raw.tobytes(), chunk=40009)
    tcp = M.TcpSource(port, Format.COMPLEX_SHORT)
    conv, gain = M.Convert(Format.COMPLEX_SHORT, Format.COMPLEX_FLOAT), M.Gain(Format.COMPLEX_FLOAT, 5.0)
    _connect([conv, gain])
    src = M.Buffer(Format.COMPLEX_FLOAT)
    gain.setWriter(src)
    clients = [_client(src, fs, out_rate, c["offset"], c["kind"]) for c in cars]
    wf, wf_rd = _waterfall(src, 1024, 7000, 4)
    raw_buf = M.Buffer(Format.COMPLEX_SHORT)
    conv.setReader(raw_buf.getReader())
    tcp.setWriter(raw_buf)
    kind = {"nfm": oracle.DEMOD_NFM, "am": oracle.DEMOD_AM}
    for (ws, rd), c in zip(clients, cars):
        want = oracle.client_chain_run(as_float, fs, out_rate, c["offset"], BANDPASS[c["kind"]], kind[c["kind"]])["audio"]
        got = np.frombuffer(b"".join(_collect(rd, 4 * (len(want) - 8), timeout=60)), np.float32)
        k = min(len(got), len(want))
        assert k >= len(want) - 8 and _rel(got[:k], want[:k]) < 1e-2
    th.join(30)
    st, wf_pin, wf_page = _runner_stats(src)
    assert src._ring and st.h2d_pageable_bytes == 0 and wf_page == 0
    assert st.h2d_pinned_bytes >= (n - 4 * decim - 30000) * 4              # 4 bytes per complex int16 sample crossed PCIe
    tcp.stop()
    for ws, rd in clients:
        ws[0].stop()
    wf[0].stop()
