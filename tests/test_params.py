"""CPU tests: the host-side parameter math (openwebrx_b200.params, mirrored in the C ABI and the
oracle) against argument lists captured from the reference's own unmodified chain classes
(tests/golden/params_reference.json, produced by tests/golden/make_golden.py)."""
import json
import os

import oracle
from openwebrx_b200 import params

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "params_reference.json")) as f:
    REF = json.load(f)


def test_fftchain_parameters_match_reference_classes():
    assert len(REF["fftchain"]) >= 7
    for case in REF["fftchain"]:
        fs, n, ov, fps = case["args"]
        avg, every_n = params.fftchain_params(fs, n, ov, fps)
        assert avg == case["avg_number"] and every_n == case["every_n_samples"], case["args"]
        assert (avg == 0) == (case["averager"] == "LogPower")
        assert oracle.fftchain_params(fs, n, ov, fps) == (avg, every_n)


def test_decimator_parameters_match_reference_classes():
    assert len(REF["decimator"]) >= 10
    for case in REF["decimator"]:
        fs, out = case["args"]
        d, frac, tr, cut = params.decimator_params(fs, out)
        assert (d, frac, tr, cut) == (case["decimation"], case["fraction"], case["transition"], case["cutoff"])
        od, ofrac, otr, ocut = oracle.decimator_params(fs, out)
        assert (od, ofrac, otr, ocut) == (d, frac, tr, cut)
        assert oracle.filter_len(tr) == params.filter_length(tr)


def test_selector_call_trace():
    log = REF["selector"]
    calls = {(c[1], json.dumps(c[2]), json.dumps(c[3], sort_keys=True)) for c in log}
    sq = params.squelch_params(12000)
    assert ("Squelch", json.dumps(["Format.COMPLEX_FLOAT"]), json.dumps(sq, sort_keys=True)) in calls
    tr, lo, hi = params.bandpass_params(12000, -5999, 5999)
    assert ("Bandpass", "[]", json.dumps({"transition": tr, "use_fft": True}, sort_keys=True)) in calls
    assert ("Bandpass.setBandpass", json.dumps([lo, hi]), "{}") in calls
    assert ("Shift.setRate", json.dumps([params.shift_rate(1234567, 10000000)]), "{}") in calls
    assert ("Squelch.setSquelchLevel", json.dumps([float(10 ** (-60 / 10))]), "{}") in calls


def test_derived_shapes_of_baseline_configs():
    # SURVEY Appendix B
    assert params.fftchain_params(2.4e6, 4096, 0.3, 9) == (93, 2867)
    assert params.fftchain_params(61.44e6, 65536, 0.3, 30) == (45, 45511)
    for fs, out, D, T in ((2.4e6, 12000, 200, 5333), (10e6, 12000, 833, 22223), (61.44e6, 12000, 5120, 136533),
                          (20e6, 250000, 80, 2133)):
        d, _, tr, _ = params.decimator_params(fs, out)
        assert d == D and params.filter_length(tr) == T
