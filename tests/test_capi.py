"""CPU tests: the C-ABI library loads, exports every symbol include/owrx_b200.h declares, and fails
loudly (no CPU fallback) when no GPU is present.  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "owrx_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(owrx_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    from openwebrx_b200 import _native as N
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(N.lib, n), "libowrx_b200.so does not export %s" % n
        assert n in N.SIGNATURES, "ctypes binding missing for %s" % n
    assert set(N.SIGNATURES) == set(names)
    assert b"sm_100a" in N.lib.owrx_version()


def test_library_embeds_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "openwebrx_b200", "libowrx_b200.so")],
                         capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu():
    from openwebrx_b200 import _native as N
    n = C.c_int()
    rc = N.lib.owrx_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    from openwebrx_b200 import ChannelBank, Waterfall
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        Waterfall(2.4e6, 4096, 0.3, 9)
    with pytest.raises(RuntimeError):
        ChannelBank(2.4e6)


def test_product_never_imports_the_oracle():
    for base in ("openwebrx_b200", "pycsdr"):
        d = os.path.join(ROOT, base)
        if not os.path.isdir(d):
            continue
        for dp, _, files in os.walk(d):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), (dp, f)
                    assert "csdr_oracle" not in txt, (dp, f)
