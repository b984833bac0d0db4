#!/usr/bin/env python
"""Opcode evidence for the shipped library, without a GPU: per kernel of libowrx_b200.so the counts of the SASS mnemonics that
prove (or disprove) a Blackwell-native path (B200_PROFILING.md "What proves a Blackwell-native kernel"):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, SYNCS = mbarrier, FFMA2/FADD2/FMUL2 = packed FP32,
LDGSTS = cp.async, HMMA = legacy mma.sync (must be absent), plus registers per thread from the cubin's resource usage.
Usage: sass_hist.py [lib.so] > profiles/rN_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "LDGSTS",
         "HMMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "MUFU", "VIMNMX", "VIMNMX3"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "openwebrx_b200", "libowrx_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
    kernels = collections.OrderedDict()
    name = None
    arch = set()
    for line in sass.splitlines():
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.search(r"\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    cols = [w for w in WATCH if any(k[w] for k in kernels.values())]
    print("SASS opcode histogram of `%s` (cuobjdump -sass; arch: %s; %d kernels)\n" % (os.path.relpath(so, ROOT), ", ".join(sorted(arch)), len(kernels)))
    print("| kernel | regs | instr | " + " | ".join(cols) + " |")
    print("|---|---:|---:|" + "---:|" * len(cols))
    tot = collections.Counter()
    for (mangled, c), nice in zip(kernels.items(), demangle):
        nice = re.sub(r"\(.*", "", nice.replace("owrx::", "").replace("(anonymous namespace)::", "")).replace("void ", "")
        print("| `%s` | %s | %d | " % (nice, regs.get(mangled, ""), sum(c.values())) + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
        tot.update(c)
    print("| **total** | | %d | " % sum(tot.values()) + " | ".join(str(tot[w]) if tot[w] else "" for w in cols) + " |")
    assert tot["HMMA"] == 0, "legacy mma.sync found"


if __name__ == "__main__":
    main()
