"""Builds libowrx_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery).

The library is the product: hand-written CUDA kernels + the C ABI of include/owrx_b200.h.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
SO = os.path.join(HERE, "libowrx_b200.so")
SOURCES = ["core.cu", "waterfall.cu", "selector.cu", "k3_fir.cu", "fastconv.cu", "fastconv_tc.cu", "iq_hop.cu"]
# per-file extra flags: K3's FFMA2 loop is scheduled better by ptxas -O1 (see k3_fir.cu)
EXTRA_FLAGS = {"k3_fir.cu": ["-Xptxas", "-O1"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "owrx_b200.h"))
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + EXTRA_FLAGS.get(src, []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed.append("%s:\n%s" % (src, out))
        elif verbose and out:
            print(out, file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(failed))
    if force or procs or _stale(SO, objs):
        # the CUDA runtime is linked dynamically (libcudart.so.12: the process shares ONE runtime with torch when both are
        # loaded, and the artefact carries only the runtime symbols it calls); rpath covers processes that load no torch
        cmd = [nvcc, "-shared", "-cudart", "shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                      "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
