#!/usr/bin/env python
"""DRAM traffic of ONE step of a BASELINE config, summed over every kernel of the step (VERDICT r1 task 2).

  run under ncu (one config per run; only the bracketed steps are profiled):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
        --profile-from-start off --csv --log-file gpurun_out/traffic_c2.csv python tools/step_traffic.py c2
  then, here:
    python tools/step_traffic.py --summarise gpurun_out/traffic_*.csv > profiles/r2_step_traffic.json

--cache-control none keeps the L2 contents between the kernels of a step (ncu's default flush would charge every inter-kernel
hand-over to DRAM); the launches are still serialised, so overlap between the pipeline's streams is not represented."""
import collections
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STEPS = 2


def run(which):
    import torch
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    rt = torch.cuda.cudart()
    if which in ("c2", "c3", "c5"):
        cfg, n_ch = {"c2": (bench.C2, 64), "c3": (bench.C3, 1024), "c5": (bench.C5, 128)}[which]
        cars = bench.channel_plan(0, n_ch, cfg["fs"], wfm=bool(cfg.get("wfm")))
        bank, _ = bench.make_bank(cfg, cars, 0)
        blocks = [bench.quick_iq_torch(cfg["block"], dev, 11), bench.quick_iq_torch(cfg["block"], dev, 12)]
        bank.set_pipelined(True)
        for i in range(3):
            bank.process_device(blocks[i & 1], cfg["block"], stream=stream.cuda_stream)
        bank.join(stream.cuda_stream); torch.cuda.synchronize()
        rt.cudaProfilerStart()
        for i in range(STEPS):
            bank.process_device(blocks[(i + 1) & 1], cfg["block"], stream=stream.cuda_stream)
        bank.join(stream.cuda_stream); torch.cuda.synchronize()
        rt.cudaProfilerStop()
    else:
        from openwebrx_b200 import Waterfall
        cfg = {"c1": bench.C1, "c4": bench.C4}[which]
        m = bench.waterfall_model(cfg)
        ns = m["every_n"] * m["avg"] * cfg["lines"] + cfg["n"]
        iq = bench.quick_iq_torch(ns, dev, 7)
        wf = Waterfall(cfg["fs"], cfg["n"], cfg["ov"], cfg["fps"], "adpcm", device=0)
        if cfg.get("noise_filter"):
            wf.set_noise_filter(True)
        out = torch.empty(cfg["lines"] * wf.line_bytes, dtype=torch.uint8, device=dev)
        wf.set_pipelined(True)
        for _ in range(2):
            wf.process_device(iq, ns, out, out.numel(), stream=stream.cuda_stream)
        wf.join(stream.cuda_stream); torch.cuda.synchronize()
        rt.cudaProfilerStart()
        for _ in range(STEPS):
            wf.process_device(iq, ns, out, out.numel(), stream=stream.cuda_stream)
        wf.join(stream.cuda_stream); torch.cuda.synchronize()
        rt.cudaProfilerStop()
    print("profiled %d steps of %s" % (STEPS, which))


def summarise(paths):
    res = {}
    for path in paths:
        which = os.path.basename(path).split("_")[-1].split(".")[0].upper()
        rows = [l for l in open(path) if l.startswith('"')]
        r = csv.reader(io.StringIO("".join(rows)))
        hdr = next(r)
        ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        per = collections.OrderedDict()
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
        for row in r:
            name = row[ki].split("(")[0].replace("void ", "").replace("owrx::", "").replace("(anonymous namespace)::", "")
            v = float(row[vi].replace(",", "")) * scale.get(row[ui], 1.0)
            d = per.setdefault(name, {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
            if row[mi] == "gpu__time_duration.sum":
                d["launches"] += 1
                d["us"] += v
            elif row[mi] == "dram__bytes_read.sum":
                d["dram_read"] += v
            elif row[mi] == "dram__bytes_write.sum":
                d["dram_write"] += v
        for d in per.values():
            for k in ("us", "dram_read", "dram_write"):
                d[k] /= STEPS
            d["launches"] /= STEPS
        total = sum(d["dram_read"] + d["dram_write"] for d in per.values())
        res[which] = {"dram_bytes_per_step": total, "kernel_us_per_step_serialised": sum(d["us"] for d in per.values()), "kernels": per,
                      "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none over %d steps (tools/step_traffic.py), "
                                "per step; launches serialised" % STEPS}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "--summarise":
        summarise(sys.argv[2:])
    else:
        run(sys.argv[1])
