#!/usr/bin/env python
"""Turns ncu outputs brought back from the GPU box (gpurun_out/) into the committed summaries under profiles/.
  ncu_summary.py launches <launches.csv> <out.md>       per-kernel table of gpu__time_duration (cold, serialised)
  ncu_summary.py kernel <report.ncu-rep> <out.md>       key metrics of one --set full capture
"""
import collections
import csv
import io
import subprocess
import sys

KEY = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "sm__inst_executed_pipe_uniform_realtime.avg.pct_of_peak_sustained_elapsed",
]


def launches(src, dst):
    rows = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(io.StringIO("".join(rows)))
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for row in r:
        name = row[ki].split("(")[0]
        v = float(row[vi].replace(",", ""))
        v = v / 1e3 if row[ui] == "ns" else (v * 1e3 if row[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f %% |\n" % (k, n, t, t / n, 100 * t / tot))
    print(open(dst).read())


def kernel(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write("### %s\n\n| metric | unit | value |\n|---|---|---:|\n" % d.get("Kernel Name", "?"))
            for h, u, v in zip(hdr, units, vals):
                if any(h == k or h.endswith("." + k) for k in KEY) or h.endswith("_per_issue_active.ratio"):
                    try:
                        if h.endswith("ratio") and float(v) < 0.02:
                            continue
                    except ValueError:
                        pass
                    f.write("| %s | %s | %s |\n" % (h, u, v))
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
