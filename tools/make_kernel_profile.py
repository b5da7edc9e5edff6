#!/usr/bin/env python
"""Turn an `ncu --metrics ... --csv` launch list of one bench.py run into profiles/<name>.json:
per kernel (base name, template arguments stripped) the per-launch averages of duration, warp instructions,
DRAM bytes and pipe utilisation.  bench.py reads it to express the dominant kernel's live CUDA-event time as an
instruction-issue rate (the kernels of this path are integer/FP32 issue bound, not HBM bound).

    python tools/make_kernel_profile.py gpurun_out/launches.csv profiles/r1_kernel_profile.json "<workload name>"
"""
import collections
import csv
import json
import re
import sys

UNIT = {"usecond": 1e-6, "us": 1e-6, "msecond": 1e-3, "ms": 1e-3, "nsecond": 1e-9, "ns": 1e-9, "second": 1.0, "s": 1.0,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1.0, "%": 1.0, "": 1.0}


def base_name(k):
    k = re.sub(r"^void\s+", "", k.strip())
    k = k.split("(")[0]
    k = re.sub(r"<.*>$", "", k)
    return k.split("::")[-1]


def main(src, dst, workload):
    rows = list(csv.reader(open(src, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, mi, ui, vi, idi = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        except ValueError:
            continue
        per.setdefault((r[idi], base_name(r[ki])), {})[r[mi]] = v
    agg = collections.OrderedDict()
    for (_, k), m in per.items():
        a = agg.setdefault(k, collections.defaultdict(list))
        for kk, v in m.items():
            a[kk].append(v)
    out = {"workload": workload, "source": src, "how": "ncu --clock-control none, per-launch averages (cold-cache, serialised: use shares and counts, not absolute times)", "kernels": {}}
    for k, a in agg.items():
        g = lambda name: (sum(a[name]) / len(a[name])) if a.get(name) else None
        out["kernels"][k] = {
            "launches_profiled": len(a["gpu__time_duration.sum"]),
            "duration_us": round(g("gpu__time_duration.sum") * 1e6, 2),
            "warp_inst": g("smsp__inst_executed.sum"),
            "dram_read_bytes": g("dram__bytes_read.sum"), "dram_write_bytes": g("dram__bytes_write.sum"),
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "threads_per_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
        }
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in out["kernels"].items():
        print("%-24s %8.1f us  %12.0f inst  issue %5.1f%%  alu %5.1f%%  fma %5.1f%%  dram %5.1f%%  rd %9.0f  wr %9.0f" % (
            k, v["duration_us"], v["warp_inst"] or 0, v["issue_active_pct"] or 0, v["alu_pipe_pct"] or 0, v["fma_pipe_pct"] or 0, v["dram_pct"] or 0,
            v["dram_read_bytes"] or 0, v["dram_write_bytes"] or 0))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
