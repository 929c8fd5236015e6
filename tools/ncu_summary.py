#!/usr/bin/env python
"""Summarise ncu output into profiles/: per-kernel launch times (launch list CSV) and, if a
.ncu-rep is given, the key metrics of each profiled kernel.

    python tools/ncu_summary.py --launches gpurun_out/launches.csv [--rep gpurun_out/prof.ncu-rep] --out profiles/r1_x
"""
import argparse
import collections
import csv
import json
import subprocess

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--out", required=True)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    lines = ["# ncu summary" + (" -- " + a.note if a.note else ""), ""]
    doc = {}
    if a.launches:
        rows = [r for r in csv.reader(open(a.launches)) if len(r) > 5]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = collections.OrderedDict()
        for r in rows[1:]:
            try:
                agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
        ours = {k: v for k, v in agg.items() if "lss::" in k}
        step = sum(sum(v) / len(v) for v in ours.values())
        lines += ["## launch list (gpu__time_duration.sum, ns; serialised, cold caches: compare shares)", "",
                  "| kernel | launches | mean ns | min ns | share of our step |", "|---|---|---|---|---|"]
        for k, v in agg.items():
            mean = sum(v) / len(v)
            share = ("%.1f %%" % (100 * mean / step)) if k in ours else "-"
            lines.append("| `%s` | %d | %.0f | %.0f | %s |" % (k[:90], len(v), mean, min(v), share))
            doc[k] = {"launches": len(v), "mean_ns": mean, "min_ns": min(v)}
        lines += ["", "sum of our kernels' mean times: %.1f us" % (step / 1e3), ""]
    if a.rep:
        out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        lines += ["## per-kernel metrics (ncu --set full, one launch each)", ""]
        seen = set()
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            if name in seen:
                continue
            seen.add(name)
            lines.append("### `%s`" % name[:100])
            lines.append("")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    lines.append("- %s = %s %s" % (m, r[i], units[i]))
            lines.append("")
    with open(a.out + ".md", "w") as f:
        f.write("\n".join(lines) + "\n")
    if doc:
        with open(a.out + ".json", "w") as f:
            json.dump(doc, f, indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
