#!/usr/bin/env python
"""Summarise ncu output into profiles/: per-kernel launch list (time, instructions, DRAM and L2
bytes) and, if a .ncu-rep is given, the key metrics of each profiled kernel (its LAST captured
launch: the first launches of a bench run are warm-ups on empty inputs).

    python tools/ncu_summary.py --launches gpurun_out/launches_TAG.csv [--rep gpurun_out/prof_TAG.ncu-rep] --out profiles/TAG
"""
import argparse
import collections
import csv
import json
import subprocess

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
           "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("lss::", "")


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
        ii, ki, mi, vi = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Value"))
        per = collections.OrderedDict()
        for r in rows[1:]:
            try:
                per.setdefault((r[ii], r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
            except ValueError:
                pass
        agg = collections.OrderedDict()
        for (_, k), m in per.items():
            if "lss::" in k:
                agg.setdefault(short(k), []).append(m)
        # the second half of each kernel's launches runs on real inputs (the first are construction warm-ups)
        med = {}
        for k, ms in agg.items():
            ms = ms[len(ms) // 2:]
            med[k] = {key: sorted(m.get(key, 0.0) for m in ms)[len(ms) // 2] for key in ms[0]}
            med[k]["launches"] = len(agg[k])
        step = sum(m["gpu__time_duration.sum"] for m in med.values())
        lines += ["## launch list (medians over the real-input launches; serialised, cold caches: compare shares)", "",
                  "| kernel | launches | time us | share of step | warp instructions | DRAM read MB | DRAM write MB | L2 MB |",
                  "|---|---|---|---|---|---|---|---|"]
        for k, m in med.items():
            lines.append("| `%s` | %d | %.2f | %.1f %% | %.0f | %.2f | %.2f | %.2f |" % (
                k, m["launches"], m["gpu__time_duration.sum"] / 1e3, 100 * m["gpu__time_duration.sum"] / step,
                m.get("smsp__inst_executed.sum", 0), m.get("dram__bytes_read.sum", 0) / 1e6,
                m.get("dram__bytes_write.sum", 0) / 1e6, m.get("lts__t_bytes.sum", 0) / 1e6))
        lines += ["", "sum of our kernels' times: %.1f us; instructions %.1f M; L2 traffic %.0f MB" % (
            step / 1e3, sum(m.get("smsp__inst_executed.sum", 0) for m in med.values()) / 1e6,
            sum(m.get("lts__t_bytes.sum", 0) for m in med.values()) / 1e6), ""]
        doc["launch_list"] = med
    if a.rep:
        out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        lines += ["## per-kernel metrics (ncu --set full, last captured launch of each kernel)", ""]
        last = collections.OrderedDict()
        for r in rows[2:]:
            last[r[hdr.index("Kernel Name")]] = r
        doc["full"] = {}
        for name, r in last.items():
            lines.append("### `%s`" % name[:100])
            lines.append("")
            d = {}
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    lines.append("- %s = %s %s" % (m, r[i], units[i]))
                    d[m] = r[i]
            doc["full"][short(name)] = d
            lines.append("")
    with open(a.out + ".md", "w") as f:
        f.write("\n".join(lines) + "\n")
    if doc:
        with open(a.out + ".json", "w") as f:
            json.dump(doc, f, indent=1)
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    main()
