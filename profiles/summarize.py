#!/usr/bin/env python3
"""Summarise ncu output brought back in gpurun_out/ into a committed markdown table.
usage: summarize.py <launches.csv> <raw.csv|-> <out.md> [title]"""
import collections
import csv
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        acc.setdefault(r[ki].split("(")[0].replace("void ", ""), []).append(v)
    return acc


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "smsp__warps_eligible.avg.per_cycle_active"]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "")}
        for w in want:
            if w in idx:
                d[w] = r[idx[w]] + " " + units[idx[w]]
        out.append(d)
    return out, want


def main():
    lpath, rpath, opath = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else "ncu summary"
    o = ["# " + title, ""]
    acc = launches(lpath)
    tot = sum(sum(v) for v in acc.values())
    o += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)",
          "", "| kernel | launches | avg ns | share of all kernel time |", "|---|---|---|---|"]
    for n, v in acc.items():
        o.append(f"| {n} | {len(v)} | {sum(v) / len(v):.0f} | {100 * sum(v) / tot:.1f}% |")
    if rpath != "-":
        rows, want = raw(rpath)
        o += ["", "## `ncu --set full` (one launch per kernel)", ""]
        for d in rows:
            o.append(f"### {d['kernel']}")
            o.append("")
            for w in want:
                if w in d:
                    o.append(f"- `{w}` = {d[w]}")
            o.append("")
    open(opath, "w").write("\n".join(o) + "\n")


if __name__ == "__main__":
    main()
