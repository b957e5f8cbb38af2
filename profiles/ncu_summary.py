#!/usr/bin/env python
"""Summaries for profiles/: python profiles/ncu_summary.py launches <csv>   |   ... kernels <ncu-rep>"""
import csv, io, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        seq.append((row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), v, row["Grid Size"], row["Block Size"]))
    # last step = everything after the last L2-flush fill kernel of torch
    last = max(i for i, s in enumerate(seq) if "FillFunctor" in s[0])
    step = seq[last + 1:]
    tot = sum(v for _, v, _, _ in step)
    print("# launches of the last bench step (ncu --metrics gpu__time_duration.sum --clock-control none; serialised, cold cache)")
    print("# %d launches, %.1f us of kernel time" % (len(step), tot))
    agg = {}
    for n, v, g, b in step:
        a = agg.setdefault(n, [0, 0.0, g, b]); a[0] += 1; a[1] += v
    print("%-44s %6s %10s %7s  %s" % ("kernel", "count", "us", "share", "grid / block (last)"))
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %6d %10.1f %6.1f%%  %s / %s" % (n[:44], a[0], a[1], 100 * a[1] / tot, a[2], a[3]))

def kernels(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        if name in seen:
            continue
        seen.add(name)
        print("## " + name)
        for w in want:
            if w in idx:
                print("  %-62s %s %s" % (w, r[idx[w]], units[idx[w]]))

if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2])
