#!/usr/bin/env python
"""Per-source-line summary of an ncu report: python profiles/ncu_lines.py <rep> <kernel regex> [top N]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; fname = ""
agg = {}
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":          # source-line row (aggregated over its SASS)
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            samp = int(d["# Samples"]); inst = int(d["Instructions Executed"])
        except ValueError:
            continue
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1]])
        a[0] += samp; a[1] += inst
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100 * a[0] / tot_s, 100 * a[1] / tot_i, f, ln, a[2].strip()[:100]))
