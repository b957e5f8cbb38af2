#!/usr/bin/env python
"""Stall-reason totals of one kernel from an ncu report: python profiles/ncu_stalls.py <rep> <kernel regex>"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
tot = {}
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        for k, v in zip(hdr, r):
            if k.startswith("stall_") and "Not Issued" not in k:
                try: tot[k] = tot.get(k, 0) + int(v)
                except ValueError: pass
s = sum(tot.values()) or 1
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-28s %8d  %5.1f%%" % (k, v, 100.0 * v / s))
