#!/usr/bin/env python
"""Top CUDA source lines by warp-stall samples from an ncu report (needs -lineinfo + --import-source on).
usage: hot_lines.py report.ncu-rep kernel_regex [launch_skip] [top_n]"""
import csv
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      f"regex:{kernel}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
path, hdr, lines = "", None, []
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr is not None and r[0].isdigit() and len(r) >= len(hdr) - 2:
        col = {c: i for i, c in enumerate(hdr)}
        raw = r[col["# Samples"]]
        n = int(raw) if raw.isdigit() else 0
        stalls = {}
        for i, c in enumerate(hdr):
            if c.startswith("stall_") and "Not Issued" not in c and i < len(r) and r[i].isdigit() and r[i] != "0":
                stalls[c[6:]] = int(r[i])
        lines.append((n, path, int(r[0]), r[1].strip()[:90], sorted(stalls.items(), key=lambda kv: -kv[1])[:2],
                      r[col["Instructions Executed"]]))
tot = sum(l[0] for l in lines) or 1
print(f"total samples {tot}")
for n, p, ln, src, st, ie in sorted(lines, key=lambda l: -l[0])[:top_n]:
    print(f"{n:6d} {100.0 * n / tot:5.1f}% {p}:{ln:<4d} inst={ie:>8s} {st}  | {src}")
