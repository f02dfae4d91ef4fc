#!/usr/bin/env python
"""Top CUDA source lines of one kernel by warp-stall samples, from an `ncu --set full --import-source on` report:
   ncu_hot_lines.py report.ncu-rep <kernel-name regex> [top N] [launch index]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = {k: j for j, k in enumerate(r)}
    elif hdr is not None and len(r) > 10 and r[0] != "":
        num = lambda v: int(v) if v.lstrip("-").isdigit() else 0   # noqa: E731
        lines.append((num(r[hdr["# Samples"]]), num(r[hdr["Instructions Executed"]]), fname, r[0], r[1].strip()))
tot = sum(x[0] for x in lines) or 1
print("kernel %s: %d samples, %d warp instructions" % (kern, tot, sum(x[1] for x in lines)))
for s, n, f, ln, src in sorted(lines, reverse=True)[:top]:
    print("%6d %5.1f%% %9d  %s:%s  %s" % (s, 100.0 * s / tot, n, f, ln, src[:120]))
