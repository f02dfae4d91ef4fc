#!/usr/bin/env python
"""SASS instructions of one kernel in address order with their stall samples and CUDA source line, from an
`ncu --set full --import-source on` report: ncu_sass_hot.py report.ncu-rep <kernel regex> [min samples] [launch index]
(tells WHICH inlined copy of a helper, e.g. which mbarrier wait, the samples belong to)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
min_s = int(sys.argv[3]) if len(sys.argv) > 3 else 50
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, cur, sass = None, None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = {k: j for j, k in enumerate(r)}
    elif hdr is not None and len(r) > 10:
        if r[0] != "":
            cur = (fname, r[0])
        elif r[2].startswith("0x"):
            num = lambda v: int(v) if v.lstrip("-").isdigit() else 0   # noqa: E731
            stalls = {k[6:]: num(r[j]) for k, j in hdr.items() if k.startswith("stall_") and "Not Issued" not in k}
            sass.append((int(r[2], 16), r[3].strip(), num(r[hdr["# Samples"]]), num(r[hdr["Instructions Executed"]]), cur, stalls))
sass.sort()
base = sass[0][0] if sass else 0
tot = sum(x[2] for x in sass) or 1
print("kernel %s: %d samples" % (kern, tot))
for a, ins, s, n, src, st in sass:
    if s >= min_s:
        top = ", ".join("%s %d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3] if kv[1] > 0)
        print("%6x %6d %5.1f%% %9d  %-22s %-60s | %s" % (a - base, s, 100.0 * s / tot, n, "%s:%s" % src, ins[:60], top))
