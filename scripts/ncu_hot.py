#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel in an .ncu-rep (source page): ncu_hot.py rep kernel-regex [N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first launch only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_i[0]]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
body = [r for r in rows[hdr_i[0] + 1:end] if len(r) == len(h)]
ix = {k: i for i, k in enumerate(h)}
tot = sum(int(r[ix["# Samples"]]) for r in body)
stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:n]
for i in sorted(order):
    r = body[i]
    s = int(r[ix["# Samples"]])
    top = sorted(((int(r[ix[k]]), k[6:]) for k in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%% exec=%-8s %-60s %s" % (i, 100.0 * s / tot, r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:60],
                                           " ".join("%s:%d" % (k, v) for v, k in top if v)))
