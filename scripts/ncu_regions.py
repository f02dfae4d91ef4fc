#!/usr/bin/env python
"""Stall samples of one kernel summed over instruction-index ranges: ncu_regions.py rep kernel-regex step"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
step = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_i[0]]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
body = [r for r in rows[hdr_i[0] + 1:end] if len(r) == len(h)]
ix = {k: i for i, k in enumerate(h)}
stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[ix["# Samples"]]) for r in body)
for a in range(0, len(body), step):
    seg = body[a:a + step]
    s = sum(int(r[ix["# Samples"]]) for r in seg)
    ex = sum(int(r[ix["Instructions Executed"]]) for r in seg)
    agg = collections.Counter()
    for r in seg:
        for k in stall_cols:
            agg[k[6:]] += int(r[ix[k]])
    ops = collections.Counter(r[ix["Source"]].split()[0 if not r[ix["Source"]].strip().startswith("@") else 1].split(".")[0] for r in seg)
    print("%5d-%5d %5.1f%% exec=%9d  %s | %s" % (a, a + len(seg), 100.0 * s / tot, ex,
          " ".join("%s:%d" % kv for kv in agg.most_common(4)), " ".join("%s:%d" % kv for kv in ops.most_common(5))))
