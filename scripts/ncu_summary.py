#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into one line per kernel + top stall reasons."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return float("nan")


cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB_rd"), ("dram__bytes_write.sum", "MB_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("smsp__inst_executed.sum", "Minst"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf")]
stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0][-40:]
    parts = []
    for k, lab in cols:
        if k in ix:
            v = f(r, k)
            u = units[ix[k]]
            if lab.startswith("MB"):
                v = v / 1e6 if u == "byte" else (v if u == "Mbyte" else v * 1e3 if u == "Gbyte" else v / 1e3)
            if lab == "Minst":
                v /= 1e6
            if lab == "us" and u == "ms":
                v *= 1e3
            parts.append("%s=%.4g" % (lab, v))
    print(name, " ".join(parts))
    sv = sorted(((f(r, h), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stalls), reverse=True)
    tot = sum(v for v, _ in sv if v == v)
    print("    stalls:", ", ".join("%s %.0f%%" % (h, 100 * v / tot) for v, h in sv[:7] if v == v))
