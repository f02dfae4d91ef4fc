#!/usr/bin/env python
"""Turn one `ncu --set full` report of scripts/profile_step.py into the files committed under profiles/:
   <prefix>.txt (one summary line + stall mix per kernel), <prefix>_raw.csv (selected raw metrics per launch) and
   r02_ncu_traffic.json (DRAM bytes per launch of each kernel, read by bench.py for roofline.traffic; stamped with the
   hash of the CUDA sources so that bench.py stops quoting it once the kernels change).
   usage: ncu_to_profiles.py report.ncu-rep profiles/r02_ncu_step_b4096_fp32 [batch]"""
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_source_hash                                          # noqa: E402

rep, prefix = sys.argv[1], sys.argv[2]
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
here = os.path.dirname(os.path.abspath(__file__))
summary = subprocess.run([sys.executable, os.path.join(here, "ncu_summary.py"), rep], capture_output=True, text=True).stdout
open(prefix + ".txt", "w").write(summary)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_config_size", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
keep = [k for k in keep if k in ix]
with open(prefix + "_raw.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keep)
    w.writerow([units[ix[k]] for k in keep])
    for r in rows[2:]:
        w.writerow([r[ix[k]] for k in keep])

NAMES = {"lbs_fwd_kernel": "lbs_fwd", "lbs_bwd_kernel": "lbs_bwd", "joints_fwd_kernel": "joints_fwd",
         "joints_bwd_kernel": "joints_bwd", "umma_gemm2_kernel": "blend_bwd_umma", "umma_gemm_kernel": "blend_bwd_umma",
         "blend_fwd_bs2_kernel": "blend_fwd_umma", "blend_fwd_ws2_kernel": "blend_fwd_umma", "blend_fwd_ws_kernel": "blend_fwd_umma",
         "pose_fwd_lb_kernel": "pose_fwd", "pose_bwd_lb_kernel": "pose_bwd", "pose_fwd_kernel": "pose_fwd", "pose_bwd_kernel": "pose_bwd"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


traffic = {}
for r in rows[2:]:                                  # the last launch of every kernel wins (warm model, steady state)
    name = r[ix["Kernel Name"]]
    for k, short in NAMES.items():
        if k + "<" in name or k + "(" in name:
            traffic[short] = (to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) +
                              to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]]))
            break
traffic["source"] = os.path.basename(prefix) + "_raw.csv (ncu --set full --clock-control none, B=%d fp32, dram__bytes_read.sum + dram__bytes_write.sum per launch)" % batch
traffic["batch"] = batch
traffic["kernel_source_hash"] = kernel_source_hash()
json.dump(traffic, open(os.path.join(os.path.dirname(prefix), "r02_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
