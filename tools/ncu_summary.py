#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md / bench.py
quote, and append them to a JSON file under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --workload c4-single --objects 67108864 --views 1 \
           --out profiles/r01_cull_kernel_summary.json
"""
import argparse
import csv
import io
import json
import os
import subprocess

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "dram__bytes_read.sum.per_second": "dram_read_rate",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed.avg.per_cycle_elapsed": "inst_per_cycle_per_sm",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__inst_executed_op_local_ld.sum": "local_loads",
    "smsp__inst_executed_op_local_st.sum": "local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "byte/s": 1, "Gbyte/s": 1e9, "Tbyte/s": 1e12, "Mbyte/s": 1e6,
         "Ghz": 1e9, "Mhz": 1e6, "hz": 1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--objects", type=int, required=True)
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--note", default="")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    raw = subprocess.check_output(["ncu", "-i", a.rep, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    caps = []
    for vals in rows[2:]:
        d = {"workload": a.workload, "objects": a.objects, "views": a.views, "note": a.note,
             "source": os.path.basename(a.rep), "report": os.path.basename(a.rep).replace(".ncu-rep", "")}
        for i, h in enumerate(hdr):
            if h == "Kernel Name":
                d["kernel"] = vals[i]
            if h in KEYS:
                try:
                    x = float(vals[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                d[KEYS[h]] = x * SCALE.get(u, 1)
                if KEYS[h] == "duration":
                    d["duration_ms"] = d.pop("duration")
        caps.append(d)
    out = {"captures": []}
    if os.path.exists(a.out):
        out = json.load(open(a.out))
    out["captures"] = [c for c in out["captures"] if not (c["source"] == os.path.basename(a.rep))] + caps
    json.dump(out, open(a.out, "w"), indent=1)
    for c in caps:
        print(json.dumps(c))


if __name__ == "__main__":
    main()
