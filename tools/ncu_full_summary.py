#!/usr/bin/env python
"""Condenses the `ncu --page raw --csv` exports of tools/ncu_capture.sh into one table (one row per captured launch):
    python tools/ncu_full_summary.py gpurun_out r02 > profiles/r02_ncu_full_summary.csv"""
import csv
import glob
import os
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_mem_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_smem_pct"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__cycles_active.avg", "sm_cycles_active"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clock"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
]


def main():
    d, tag = sys.argv[1], sys.argv[2]
    w = csv.writer(sys.stdout)
    w.writerow(["capture", "id", "kernel", "grid", "block"] + [m[1] for m in METRICS] + [m[1] + "_unit" for m in METRICS if m[1] in ("time", "dram_read", "dram_write", "sm_clock")])
    for path in sorted(glob.glob(os.path.join(d, f"{tag}_full_*.csv"))):
        if path.endswith("_source.csv"):   # the per-instruction source-page exports of tools/ncu_capture.sh have another layout
            continue
        rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            vals = [r[col[m]] if m in col else "" for m, _ in METRICS]
            us = [units[col[m]] if m in col else "" for m, n in METRICS if n in ("time", "dram_read", "dram_write", "sm_clock")]
            w.writerow([os.path.basename(path)[len(tag) + 6:-4], r[col["ID"]], r[col["Kernel Name"]][:110], r[col["Grid Size"]], r[col["Block Size"]]] + vals + us)


if __name__ == "__main__":
    main()
