#!/usr/bin/env python
"""Where the three roles of one gemm_kernel launch wait (ncu source page CSV + raw page CSV of tools/ncu_gemm_shape.sh):
samples at the try_wait branches of the operand-empty (TMA producer), operand-full / accumulator-empty (MMA issuer) and
accumulator-full (epilogue) barriers, next to the expected sample share of the role's warps.   python tools/ncu_gemm_roles.py gpurun_out/ncu_<tag>"""
import csv
import re
import sys


def main(prefix):
    raw = list(csv.reader(open(prefix + "_raw.csv")))
    d = {h: v for h, v in zip(raw[0], raw[2])}
    rows = list(csv.reader(open(prefix + "_source.csv")))
    hdr, data = rows[1], rows[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    total = sum(int(r[isamp] or 0) for r in data)
    print(prefix, "time us", d.get("gpu__time_duration.sum"), "tensor pipe %", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "")[:5],
          "issue %", d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "")[:5], "dram %", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "")[:5],
          "L2 hit %", d.get("lts__t_sector_hit_rate.pct", "")[:5], "regs", d.get("launch__registers_per_thread"), "samples", total)
    # barrier offsets: full 0x38400.., empty +0x28, tfull +0x50, tempty +0x60 (5 stages)
    names = {}
    for i, r in enumerate(data):
        m = re.search(r"TRYWAIT P\d, \[(\w+)\+URZ(\+0x[0-9a-f]+)?\]", r[isrc])
        if not m:
            continue
        off = int(m.group(2)[1:], 16) if m.group(2) else 0
        # samples of the wait = this instruction + the following few (branch on the predicate)
        s = sum(int(x[isamp] or 0) for x in data[i:i + 3])
        names.setdefault(off, [0, 0])
        names[off][0] += s
        names[off][1] += int(r[iex] or 0)
    for off in sorted(names):
        print(f"  try_wait @ +0x{off:x}: samples {names[off][0]:6d}  executions {names[off][1]}")
    idle = sum(int(r[isamp] or 0) for r in data if "UCGABAR_WAIT" in r[isrc])
    print("  idle warps at the final cluster barrier:", idle)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
