#!/usr/bin/env python
"""Summarise an `ncu --csv --log-file` launch list (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum)
of `bench.py --profile-mode` into per-kernel-category numbers:

    python tools/ncu_summary.py gpurun_out/launches.csv profiles/r01_ncu_launches_summary.txt profiles/ncu_traffic.json

* share of the step per category (ncu times are cold-cache and serialised: compare SHARES with bench.py's live event times);
* DRAM traffic per launch (read + write) per category -> `roofline.traffic` of bench.py (profiles/ncu_traffic.json).
Categories follow the engine profiler (include/clipdlm.h CLIPDLM_PROF_*)."""
import collections
import csv
import json
import re
import sys


def category(name: str) -> str:
    m = re.search(r"gemm_kernel<\(?(?:int\))?(\d), \(?(?:int\))?(\d), \(?(?:int\))?(\d)", name)
    if m:
        amaj, bmaj, epi = (int(x) for x in m.groups())
        return {1: "gemm_wgrad", 2: "gemm_lse", 3: "gemm_smgrad", 4: "gemm_lse"}.get(epi, "gemm_dgrad" if bmaj else "gemm_fwd")
    name = name.replace("(bool)", "")
    name = name.replace("(int)", "")
    for key, cat in (("softmax_grad_inplace", "gemm_smgrad"), ("attn_packed_kernel<0", "attn_fwd"), ("attn_packed_kernel<1", "attn_bwd"),
                     ("ce_row_terms", "loss"), ("attn_umma_kernel<0", "attn_fwd"), ("attn_umma_kernel<1", "attn_bwd"),
                     ("attn_ring_kernel<0", "attn_fwd"), ("attn_ring_kernel<1", "attn_bwd"), ("attn_fwd", "attn_fwd"),
                     ("attn_bwd", "attn_bwd"), ("layernorm_fwd", "ln_fwd"), ("layernorm_bwd", "ln_bwd"), ("embed_fwd", "embed"),
                     ("embed_loss", "loss"), ("lse_combine", "loss"), ("colsum", "colsum"), ("adamw", "adamw")):
        if key in name:
            return cat
    return "other (clipdlm)" if "clipdlm::" in name else "torch"


def main():
    src, out_txt, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.defaultdict(lambda: dict(n=0, ms=0.0, rd=0.0, wr=0.0))
    launches = collections.defaultdict(dict)
    if lines and lines[0].startswith("id,kernel,time_us"):   # the compact per-launch table kept under profiles/ (one row per launch)
        for row in csv.DictReader(lines):
            launches[row["id"]] = {"name": row["kernel"], "gpu__time_duration.sum": float(row["time_us"]) * 1e-3,
                                   "dram__bytes_read.sum": float(row["dram_read_mb"]) * 1e6, "dram__bytes_write.sum": float(row["dram_write_mb"]) * 1e6}
        lines = []
    for row in csv.DictReader(lines):
        launches[row["ID"]]["name"] = row["Kernel Name"]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        metric = row["Metric Name"]
        if metric == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(unit, 1e-6)
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        launches[row["ID"]][metric] = v
    for rec in launches.values():
        c = per[category(rec["name"])]
        c["n"] += 1
        c["ms"] += rec.get("gpu__time_duration.sum", 0.0)
        c["rd"] += rec.get("dram__bytes_read.sum", 0.0)
        c["wr"] += rec.get("dram__bytes_write.sum", 0.0)
    tot = sum(c["ms"] for c in per.values())
    with open(out_txt, "w") as f:
        f.write(f"# {src}: {sum(c['n'] for c in per.values())} launches, {tot:.1f} ms of kernel time under ncu (cold cache, serialised)\n")
        f.write(f"{'category':18s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'DRAM MB/launch':>15s} {'GB/s':>8s}\n")
        for k, c in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
            mb = (c["rd"] + c["wr"]) / max(c["n"], 1) / 1e6
            gbs = (c["rd"] + c["wr"]) / max(c["ms"], 1e-9) / 1e6
            f.write(f"{k:18s} {c['n']:8d} {c['ms']:9.2f} {100 * c['ms'] / tot:6.1f}% {mb:15.1f} {gbs:8.0f}\n")
    json.dump({k: dict(launches=c["n"], ms=c["ms"], share=c["ms"] / tot, dram_bytes_per_launch=(c["rd"] + c["wr"]) / max(c["n"], 1))
               for k, c in per.items()}, open(out_json, "w"), indent=1)
    print(open(out_txt).read())


if __name__ == "__main__":
    main()
