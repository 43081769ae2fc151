#!/bin/bash
# Same-box A/B of two builds of libclipdlm.so (interleaved, the in-tree library is restored at the end):
#   gpurun --timeout 900 -- 'bash tools/ab_library.sh build_ab/libclipdlm_r02base.so [rounds]'
set -u
cd "$(dirname "$0")/.."
OLD=$1; ROUNDS=${2:-2}
LIB=diffusion-image-captioning_b200/libclipdlm.so
mkdir -p gpurun_out
cp $LIB /tmp/libclipdlm_new.so
for r in $(seq 1 $ROUNDS); do
  for tag in new old; do
    if [ $tag = old ]; then cp $OLD $LIB; else cp /tmp/libclipdlm_new.so $LIB; fi
    timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --blocks none > gpurun_out/ab_${tag}_$r.json 2> gpurun_out/ab_${tag}_$r.err
    timeout 300 python bench.py --workload denoise --steps 3 --warmup 3 --no-cpu-baseline --blocks none > gpurun_out/ab_denoise_${tag}_$r.json 2>> gpurun_out/ab_${tag}_$r.err
  done
done
cp /tmp/libclipdlm_new.so $LIB
python - <<PY
import json
for r in range(1, $ROUNDS + 1):
    for tag in ("new", "old"):
        try:
            d = json.loads(open(f"gpurun_out/ab_{tag}_{r}.json").read().strip().splitlines()[-1])
            k = d["kernels"]
            print(tag, r, "train", round(d["ms_per_step"], 2), "ms/step;", {a: round(k[a]["ms_per_step"], 2) for a in ("gemm_fwd", "gemm_dgrad", "gemm_wgrad", "gemm_lse", "ln_fwd", "ln_bwd", "attn_fwd", "attn_bwd")}, "clk", d["clocks"]["sm_mhz"])
            d = json.loads(open(f"gpurun_out/ab_denoise_{tag}_{r}.json").read().strip().splitlines()[-1])
            print(tag, r, "denoise", round(d["value"], 1), d["unit"], round(d.get("ms_per_step", 0), 2), "ms")
        except Exception as ex:
            print(tag, r, "no result:", ex)
PY
