#!/bin/bash
# One gpurun call that validates and measures the two experimental paths of DESIGN.md section 8 (factored softmax-CE gradient,
# gelu'(u) stored by the forward):
#   gpurun --timeout 600 -- 'bash tools/validate_experimental.sh'
# 1. the experimental GPU tests; 2. the WHOLE GPU suite with the path switched on through the environment; 3. bench A/B on the same box.
# Everything lands in gpurun_out/fsg_*.  Make it the default (model.py: fused_softmax_grad default) only if 1 and 2 are green and 3 is a gain.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CLIPDLM_TEST_EXPERIMENTAL=1
python -m pytest tests/test_experimental_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/fsg_tests.log
CLIPDLM_FUSED_SOFTMAX_GRAD=1 CLIPDLM_GELU_DERIV_STORE=2 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/fsg_full_suite_switched_on.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/fsg_bench_default.json 2> gpurun_out/fsg_bench_default.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fused-softmax-grad > gpurun_out/fsg_bench_fused.json 2> gpurun_out/fsg_bench_fused.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gelu-deriv-store > gpurun_out/fsg_bench_gelud.json 2> gpurun_out/fsg_bench_gelud.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gelu-deriv-store 2 > gpurun_out/fsg_bench_gelud2.json 2> gpurun_out/fsg_bench_gelud2.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fused-softmax-grad --gelu-deriv-store 2 > gpurun_out/fsg_bench_both.json 2> gpurun_out/fsg_bench_both.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/fsg_bench_default_again.json 2>> gpurun_out/fsg_bench_default.err
python - <<'PY'
import json
for n in ("default", "fused", "gelud", "gelud2", "both", "default_again"):
    try:
        d = json.load(open(f"gpurun_out/fsg_bench_{n}.json"))
        k = d["kernels"]
        print(n, round(d["ms_per_step"], 2), "ms/step", round(d["value"], 1), "captions/s; lse", round(k["gemm_lse"]["ms_per_step"], 2),
              "smgrad", round(k.get("gemm_smgrad", {}).get("ms_per_step", 0.0), 2), "fwd", round(k["gemm_fwd"]["ms_per_step"], 2), "dgrad", round(k["gemm_dgrad"]["ms_per_step"], 2),
              "loss", round(k["loss"]["ms_per_step"], 2), "colsum", round(k["colsum"]["ms_per_step"], 2))
    except Exception as ex:
        print(n, "no result:", ex)
PY
# parity-mode (bf16x3) step time on the same box (VERDICT r1 missing #3)
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precision bf16x3 > gpurun_out/fsg_bench_bf16x3.json 2> gpurun_out/fsg_bench_bf16x3.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/fsg_bench_bf16x3.json"))
    print("bf16x3", round(d["ms_per_step"], 2), "ms/step", round(d["value"], 1), "captions/s", {k: round(v["ms_per_step"], 1) for k, v in d["kernels"].items()})
except Exception as ex:
    print("bf16x3 no result:", ex)
PY
