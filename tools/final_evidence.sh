#!/bin/bash
# One gpurun call (1 GPU) that refreshes the round's record with the final kernels:  gpurun --timeout 1500 -- 'bash tools/final_evidence.sh r02'
# GPU test suite, smoke(), compute-sanitizer (memcheck + synccheck) on the C host, the driver-style default bench line.
set -u
TAG=${1:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
SANITIZE_TOOLS="memcheck synccheck" SANITIZE_DROPOUTS="0.1" SANITIZE_TIMEOUT=60 bash tools/sanitize_c_host.sh gpurun_out/${TAG}_sanitizer_memcheck_synccheck.log > /dev/null 2>&1
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1_final.json 2> gpurun_out/${TAG}_bench_n1_final.err
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log; grep -c "ERROR SUMMARY: 0 errors" gpurun_out/${TAG}_sanitizer_memcheck_synccheck.log; grep "exit code" gpurun_out/${TAG}_sanitizer_memcheck_synccheck.log
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n1_final.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], d["roofline"]["frac"], d["clocks"])
for b in ("parity_mode", "denoise", "denoise_b8", "layers12"):
    print(b, {k: v for k, v in d.get(b, {}).items() if k in ("value", "ms_per_step", "ms_per_loop", "ms_per_batch")})
PY
