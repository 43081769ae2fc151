#!/bin/bash
# compute-sanitizer (memcheck) over the C-ABI hot path with no Python / torch in the process: examples/c_host.c runs one train step
# (embed + q_sample + fusion, 2 encoder layers fwd/bwd, fused lm_head CE, AdamW) and a 3-step denoise loop with the fused arg-max.
# Usage (GPU box): [SANITIZE_TOOLS="memcheck synccheck racecheck initcheck"] [SANITIZE_DROPOUTS="0.1 0"] bash tools/sanitize_c_host.sh [logfile]
# exit code != 0 if a tool reports an error.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
LOG="${1:-$ROOT/gpurun_out/sanitizer_memcheck.log}"
PKG="$ROOT/diffusion-image-captioning_b200"
CUDA=/usr/local/cuda
mkdir -p "$(dirname "$LOG")"
gcc -O2 -std=c99 -Wall -Werror "$ROOT/examples/c_host.c" -I"$ROOT/include" -I"$CUDA/include" -L"$PKG" -lclipdlm -L"$CUDA/lib64" -lcudart -lm \
    -Wl,-rpath,"$PKG" -Wl,-rpath,"$CUDA/lib64" -o /tmp/c_host_sanitize || exit 3
rc=0
: > "$LOG"
for tool in ${SANITIZE_TOOLS:-memcheck}; do
  for p in ${SANITIZE_DROPOUTS:-0.1 0}; do
   for fused in ${SANITIZE_FUSED:-1 0}; do   # 1 = the fused lm_head / gelu' options of the Python host's default path, 0 = the C-ABI defaults
    echo "===== $tool, dropout $p, fused options $fused =====" >> "$LOG"
    C_HOST_FUSED=$fused C_HOST_DROPOUT=$p timeout "${SANITIZE_TIMEOUT:-30}" "$CUDA/bin/compute-sanitizer" --tool "$tool" --error-exitcode 9 --print-limit 10 /tmp/c_host_sanitize >> "$LOG" 2>&1
    r=$?
    echo "exit code $r" >> "$LOG"
    [ $r -ne 0 ] && rc=$r
   done
  done
done
grep -v "^=========     \|^=========$" "$LOG" | tail -n 60
exit $rc
