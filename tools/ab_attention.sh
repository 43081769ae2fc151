#!/bin/bash
# Same-box A/B of attention kernels across library builds (interleaved):  gpurun -- 'bash tools/ab_attention.sh build_ab/a.so build_ab/b.so ...'
set -u
cd "$(dirname "$0")/.."
LIB=diffusion-image-captioning_b200/libclipdlm.so
mkdir -p gpurun_out
cp $LIB /tmp/libclipdlm_keep.so
: > gpurun_out/ab_attention.log
for r in 1 2 3; do
  for so in "$@"; do
    cp $so $LIB
    echo "== $so round $r" | tee -a gpurun_out/ab_attention.log
    ATTN_PATHS=0 timeout 120 python tools/attn_perf.py 2>&1 | tee -a gpurun_out/ab_attention.log
    ROWS=1024 ATTN_PATHS=0 timeout 120 python tools/attn_perf.py 2>&1 | sed 's/^/  [R=1024] /' | tee -a gpurun_out/ab_attention.log
  done
done
cp /tmp/libclipdlm_keep.so $LIB
