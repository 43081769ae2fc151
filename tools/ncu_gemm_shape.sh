#!/bin/bash
# ncu --set full + source view of single tools/gemm_perf.py shapes:  gpurun -- 'bash tools/ncu_gemm_shape.sh tag1 "pattern 1" tag2 "pattern 2" ...'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  tag=$1; pat=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/ncu_$tag \
      python tools/gemm_perf.py --rows 8192 --only "$pat" --flags 0 --rounds 1 --no-cublas > gpurun_out/ncu_$tag.log 2>&1
  ncu -i gpurun_out/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_${tag}_source.csv 2>/dev/null
  rm -f gpurun_out/ncu_$tag.ncu-rep
done
