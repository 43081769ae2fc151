#!/bin/bash
# One gpurun call (1 GPU) that captures the ncu evidence bench.py's roofline block and DESIGN.md cite:
#   gpurun --timeout 1500 -- 'bash tools/ncu_capture.sh r02'
# 1. launch list of one full train step (time + DRAM bytes per launch)  -> gpurun_out/<tag>_launches.csv -> tools/ncu_summary.py
# 2. `--set full` captures of the forward GEMMs (bias / bias+GELU dual store / dropout+residual), the lm_head LSE pass, the dgrad / wgrad GEMMs,
#    attention forward / backward, LayerNorm forward / backward: one short bench process per kernel family (ncu replays each launch ~40 x).
set -u
TAG=${1:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BENCH="python bench.py --profile-mode --blocks none --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    $BENCH --steps 1 --warmup 1 > gpurun_out/${TAG}_launches_bench.json 2> gpurun_out/${TAG}_launches_bench.err
python tools/ncu_summary.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_ncu_launches_summary.txt gpurun_out/${TAG}_ncu_traffic.json > /dev/null 2>&1
full() {  # name, kernel regex (demangled), launches to capture, launches to skip
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $4 -c $3 -f -o gpurun_out/${TAG}_full_$1 \
      $BENCH --steps 1 --warmup 0 --samples 16 > /dev/null 2> gpurun_out/${TAG}_full_$1.err
  ncu -i gpurun_out/${TAG}_full_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_$1.csv 2>/dev/null
  if [ "${5:-}" = "source" ]; then ncu -i gpurun_out/${TAG}_full_$1.ncu-rep --page source --csv > gpurun_out/${TAG}_full_$1_source.csv 2>/dev/null; fi
  rm -f gpurun_out/${TAG}_full_$1.ncu-rep      # gpurun brings back at most 64 MiB: keep the CSV exports, not the reports
}
full gemm_fwd   'gemm_kernel<(\(int\))?0, (\(int\))?0, (\(int\))?(0|6),' 12 13 source
full gemm_lse   'gemm_kernel<(\(int\))?0, (\(int\))?0, (\(int\))?(2|4),' 2 0 source
full gemm_dgrad 'gemm_kernel<(\(int\))?0, (\(int\))?1, ' 8 2
full gemm_wgrad 'gemm_kernel<(\(int\))?1, (\(int\))?1, ' 5 0
full attn       'attn_packed' 4 4
full ln         'layernorm_' 6 12
python tools/ncu_full_summary.py gpurun_out ${TAG} > gpurun_out/${TAG}_ncu_full_summary.csv 2> gpurun_out/${TAG}_ncu_full_summary.err
ls -la gpurun_out | head -40
