#!/bin/bash
# The two-GPU fused-vs-NCCL optimizer-step check of tests/test_dp_fused_gpu.py without pytest (one torchrun per exchange path):
#   gpurun --gpus 2 -- 'bash tools/dp_fused_quick.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r02_dp_fused_workers.log
for mc in 0 1; do
  DP_TEST_VERBOSE=1 DP_TEST_TE=0 CLIPDLM_DP_MULTICAST=$mc timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$mc tests/_dp_fused_worker.py > /tmp/w.log 2>&1
  echo "== TRAIN_EMBEDDING=0 multicast=$mc rc=$?" | tee -a gpurun_out/r02_dp_fused_workers.log
  grep -h "DP_FUSED_OK\|AssertionError" /tmp/w.log | cut -c1-300 | tee -a gpurun_out/r02_dp_fused_workers.log
done
