"""-m gpu, needs >= 2 GPUs on the box (skipped on a single-GPU box): the fused data-parallel optimizer step over NVLink peer memory
(reduce-scatter + AdamW + all-gather in one kernel) against the NCCL all-reduce path. The single-GPU emulation of its peer-pointer path
lives in tests/test_kernels_gpu.py::test_adamw_dp_peer_path_emulated."""
import os
import subprocess
import sys

import pytest
import torch

from _util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("te", ["0", "1"])
@pytest.mark.parametrize("multicast", ["1", "0"])
def test_fused_step_matches_nccl(te, multicast):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, DP_TEST_TE=te, CLIPDLM_DP_MULTICAST=multicast)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           str(29611 + int(te) * 2 + int(multicast)), os.path.join(ROOT, "tests", "_dp_fused_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    if r.returncode != 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"dp_fused_te{te}_mc{multicast}.log"), "w") as f:
            f.write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    tb = [l for l in (r.stdout + r.stderr).splitlines() if "Error" in l or "assert" in l]
    assert r.returncode == 0, "\n".join(tb[-12:])
    assert r.stdout.count("DP_FUSED_OK") == 2, r.stdout[-2000:]
    if multicast == "0":
        assert "multicast=False" in r.stdout


@pytest.mark.parametrize("fused", ["1", "0"])
def test_data_parallel_equals_single_device(fused):
    """2 ranks x B captions == 1 device x 2B captions (same t, same per-caption noise): losses and post-step weights, both exchanges."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, DP_TEST_FUSED=fused)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           str(29621 + int(fused)), os.path.join(ROOT, "tests", "_dp_equiv_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dp_equiv_fused{fused}.log"), "w") as f:
        f.write(r.stdout + "\n---- stderr ----\n" + r.stderr[-4000:])
    tb = [l for l in (r.stdout + r.stderr).splitlines() if "Error" in l or "assert" in l]
    assert r.returncode == 0, "\n".join(tb[-12:])
    assert r.stdout.count("DP_EQUIV_OK") == 2, r.stdout[-2000:]
