"""-m gpu: the C-ABI driven from a plain C host (examples/c_host.c): one train step + a denoise loop, no Python in that process."""
import subprocess

import pytest

from test_host_cpu import _build_c_host

pytestmark = pytest.mark.gpu


def test_c_host_runs(tmp_path):
    exe = str(tmp_path / "c_host")
    r = _build_c_host(exe)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "c_host ok" in run.stdout and "train step: x_t_loss" in run.stdout
