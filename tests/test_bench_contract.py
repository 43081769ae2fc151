"""bench.py contract, the part that runs without a GPU: the reference arm (`--impl reference`, the CPU path through the oracle port)
prints exactly ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0", "--ref-samples", "3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "training samples/sec (seq=16)" and d["higher_is_better"] is True
    for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    assert d["unit"] == d["e2e"]["unit"] == d["cpu_baseline"]["unit"] == "captions/s"   # the two arms must print the same unit string
    assert d["warmup"] >= 2 and d["steps"] == 2 and "median" in d["sample"]
    # the config block is a function of the command line only: our arm prints the same one (same_config for the driver)
    import argparse, importlib.util
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": workload_config(args)') == 2


def test_algorithmic_flop_constants_follow_the_survey_formula():
    """bench.py's TFLOP/s figures rest on SURVEY.md 8(d): per layer fwd = 8 L d^2 + 4 L^2 d + 4 L d F, head = 2 L d^2, CLIP linears =
    4 * 512 * d, lm_head = 2 L_txt d V; training = 3 x the trainable part (fwd + dgrad + wgrad) + 2 x the frozen lm_head (fwd + dgrad)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    src = open(spec.origin).read()
    ns = {}
    for name in ("TRAIN_GFLOP_PER_ROW", "FWD_GFLOP_PER_ROW_NO_HEAD", "LM_HEAD_GFLOP_PER_ROW"):
        line = next(l for l in src.splitlines() if l.startswith(name + " ="))
        exec(line.split("#")[0], ns)   # the constants only: importing bench.py would redirect this process's stdout

    def flops(nl, d, f, L, ltxt, v=30522, clip=512):
        body = nl * (8 * L * d * d + 4 * L * L * d + 4 * L * d * f) + 2 * L * d * d + 4 * clip * d
        lm = 2 * ltxt * d * v
        return (3 * body + 2 * lm) / 1e9, body / 1e9, lm / 1e9

    for key, cfg in ((6, (6, 768, 3072, 18, 16)), (12, (12, 768, 3072, 18, 16)), ("bert-large", (24, 1024, 4096, 66, 64))):
        train, body, lm = flops(*cfg)
        assert abs(ns["TRAIN_GFLOP_PER_ROW"][key] - train) < 5e-4 * train, key
        if key in ns["FWD_GFLOP_PER_ROW_NO_HEAD"]:
            assert abs(ns["FWD_GFLOP_PER_ROW_NO_HEAD"][key] - body) < 5e-4 * body
    assert abs(ns["LM_HEAD_GFLOP_PER_ROW"] - flops(6, 768, 3072, 18, 16)[2]) < 1e-4
