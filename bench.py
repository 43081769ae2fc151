#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: training samples/sec of the CLIP-Diffusion-LM train step
(CLIP-DDPM.py train_func: q_sample -> 2 encoder passes -> L1 + rounding-CE loss -> backward -> AdamW), synthetic CLIP features
and random token ids, seq_len 16, bs = 512 captions per GPU x SAMPLE_SIZE = 100 noise levels (51 712 encoder rows per step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|denoise] [--layers 6]

One JSON line on stdout (rank 0). `value` = whole-job captions/s with inputs resident in HBM; `e2e` = the same through the public
train_func() call with the batch copied from pinned host memory and the loss read back every step; `roofline` = the dominant
kernel (tcgen05 GEMM) timed live with CUDA events inside the timed steps; `cpu_baseline` = the oracle port of the reference's CPU
path on this box's host cores (bounded sample). `--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import contextlib
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line. Libraries print there too (the image sets NCCL_DEBUG=VERSION: "NCCL version ..." comes from
# NCCL's C code), so file descriptor 1 is pointed at stderr for the whole run and the result line is written to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import torch  # noqa: E402

TRAIN_GFLOP_PER_ROW = {6: 6.1730, 12: 10.7774, "bert-large": 129.2953}  # algorithmic 2*MAC, fwd+dgrad+wgrad, frozen lm_head x2, no recompute
FWD_GFLOP_PER_ROW_NO_HEAD = {6: 1.5576, 12: 3.0924}
LM_HEAD_GFLOP_PER_ROW = 0.7501


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def ncu_traffic(category: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py from the .ncu-rep brought back from the GPU box)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        return d.get(category, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_host_batch(B: int, seed: int, ML: int = 16):
    g = torch.Generator().manual_seed(seed)
    b = {"input_ids": torch.randint(0, 30522, (B, ML), generator=g), "attention_mask": torch.ones(B, ML, dtype=torch.int64),
         "image_clip": torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1),
         "text_clip": torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1)}
    return {k: v.pin_memory() if torch.cuda.is_available() else v for k, v in b.items()}


# ------------------------------------------------------------------------------------------------------------ CPU (reference) arm
def cpu_reference_steps(layers: int, steps: int, warmup: int, B: int = 8, S: int = 24, workload: str = "train", device: str = "cpu",
                        autocast: bool = False):
    """The reference's CPU path through the oracle port (the reference is Python + HF transformers and cannot travel to the
    GPU box; oracle/clipdlm_oracle.py restates it op for op and is pinned against it). fp32, train mode (dropout on), AdamW step
    included, all host threads. Each step is a bounded sample: B captions x S noise levels (+ the x_1 pass)."""
    from oracle import clipdlm_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    hp = O.default_hparams()
    hp.update(BATCH_SIZE=B, SAMPLE_SIZE=S, N_LAYERS=layers)
    P = {k: v.to(device) for k, v in O.init_params(hp, seed=0).items()}
    acp = O.alpha_cumprod(hp).to(device)
    batch = {k: v.to(device) for k, v in O.synthetic_batch(hp, seed=0).items()}
    times = []
    on_gpu = device != "cpu"  # --ref-device cuda: the same eager PyTorch ops on the B200 itself, optionally under autocast(bf16)

    def clock():
        if on_gpu:
            torch.cuda.synchronize()
        return time.perf_counter()

    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if (on_gpu and autocast) else contextlib.nullcontext
    if workload == "train":
        opt = O.AdamW(O.make_trainable(P, hp), lr=1e-4)
        for i in range(warmup + steps):
            t0 = clock()
            with ctx():
                O.train_func(P, opt, batch, hp, acp, True)
            t1 = clock()
            if i >= warmup:
                times.append(t1 - t0)
        rows = B * (S + 1)
    else:
        for i in range(warmup + steps):
            t0 = clock()
            with ctx():
                O.sample(P, batch["image_clip"], hp, n_steps=S)
            t1 = clock()
            if i >= warmup:
                times.append(t1 - t0)
        rows = B
    return times, rows, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = min(args.warmup, 1)
    gpu_ref = args.ref_device == "cuda"
    if gpu_ref:
        return run_reference_on_gpu(args)
    if args.workload == "train":
        B, S = 8, 24
        times, rows, threads = cpu_reference_steps(args.layers, args.steps, W, B, S, "train")
        sec = sum(times)
        rows_per_s = rows * len(times) / sec
        value = rows_per_s / 101.0  # captions/s of the full workload: each caption = SAMPLE_SIZE + 1 = 101 encoder rows
        sample = f"{len(times)} train steps of {B} captions x {S}+1 noise levels ({rows} encoder rows/step), 6-layer fp32 oracle port, dropout on, AdamW"
        metric, unit = "training samples/sec (seq=16)", "captions/s (1 caption = 101 noised sequences)"
    else:
        B, S = 8, 10
        times, rows, threads = cpu_reference_steps(args.layers, args.steps, W, B, S, "denoise")
        sec = sum(times)
        value = rows * len(times) / sec * (S / 100.0)  # captions/s at 100 denoise steps
        sample = f"{len(times)} denoise loops of {B} captions x {S} steps (lm_head every step, as the reference), scaled to 100 steps"
        metric, unit = "denoise-loop captions/sec (100 steps)", "captions/s"
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": len(times), "warmup": W,
            "ms_per_step": 1e3 * sec / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, B_override=None),
            "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def run_reference_on_gpu(args):
    """Extra comparison, not the driver's reference arm: the reference's own eager PyTorch op sequence (the oracle port) executed on
    the B200 (`--impl reference --ref-device cuda [--ref-autocast]`), fp32 or autocast(bf16), train mode, AdamW step included.
    The reference materialises fp32 logits [rows, 16, 30522] twice (lm_head output + softmax), so its batch is capped by memory:
    --ref-batch captions x --ref-samples noise levels per step (default 8 x 100, CLIP-DDPM.py's own defaults)."""
    B, S = args.ref_batch, args.ref_samples
    W = max(1, min(args.warmup, 2))
    if args.workload == "train":
        times, rows, _ = cpu_reference_steps(args.layers, args.steps, W, B, S, "train", "cuda", args.ref_autocast)
        sec = sum(times)
        value = rows * len(times) / sec / (S + 1.0)
        metric, unit = "training samples/sec (seq=16)", f"captions/s (1 caption = {S + 1} noised sequences)"
        what = f"train step, {B} captions x {S}+1 noise levels = {rows} encoder rows"
    else:
        times, rows, _ = cpu_reference_steps(args.layers, args.steps, W, B, S, "denoise", "cuda", args.ref_autocast)
        sec = sum(times)
        value = rows * len(times) / sec
        metric, unit = f"denoise-loop captions/sec ({S} steps)", "captions/s"
        what = f"denoise loop, {B} captions x {S} steps, lm_head every step"
    emit({"impl": "reference", "ref_device": "cuda", "ref_dtype": "autocast-bf16" if args.ref_autocast else "f32", "metric": metric, "value": value,
          "unit": unit, "n_gpus": 1, "steps": len(times), "warmup": W, "ms_per_step": 1e3 * sec / len(times), "higher_is_better": True,
          "data": "synthetic", "config": {"workload": f"eager PyTorch restatement of CLIP-DDPM.py on the B200: {what}, {args.layers}L", "global_batch": B,
                                           "sample_size": S, "layers": args.layers},
          "max_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30})


def workload_config(args, B_override=None):
    B = args.batch
    if getattr(args, "model", "distilbert") == "bert-large":
        return {"workload": f"CLIP-DDPM.py train_func on a bert-large-shaped encoder (24L/1024/16H/4096), seq_len=64 (+2 CLIP positions), bs={B} captions/GPU x "
                            f"SAMPLE_SIZE={args.samples} (+x_1 pass) = {B * (args.samples + 1)} encoder rows/step/GPU, x_0-predict, concat fusion, L1 + rounding CE, "
                            f"dropout 0.1, AdamW (BASELINE.json configs[4])", "global_batch": B * args.gpus, "seq_len": 64, "sample_size": args.samples,
                "layers": 24, "parallelism": f"dp{args.gpus}", "chunk_rows": args.chunk_rows, "l2": "working set >> L2, 256 MB L2 flush between timed steps"}
    if args.workload == "train" and getattr(args, "train_embedding", False):
        return {"workload": f"CLIP-DDPM.py train_func with TRAIN_EMBEDDING=True (IN_CHANNEL=16 learned embedding, trainable lm_head + in/out projections), "
                            f"{args.layers}L/768, seq_len=16, bs={B} x SAMPLE_SIZE={args.samples}", "global_batch": B * args.gpus, "seq_len": 16,
                "sample_size": args.samples, "layers": args.layers, "parallelism": f"dp{args.gpus}", "chunk_rows": args.chunk_rows,
                "l2": "working set >> L2, 256 MB L2 flush between timed steps"}
    if args.workload == "train":
        return {"workload": f"CLIP-DDPM.py train_func, DistilBertConfig() {args.layers}L/768/12H/3072 ('bert-base' in BASELINE.json), seq_len=16 (+2 CLIP positions), "
                            f"bs={B} captions/GPU x SAMPLE_SIZE={args.samples} (+x_1 pass) = {B * (args.samples + 1)} encoder rows/step/GPU, x_0-predict, concat fusion, "
                            f"L1 + rounding CE, dropout 0.1, AdamW", "global_batch": B * args.gpus, "seq_len": 16, "sample_size": args.samples,
                "layers": args.layers, "parallelism": f"dp{args.gpus}", "chunk_rows": args.chunk_rows,
                "l2": "per-step working set (>10 GB of activations) >> 126 MB L2, plus a 256 MB L2 flush between timed steps"}
    return {"workload": f"CLIP-DDPM.py denoise loop :611-621, {args.layers}L model, bs={args.denoise_batch}/GPU, n_steps={args.denoise_steps}, eval mode, "
                        f"fused lm_head+argmax on the last step", "global_batch": args.denoise_batch * args.gpus, "seq_len": 16, "layers": args.layers,
            "parallelism": f"shard{args.gpus}", "l2": "256 MB L2 flush between timed loops"}


# ------------------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import clipdlm
    from clipdlm import parallel
    rank, local_rank, world = parallel.init_process_group_from_env("nccl")
    assert world == args.gpus or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, S = args.batch, args.samples
    hp = clipdlm.default_hparams(BATCH_SIZE=B, SAMPLE_SIZE=S, N_LAYERS=args.layers)
    if args.model == "bert-large":  # BASELINE.json configs[4]: bert-large-shaped 24L/1024/16H/4096, seq_len 64 (+2 CLIP positions)
        hp = clipdlm.default_hparams(BATCH_SIZE=B, SAMPLE_SIZE=S, N_LAYERS=24, DIM=1024, N_HEADS=16, HIDDEN_DIM=4096, MAX_LENGTH=64)
    if args.train_embedding:  # CLIP-DDPM.py:98-102: 16-channel learned embedding, trainable lm_head and in/out projections
        hp.update(TRAIN_EMBEDDING=True, IN_CHANNEL=16)
    torch.manual_seed(0)
    model = clipdlm.DistilBertModel(None, None, None, hp=hp, precision=args.precision, seed=0, chunk_rows=args.chunk_rows,
                                    fused_softmax_grad=True if args.fused_softmax_grad else None,
                                    gelu_deriv_store=args.gelu_deriv_store if args.gelu_deriv_store else None)
    dp_mode = "single GPU"
    if world > 1:
        dp_mode = "nccl all-reduce of the flat fp32 gradients + AdamW on every rank"
        if args.workload == "train" and args.dp_exchange in ("auto", "fused"):
            try:
                parallel.enable_data_parallel(model, fused=True)
                dp_mode = ("fused reduce-scatter + AdamW + all-gather kernel over " +
                           ("NVSwitch multicast (multimem.ld_reduce / multimem.st)" if model.dp_fused["multicast"] else "NVLink peer loads / stores"))
            except Exception as ex:  # symmetric memory unavailable on this box: the NCCL exchange is the other GPU path, not a CPU fallback
                if args.dp_exchange == "fused":
                    raise
                sys.stderr.write(f"bench: fused data-parallel step unavailable ({type(ex).__name__}: {ex}); using the NCCL all-reduce\n")
                model.dp_fused = None
                parallel.enable_data_parallel(model, fused=False)
        else:
            parallel.enable_data_parallel(model, fused=False)
    trainer = clipdlm.AdamW(model.parameters(), lr=hp["LEARNING_RATE"])
    host = synthetic_host_batch(B if args.workload == "train" else args.denoise_batch, seed=rank, ML=hp["MAX_LENGTH"])
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    if args.workload == "train":
        model.train()
        dev_batch = {k: v.to(dev) for k, v in host.items()}

        def step_resident():
            return clipdlm.train_func(model, trainer, dev_batch)[0]

        def step_e2e():
            b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            return clipdlm.train_func(model, trainer, b)[0].item()  # loss read back: device -> host every step
        units = B
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = 4
    else:
        model.eval()
        img_dev = host["image_clip"].to(dev)

        def step_resident():
            return clipdlm.sample(model, img_dev, n_steps=args.denoise_steps)[0]

        def step_e2e():
            return clipdlm.sample(model, host["image_clip"].to(dev, non_blocking=True), n_steps=args.denoise_steps)[0].cpu()
        units = args.denoise_batch
        h2d = host["image_clip"].numel() * 4
        d2h = units * 16 * 8

    def timed(fn, n, profile=False):
        evs = []
        barrier()
        if profile:
            model.profile(True)
            model.profile_read(reset=True)
        launches0 = model.launch_count()
        t_wall = time.perf_counter()
        for _ in range(n):
            flush_buf.zero_()  # L2 flush, outside the per-step event pair
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        wall = time.perf_counter() - t_wall
        ms = sum(a.elapsed_time(b) for a, b in evs)
        prof = model.profile_read(reset=True) if profile else None
        if profile:
            model.profile(False)
        launches = model.launch_count() - launches0
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, prof, launches, wall

    n_warm = args.warmup if args.profile_mode else max(args.warmup, 3)
    for _ in range(n_warm):
        step_resident()
    if args.profile_mode:   # under ncu: one pass, nothing else
        with ClockSampler(local_rank) as clk:
            ms, prof, launches, wall = timed(step_resident, args.steps, profile=True)
        clocks = clk.summary()
        ms_e2e, ms_prof = ms, ms
    else:
        # `value`: K steps, nothing between the launches. The per-kernel event brackets of the roofline leg cost a few % (they
        # serialise back-to-back launches), so that leg is a SECOND pass over the same K steps; its own step time is reported too.
        with ClockSampler(local_rank) as clk:
            ms, _, launches, wall = timed(step_resident, args.steps)
        clocks = clk.summary()
        ms_prof, prof, _, _ = timed(step_resident, args.steps, profile=True)
        for _ in range(2):
            step_e2e()
        ms_e2e, _, _, _ = timed(step_e2e, args.steps)
    value = units * world * args.steps / (ms / 1e3)
    e2e_value = units * world * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        pk = peaks()
        gemm_cats = ("gemm_fwd", "gemm_dgrad", "gemm_wgrad", "gemm_lse", "gemm_smgrad")
        dom = max(gemm_cats, key=lambda k: prof[k]["ms"])
        d = prof[dom]
        ach = d["flops"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else 0.0
        kernels = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                       "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None,
                       "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 and v["bytes"] > 0 else None}
                   for k, v in prof.items() if v["launches"]}
        roofline = {"bound": "tensor", "kernel": f"gemm_kernel ({dom}: tcgen05.mma.cta_group::2 256x256x16 CTA pairs, TMA-fed 5-stage ring, fused epilogue with TMA stores)", "achieved": ach,
                    "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"], "peak_kind": f"{pk['source']} sustained bf16 (burst {pk['tf_burst']})",
                    "avg_launch_ms": d["ms"] / max(d["launches"], 1), "flops_per_launch": d["flops"] / max(d["launches"], 1),
                    "traffic": ncu_traffic(dom), "ms_per_step_with_event_brackets": ms_prof / args.steps,
                    "all_gemm_tflops": sum(prof[k]["flops"] for k in gemm_cats) / max(sum(prof[k]["ms"] for k in gemm_cats), 1e-9) / 1e9}
        line = {"metric": "training samples/sec (seq=16)" if args.workload == "train" else "denoise-loop captions/sec (100 steps)",
                "value": value, "unit": "captions/s (1 caption = 101 noised sequences)" if args.workload == "train" else "captions/s",
                "n_gpus": world, "steps": args.steps, "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": dict(workload_config(args), dp_exchange=dp_mode, **({"fused_softmax_grad": True} if model.fused_softmax_grad else {}),
                                               **({"gelu_deriv_store": model.gelu_deriv_store} if model.gelu_deriv_store else {})),
                "e2e": {"value": e2e_value, "unit": "captions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches + (args.steps if args.workload == "train" else 0), "clocks": clocks, "roofline": roofline,
                "kernels": kernels}
        if args.workload == "train":
            rows_s = value * (S + 1)
            alg = rows_s * TRAIN_GFLOP_PER_ROW.get("bert-large" if args.model == "bert-large" else args.layers, float("nan")) / 1e3
            line["noised_sequences_per_s"] = rows_s
            line["algorithmic_tflops"] = alg
            line["algorithmic_frac_of_sustained_peak"] = alg / pk["tf_sustained"] / world
        if world == 1 and not args.no_cpu_baseline and not args.profile_mode:
            if args.workload == "train":
                times, rows, threads = cpu_reference_steps(args.layers, 2, 1, 8, 24, "train")
                v = rows * len(times) / sum(times) / 101.0
                sample = f"2 train steps of 8 captions x 24+1 noise levels ({rows} rows/step) after 1 warm-up, fp32 oracle port, dropout on, AdamW"
            else:
                times, rows, threads = cpu_reference_steps(args.layers, 2, 1, 8, 10, "denoise")
                v = rows * len(times) / sum(times) * (10 / 100.0)
                sample = "2 denoise loops of 8 captions x 10 steps, scaled to 100 steps"
            line["cpu_baseline"] = {"value": v, "unit": line["unit"], "cores": threads, "kind": "port", "sample": sample}
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "denoise"])
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--model", default="distilbert", choices=["distilbert", "bert-large"],
                    help="bert-large = BASELINE.json configs[4] (24L/1024/16H/4096, seq_len 64): use with --batch 64 --chunk-rows 1024")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--samples", type=int, default=100)
    ap.add_argument("--chunk-rows", type=int, default=8192)
    ap.add_argument("--denoise-batch", type=int, default=1024)
    ap.add_argument("--denoise-steps", type=int, default=100)
    ap.add_argument("--train-embedding", action="store_true", help="TRAIN_EMBEDDING=True variant of the train step (use with --no-cpu-baseline)")
    ap.add_argument("--dp-exchange", default="auto", choices=["auto", "fused", "nccl"],
                    help="N > 1 gradient exchange: fused = reduce-scatter + AdamW + all-gather in one kernel over NVLink peer memory; auto = fused, NCCL if unavailable")
    ap.add_argument("--fused-softmax-grad", action="store_true",
                    help="experimental: factored softmax-CE gradient of the lm_head (no in-place pass over the stored logits); default off")
    ap.add_argument("--gelu-deriv-store", type=int, nargs="?", const=1, default=0, choices=[0, 1, 2],
                    help="experimental: lin1 stores gelu'(u), the lin2 gradient GEMM multiplies by it instead of evaluating gelu' "
                         "(2: that GEMM also sums the lin1 bias gradient in its epilogue); default off")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"],
                    help="bf16 = speed mode (BASELINE.json configs[1]); bf16x3 = parity mode (split-bf16 operands, fp32-class results)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"], help="--impl reference: 'cuda' runs the eager PyTorch port on the GPU (extra comparison)")
    ap.add_argument("--ref-autocast", action="store_true", help="--ref-device cuda under torch.autocast(bfloat16)")
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--ref-samples", type=int, default=100)
    ap.add_argument("--profile-mode", action="store_true", help="for runs under ncu: no forced warm-up, no e2e leg, no CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
