#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: training samples/sec of the CLIP-Diffusion-LM train step
(CLIP-DDPM.py train_func: q_sample -> 2 encoder passes -> L1 + rounding-CE loss -> backward -> AdamW), synthetic CLIP features
and random token ids, seq_len 16, bs = 512 captions per GPU x SAMPLE_SIZE = 100 noise levels (51 712 encoder rows per step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|denoise] [--layers 6] [--blocks auto|none|a,b,..]

One JSON line on stdout (rank 0). `value` = whole-job captions/s with inputs resident in HBM; `e2e` = the same through the public
train_func() call with the batch drawn from a pinned-host caption set and the loss read back every step; `roofline` = the dominant
kernel (tcgen05 GEMM) timed live with CUDA events inside the timed steps; `cpu_baseline` = the oracle port of the reference's CPU
path on this box's host cores (bounded sample: the reference's own B = 8 x S = 100 step). `--impl reference` times that CPU path alone.

The same line carries short sub-benchmarks of the other BASELINE.json configurations (`--blocks`, default "auto"):
  parity_mode  the same train step in precision="bf16x3" (the mode that meets the 1e-3 / bit-exact-argmax gate)           N = 1
  denoise      configs[3]: p_sample loop, 100 steps, bs = 1024 per GPU (value, e2e with the ids copied back, roofline)   every N
  denoise_b8   the reference's own evaluation shape (B = 8, 5 steps, CLIP-DDPM.py:613-617): latency per batch            N = 1
  layers12     SURVEY M1's reading of "bert-base" (12-layer encoder), same train step                                    N = 1
  eager_b200   the reference's eager PyTorch op sequence (oracle port) on this B200, fp32 and autocast(bf16), B = 8     N = 1
  bert_large   configs[4]: 24L/1024/16H/4096, seq_len 64, bs = 64 per GPU                                              N = 8
  dp_check     N > 1: weights bit-identical on every rank after the timed steps; fused exchange vs NCCL exchange on one extra step
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
REF_B, REF_S = 8, 100   # CLIP-DDPM.py:57,109 = BASELINE.json configs[0]: the reference's own step (808 encoder rows)


def cpu_reference_steps(layers: int, steps: int, warmup: int, B: int = REF_B, S: int = REF_S, workload: str = "train", device: str = "cpu",
                        autocast: bool = False, budget_s: float = None, min_steps: int = 3):
    """The reference's CPU path through the oracle port (the reference is Python + HF transformers and cannot travel to the
    GPU box; oracle/clipdlm_oracle.py restates it op for op and is pinned against it). fp32, train mode (dropout on), AdamW step
    included, all host threads. One step = the reference's own step shape: B captions x S noise levels (+ the x_1 pass).
    budget_s bounds the wall time: once it is spent and min_steps timed steps exist, the loop stops (the count is reported)."""
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
    opt = O.AdamW(O.make_trainable(P, hp), lr=1e-4) if workload == "train" else None
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = clock()
        with ctx():
            if workload == "train":
                O.train_func(P, opt, batch, hp, acp, True)
            else:
                O.sample(P, batch["image_clip"], hp, n_steps=S)
        t1 = clock()
        if i >= warmup:
            times.append(t1 - t0)
        if budget_s is not None and time.perf_counter() - t_start > budget_s and len(times) >= min_steps:
            break
    rows = B * (S + 1) if workload == "train" else B
    return times, rows, threads


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port, kind "port") on this box's host cores, all threads.
    Same metric / unit / config as our arm; each step is a bounded sample of that workload: the reference's own step (BASELINE.json
    configs[0]: 8 of the captions x all S = 100 (+1) noise levels, fp32, dropout on, AdamW) - value = 8 captions / median step time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device == "cuda":
        return run_reference_on_gpu(args)
    W = max(2, args.warmup)   # drop (at least) 2, median of the rest
    REF_B, REF_S = args.ref_batch, args.ref_samples   # defaults 8 x 100 = the reference's own step; smaller only for the CPU contract test
    if args.workload == "train":
        times, rows, threads = cpu_reference_steps(args.layers, args.steps, W, REF_B, REF_S, "train", budget_s=args.ref_budget_s, min_steps=min(5, args.steps))
        med = statistics.median(times)
        value = REF_B / med
        sample = (f"median of {len(times)} train steps after {W} warm-up steps; one step = {REF_B} captions x {REF_S}+1 noise levels = {rows} encoder rows "
                  f"(CLIP-DDPM.py's own BATCH_SIZE x SAMPLE_SIZE, BASELINE.json configs[0]), {args.layers}-layer fp32 oracle port, dropout on, AdamW")
        metric = METRIC_TRAIN
    else:
        times, rows, threads = cpu_reference_steps(args.layers, args.steps, W, REF_B, args.denoise_steps, "denoise", budget_s=args.ref_budget_s, min_steps=min(3, args.steps))
        med = statistics.median(times)
        value = REF_B / med
        sample = f"median of {len(times)} denoise loops of {REF_B} captions x {args.denoise_steps} steps (lm_head every step, as CLIP-DDPM.py:616-617) after {W} warm-up loops"
        metric = METRIC_DENOISE
    line = {"impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "warmup": W,
            "ms_per_step": 1e3 * med, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args),
            "sample": sample, "ms_per_step_all": [round(1e3 * x, 1) for x in times],
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def eager_on_gpu(layers: int, B: int, S: int, autocast: bool, steps: int = 3, warmup: int = 2, workload: str = "train") -> dict:
    """The reference's own eager PyTorch op sequence (the oracle port) executed on the B200, fp32 or autocast(bf16), train mode, AdamW
    step included - SURVEY 2.1's "bar to beat". The reference materialises fp32 logits [rows, 16, 30522] twice (lm_head output +
    softmax), so its batch is capped by memory (B = 8 x S = 100 by default, CLIP-DDPM.py's own)."""
    torch.cuda.reset_peak_memory_stats()
    times, rows, _ = cpu_reference_steps(layers, steps, warmup, B, S, workload, "cuda", autocast)
    med = statistics.median(times)
    out = {"value": B / med, "unit": UNIT, "ms_per_step": 1e3 * med, "steps": len(times), "captions_per_step": B, "sample_size": S,
           "dtype": "autocast-bf16" if autocast else "f32", "max_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    return out


def run_reference_on_gpu(args):
    """Extra comparison, not the driver's reference arm (`--impl reference --ref-device cuda [--ref-autocast]`)."""
    r = eager_on_gpu(args.layers, args.ref_batch, args.ref_samples if args.workload == "train" else args.denoise_steps, args.ref_autocast,
                     args.steps, max(1, min(args.warmup, 2)), args.workload)
    emit({"impl": "reference", "ref_device": "cuda", "ref_dtype": r["dtype"], "metric": METRIC_TRAIN if args.workload == "train" else METRIC_DENOISE,
          "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": r["steps"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "data": "synthetic",
          "config": {"workload": f"eager PyTorch restatement of CLIP-DDPM.py on the B200, {args.workload}, {args.layers}L", "global_batch": args.ref_batch,
                     "sample_size": args.ref_samples, "layers": args.layers}, "max_memory_gb": r["max_memory_gb"]})


METRIC_TRAIN = "training samples/sec (seq=16)"
METRIC_DENOISE = "denoise-loop captions/sec (100 steps)"
UNIT = "captions/s"   # train: one caption = SAMPLE_SIZE + 1 = 101 noised sequences (encoder rows) per step; the same string in both arms and in e2e
L2_NOTE = "per-step working set (>10 GB of activations) >> 126 MB L2, plus a 256 MB L2 flush between timed steps"


def workload_config(args, model: str = None, layers: int = None):
    """`config` of the JSON line: a function of the command line only, so both arms (--impl ours / reference) print the same block."""
    B = args.batch
    model = model or args.model
    layers = layers or args.layers
    if model == "bert-large":
        return {"workload": f"CLIP-DDPM.py train_func on a bert-large-shaped encoder (24L/1024/16H/4096), seq_len=64 (+2 CLIP positions), bs={B} captions/GPU x "
                            f"SAMPLE_SIZE={args.samples} (+x_1 pass) = {B * (args.samples + 1)} encoder rows/step/GPU, x_0-predict, concat fusion, L1 + rounding CE, "
                            f"dropout 0.1, AdamW (BASELINE.json configs[4])", "global_batch": B * args.gpus, "seq_len": 64, "sample_size": args.samples,
                "layers": 24, "parallelism": f"dp{args.gpus}", "chunk_rows": args.chunk_rows, "l2": "working set >> L2, 256 MB L2 flush between timed steps"}
    if args.workload == "train" and getattr(args, "train_embedding", False):
        return {"workload": f"CLIP-DDPM.py train_func with TRAIN_EMBEDDING=True (IN_CHANNEL=16 learned embedding, trainable lm_head + in/out projections), "
                            f"{layers}L/768, seq_len=16, bs={B} x SAMPLE_SIZE={args.samples}", "global_batch": B * args.gpus, "seq_len": 16,
                "sample_size": args.samples, "layers": layers, "parallelism": f"dp{args.gpus}", "chunk_rows": args.chunk_rows,
                "l2": "working set >> L2, 256 MB L2 flush between timed steps"}
    if args.workload == "train":
        which = ("the model the reference actually trains (CLIP-DDPM.py:326,330: DistilBertConfig() = 6 layers; BASELINE.json calls it 'bert-base')" if layers == 6
                 else "SURVEY M1's reading of BASELINE.json's 'bert-base' (12-layer encoder)" if layers == 12 else f"{layers}-layer variant")
        return {"workload": f"CLIP-DDPM.py train_func, {layers}L/768/12H/3072 encoder - {which}, seq_len=16 (+2 CLIP positions), "
                            f"bs={B} captions/GPU x SAMPLE_SIZE={args.samples} (+x_1 pass) = {B * (args.samples + 1)} encoder rows/step/GPU, x_0-predict, concat fusion, "
                            f"L1 + rounding CE, dropout 0.1, AdamW", "global_batch": B * args.gpus, "seq_len": 16, "sample_size": args.samples,
                "layers": layers, "parallelism": f"dp{args.gpus}", "chunk_rows": args.chunk_rows, "l2": L2_NOTE}
    return {"workload": f"CLIP-DDPM.py denoise loop :611-621, {layers}L model, bs={args.denoise_batch}/GPU, n_steps={args.denoise_steps}, eval mode, "
                        f"fused lm_head+argmax on the last step", "global_batch": args.denoise_batch * args.gpus, "seq_len": 16, "layers": layers,
            "parallelism": f"shard{args.gpus}", "l2": "256 MB L2 flush between timed loops"}


# ------------------------------------------------------------------------------------------------------------ ours
class Ctx:
    """Per-process bench context: ranks, device, L2-flush buffer, barrier."""

    def __init__(self, args):
        from clipdlm import parallel
        self.rank, self.local_rank, self.world = parallel.init_process_group_from_env("nccl")
        assert self.world == args.gpus or self.world == 1, f"launched with WORLD_SIZE={self.world} but --gpus {args.gpus}"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def timed(self, model, fn, n: int, profile: bool = False):
        """n calls of fn, each bracketed by CUDA events on the launching stream, an L2 flush before each (outside the bracket), a
        barrier + synchronize on both sides; returns (sum of device ms = max over ranks, per-category profile, launches)."""
        evs = []
        self.barrier()
        if profile:
            model.profile(True)
            model.profile_read(reset=True)
        launches0 = model.launch_count()
        for _ in range(n):
            self.flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        self.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        prof = model.profile_read(reset=True) if profile else None
        if profile:
            model.profile(False)
        return self.max_over_ranks(ms), prof, model.launch_count() - launches0


def make_model(ctx: Ctx, args, kind: str, layers: int, precision: str, B: int, S: int, chunk_rows: int, dp: bool, train_embedding: bool = False):
    import clipdlm
    from clipdlm import parallel
    if kind == "bert-large":  # BASELINE.json configs[4]: bert-large-shaped 24L/1024/16H/4096, seq_len 64 (+2 CLIP positions)
        hp = clipdlm.default_hparams(BATCH_SIZE=B, SAMPLE_SIZE=S, N_LAYERS=24, DIM=1024, N_HEADS=16, HIDDEN_DIM=4096, MAX_LENGTH=64)
    else:
        hp = clipdlm.default_hparams(BATCH_SIZE=B, SAMPLE_SIZE=S, N_LAYERS=layers)
    if train_embedding:  # CLIP-DDPM.py:98-102: 16-channel learned embedding, trainable lm_head and in/out projections
        hp.update(TRAIN_EMBEDDING=True, IN_CHANNEL=16)
    torch.manual_seed(0)
    model = clipdlm.DistilBertModel(None, None, None, hp=hp, precision=precision, seed=0, chunk_rows=chunk_rows,
                                    fused_softmax_grad=None if args.fused_softmax_grad < 0 else bool(args.fused_softmax_grad),
                                    gelu_deriv_store=None if args.gelu_deriv_store < 0 else args.gelu_deriv_store)
    dp_mode = "single GPU"
    if ctx.world > 1 and dp:
        dp_mode = "nccl all-reduce of the flat fp32 gradients + AdamW on every rank"
        if args.dp_exchange in ("auto", "fused"):
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
    return model, trainer, hp, dp_mode


def free(*objs):
    import gc
    for o in objs:
        del o
    gc.collect()
    torch.cuda.empty_cache()


GEMM_CATS = ("gemm_fwd", "gemm_dgrad", "gemm_wgrad", "gemm_lse", "gemm_smgrad")


def kernel_table(prof, steps):
    return {k: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] / steps,
                "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None,
                "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 and v["bytes"] > 0 else None}
            for k, v in prof.items() if v["launches"]}


def roofline_block(prof, steps, ms_prof, pk):
    dom = max(GEMM_CATS, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    ach = d["flops"] / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else 0.0
    return {"bound": "tensor", "kernel": f"gemm_kernel ({dom}: tcgen05.mma.cta_group::2 256x256x16 CTA pairs, TMA-fed 5-stage ring, fused epilogue with TMA stores)",
            "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
            "peak_kind": f"{pk['source']} sustained bf16 (burst {pk['tf_burst']})", "avg_launch_ms": d["ms"] / max(d["launches"], 1),
            "flops_per_launch": d["flops"] / max(d["launches"], 1), "traffic": ncu_traffic(dom), "ms_per_step_with_event_brackets": ms_prof / steps,
            "all_gemm_tflops": sum(prof[k]["flops"] for k in GEMM_CATS) / max(sum(prof[k]["ms"] for k in GEMM_CATS), 1e-9) / 1e9}


def host_caption_set(B: int, n_batches: int, seed: int, ML: int):
    """Pinned-host caption set for the e2e leg: the product's own feed (data.py: DeviceCaptionDataset held in pinned host memory here, so
    that every step's batch really crosses PCIe) - n_batches x B pre-tokenised captions + CLIP features, random order per pass."""
    import clipdlm
    ds = clipdlm.synthetic_dataset(B * n_batches, max_length=ML, seed=seed, device="cpu")
    ds.attention_mask = torch.ones_like(ds.attention_mask)   # BASELINE configs[1]: full-length captions, like the resident batch
    ds.image = ds.text = None
    ds.pin_memory()
    return ds


def train_bench(ctx: Ctx, args, model, trainer, hp, steps: int, warm: int, legs=("resident", "profile", "e2e"), clocks: bool = False):
    """Times the train step of `model`: `resident` (batch already in HBM), `profile` (second pass with per-kernel event brackets),
    `e2e` (public API: loader over a pinned-host caption set -> H2D -> train_func -> loss read back)."""
    import clipdlm
    B = hp["BATCH_SIZE"]
    host = synthetic_host_batch(B, seed=ctx.rank, ML=hp["MAX_LENGTH"])
    dev_batch = {k: v.to(ctx.dev) for k, v in host.items()}
    model.train()
    out = {}

    def step_resident():
        return clipdlm.train_func(model, trainer, dev_batch)[0]

    for _ in range(warm):
        step_resident()
    if clocks:
        with ClockSampler(ctx.local_rank) as clk:
            ms, _, launches = ctx.timed(model, step_resident, steps, profile=args.profile_mode)
        out["clocks"] = clk.summary()
    else:
        ms, _, launches = ctx.timed(model, step_resident, steps)
    out.update(ms=ms, launches=launches + steps)   # + the optimizer kernel (launched by the trainer, not by an engine)
    if "profile" in legs and not args.profile_mode:
        out["ms_prof"], out["prof"], _ = ctx.timed(model, step_resident, steps, profile=True)
    if "e2e" in legs and not args.profile_mode:
        ds = host_caption_set(B, 4, seed=100 + ctx.rank, ML=hp["MAX_LENGTH"])
        loader = clipdlm.CaptionSubset(ds, torch.arange(len(ds))).loader(B, shuffle=True, generator=torch.Generator().manual_seed(ctx.rank), pin_staging=True)
        state = {"it": iter(loader)}

        def step_e2e():
            try:
                x = next(state["it"])
            except StopIteration:
                state["it"] = iter(loader)
                x = next(state["it"])
            b = {k: v.to(ctx.dev, non_blocking=True) for k, v in x.items()}
            return clipdlm.train_func(model, trainer, b)[0].item()  # loss read back: device -> host every step
        for _ in range(2):
            step_e2e()
        out["ms_e2e"], _, _ = ctx.timed(model, step_e2e, steps)
        out["h2d"] = sum(v.numel() * v.element_size() for v in host.values())
        out["d2h"] = 4
    return out


def denoise_bench(ctx: Ctx, args, model, B: int, n_steps: int, loops: int, warm: int, profile: bool = True):
    import clipdlm
    host = synthetic_host_batch(B, seed=ctx.rank)
    img_dev = host["image_clip"].to(ctx.dev)
    model.eval()

    def loop_resident():
        return clipdlm.sample(model, img_dev, n_steps=n_steps)[0]

    def loop_e2e():
        return clipdlm.sample(model, host["image_clip"].to(ctx.dev, non_blocking=True), n_steps=n_steps)[0].cpu()   # ids copied back
    for _ in range(warm):
        loop_resident()
    out = {}
    out["ms"], _, out["launches"] = ctx.timed(model, loop_resident, loops)
    if profile:
        out["ms_prof"], out["prof"], _ = ctx.timed(model, loop_resident, loops, profile=True)
    loop_e2e()
    out["ms_e2e"], _, _ = ctx.timed(model, loop_e2e, loops)
    out["h2d"], out["d2h"] = host["image_clip"].numel() * 4, B * 16 * 8
    return out


def dp_check(ctx: Ctx, model, trainer, hp):
    """N > 1, after the timed steps. (1) Every rank must hold bit-identical weights: spread of an integer checksum of the fp32 master weights
    across ranks (0 = identical). (2) One extra step with pinned draws, twice from the same state: through the model's own exchange (the fused
    reduce-scatter + AdamW + all-gather kernel) and through NCCL (all-reduce of the flat gradients, the single-GPU AdamW kernel on this rank's
    slice, broadcast of the slices) - the weights after the step must agree to summation order."""
    import ctypes as C
    import clipdlm
    from clipdlm import _lib as L
    dist = torch.distributed
    torch.cuda.synchronize()
    csum = model.flat.view(torch.int32).to(torch.int64).sum().reshape(1)
    allc = [torch.zeros_like(csum) for _ in range(ctx.world)]
    dist.all_gather(allc, csum)
    spread = int(max(int(c.item()) for c in allc) - min(int(c.item()) for c in allc))
    out = {"weights_checksum_spread_across_ranks": spread}
    if model.dp_fused is None:
        return out
    B, S, ML, D = hp["BATCH_SIZE"], hp["SAMPLE_SIZE"], hp["MAX_LENGTH"], hp["IN_CHANNEL"]
    g = torch.Generator().manual_seed(1000 + ctx.rank)
    batch = {k: v.to(ctx.dev) for k, v in synthetic_host_batch(B, seed=50 + ctx.rank, ML=ML).items()}
    t = torch.randint(0, 1000, (S, 1, 1), generator=torch.Generator().manual_seed(7))
    n_t, n_1 = torch.randn(B, ML, D, generator=g).to(ctx.dev), torch.randn(B, ML, D, generator=g).to(ctx.dev)
    w0, m0, v0, t0 = model.flat.clone(), trainer.m.clone(), trainer.v.clone(), trainer.t
    model.train()
    loss_f = clipdlm.train_func(model, trainer, batch, t=t, noise_t=n_t, noise_1=n_1, dropout_seed=4242)[0].item()
    w_fused = model.flat.clone()
    # same state again, NCCL exchange
    model.flat.copy_(w0); trainer.m.copy_(m0); trainer.v.copy_(v0); trainer.t = t0
    model.sync_shadow()
    ctx.barrier()
    step = trainer.step
    lo, hi = model.dp_fused["slice"]

    def nccl_step():
        dist.all_reduce(model.grad, group=model.dp_group)
        gp = trainer.param_groups[0]
        trainer.t += 1
        with torch.cuda.device(model.device):
            L.check(L.load().clipdlm_adamw(model.flat.data_ptr() + 4 * lo, model.grad.data_ptr() + 4 * lo, L.ptr(trainer.m), L.ptr(trainer.v),
                                           model.shadow_hi.data_ptr() + 2 * lo, None, hi - lo, gp["lr"], gp["betas"][0], gp["betas"][1], gp["eps"],
                                           gp["weight_decay"], trainer.t, 1.0 / ctx.world, 1, model._stream()))
        b, e = C.c_int64(), C.c_int64()
        for r in range(ctx.world):
            L.check(L.load().clipdlm_dp_slice(model.n_params, r, ctx.world, C.byref(b), C.byref(e)))
            dist.broadcast(model.flat[int(b.value):int(e.value)], src=r, group=model.dp_group)
        model.grad.zero_()
        model._grads_dirty = False
    trainer.step = nccl_step
    loss_n = clipdlm.train_func(model, trainer, batch, t=t, noise_t=n_t, noise_1=n_1, dropout_seed=4242)[0].item()
    trainer.step = step
    torch.cuda.synchronize()
    d = (model.flat.double() - w_fused.double())
    upd = (w_fused.double() - w0.double())
    out["fused_vs_nccl_one_step"] = {"loss_fused": loss_f, "loss_nccl": loss_n, "loss_delta": abs(loss_f - loss_n),
                                     "weights_max_abs_diff": float(d.abs().max()), "weights_rel_diff_of_update": float(d.norm() / upd.norm().clamp_min(1e-30)),
                                     "note": "same weights, batch, t, noise and dropout seed; the two runs differ by summation order only (fp32 atomics of the split-K "
                                             "weight gradients, cross-rank sum); Adam turns gradient noise on zero-gradient elements into +-lr steps, "
                                             f"lr = {trainer.param_groups[0]['lr']}"}
    model.flat.copy_(w_fused)
    model.sync_shadow()
    return out


def run_ours(args):
    import clipdlm
    if args.gemm_debug_flags:
        from clipdlm import _lib
        _lib.load().clipdlm_gemm_debug_flags(args.gemm_debug_flags)
    if args.attn_path:
        from clipdlm import _lib
        _lib.load().clipdlm_attn_force_simt(args.attn_path)
    ctx = Ctx(args)
    rank, world = ctx.rank, ctx.world
    pk = peaks()
    K = args.steps
    n_warm = args.warmup if args.profile_mode else max(args.warmup, 3)
    sub_K, sub_W = min(K, 3), 3
    if args.blocks == "auto":
        blocks = ([] if args.profile_mode or args.workload != "train" or args.model != "distilbert" or args.train_embedding or args.precision != "bf16" else
                  (["parity_mode", "denoise", "denoise_b8", "layers12", "eager_b200"] if world == 1 else ["denoise", "dp_check"] + (["bert_large"] if world == 8 else [])))
    else:
        blocks = [b for b in args.blocks.split(",") if b and b != "none"]
    line = {}
    if args.workload == "train":
        model, trainer, hp, dp_mode = make_model(ctx, args, args.model, args.layers, args.precision, args.batch, args.samples, args.chunk_rows, dp=True,
                                                 train_embedding=args.train_embedding)
        r = train_bench(ctx, args, model, trainer, hp, K, n_warm, clocks=True)
        units, S = args.batch, args.samples
        metric = METRIC_TRAIN
    else:
        model, trainer, hp, dp_mode = make_model(ctx, args, "distilbert", args.layers, args.precision, 8, 1, args.chunk_rows, dp=False)
        with ClockSampler(ctx.local_rank) as clk:
            r = denoise_bench(ctx, args, model, args.denoise_batch, args.denoise_steps, K, n_warm, profile=not args.profile_mode)
        r["clocks"] = clk.summary()
        units = args.denoise_batch
        metric = METRIC_DENOISE
        dp_mode = "batch shard, no collective"
    value = units * world * K / (r["ms"] / 1e3)
    ms_e2e = r.get("ms_e2e", r["ms"])
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": n_warm, "ms_per_step": r["ms"] / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args),
            "options": {"dp_exchange": dp_mode, "fused_softmax_grad": bool(model.fused_softmax_grad), "gelu_deriv_store": int(model.gelu_deriv_store),
                        "e2e_feed": "CaptionLoader over a pinned-host caption set (4 batches, shuffled), batch copied H2D every step, loss read back"},
            "e2e": {"value": units * world * K / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": r.get("h2d", 0), "d2h_bytes_per_step": r.get("d2h", 0),
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": r["launches"], "clocks": r.get("clocks")}
    if r.get("prof") is not None:
        line["roofline"] = roofline_block(r["prof"], K, r["ms_prof"], pk)
        line["kernels"] = kernel_table(r["prof"], K)
    if args.workload == "train":
        rows_s = value * (S + 1)
        alg = rows_s * TRAIN_GFLOP_PER_ROW.get("bert-large" if args.model == "bert-large" else args.layers, float("nan")) / 1e3
        line.update(noised_sequences_per_s=rows_s, algorithmic_tflops=alg, algorithmic_frac_of_sustained_peak=alg / pk["tf_sustained"] / world)
    if "dp_check" in blocks and world > 1 and args.workload == "train":
        try:
            line["dp_check"] = dp_check(ctx, model, trainer, hp)
        except Exception as ex:
            line["dp_check"] = {"error": f"{type(ex).__name__}: {ex}"}
    free(model, trainer)
    model = trainer = None

    def block(name, fn):
        if name not in blocks:
            return
        try:
            res = fn()
        except Exception as ex:   # a sub-benchmark must not take the headline down with it
            res = {"error": f"{type(ex).__name__}: {ex}"}
        if rank == 0:
            line[name] = res
        torch.cuda.empty_cache()

    def parity_mode():
        m, tr, hp2, _ = make_model(ctx, args, "distilbert", args.layers, "bf16x3", args.batch, args.samples, args.chunk_rows, dp=True)
        rr = train_bench(ctx, args, m, tr, hp2, sub_K, sub_W, legs=("resident",))
        free(m, tr)
        v = args.batch * world * sub_K / (rr["ms"] / 1e3)
        return {"what": "the same train step (same config) in precision='bf16x3': split-bf16 operands, three tensor-core passes per GEMM, fp32-class "
                        "results - the mode that meets the north star's 1e-3 / bit-exact-argmax gate (tests/test_parity_gpu.py::test_real_width_parity_mode_vs_reference)",
                "value": v, "unit": UNIT, "ms_per_step": rr["ms"] / sub_K, "steps": sub_K, "warmup": sub_W, "dtype": "bf16x3",
                "algorithmic_tflops": v * (args.samples + 1) * TRAIN_GFLOP_PER_ROW[args.layers] / 1e3, "slowdown_vs_bf16": (rr["ms"] / sub_K) / (r["ms"] / K)}

    def denoise():
        m, tr, hp2, _ = make_model(ctx, args, "distilbert", args.layers, "bf16", 8, 1, args.chunk_rows, dp=False)
        rr = denoise_bench(ctx, args, m, args.denoise_batch, args.denoise_steps, sub_K, sub_W)
        free(m, tr)
        B, n = args.denoise_batch, args.denoise_steps
        v = B * world * sub_K / (rr["ms"] / 1e3)
        flops = B * (n * FWD_GFLOP_PER_ROW_NO_HEAD[args.layers] + LM_HEAD_GFLOP_PER_ROW) * 1e9   # lm_head on the last step only
        return {"what": f"BASELINE.json configs[3]: CLIP-DDPM.py denoise loop :611-621, {args.layers}L, bs={B}/GPU x {n} steps, eval mode, fused lm_head+argmax on the last step; "
                        "images sharded across GPUs, no collective", "metric": METRIC_DENOISE, "value": v, "unit": UNIT, "ms_per_loop": rr["ms"] / sub_K,
                "loops": sub_K, "warmup": sub_W, "n_gpus": world, "global_batch": B * world,
                "e2e": {"value": B * world * sub_K / (rr["ms_e2e"] / 1e3), "unit": UNIT, "h2d_bytes_per_step": rr["h2d"], "d2h_bytes_per_step": rr["d2h"],
                        "ms_per_loop": rr["ms_e2e"] / sub_K},
                "algorithmic_tflops_per_gpu": flops / (rr["ms"] / sub_K / 1e3) / 1e12, "gpu_launches": rr["launches"],
                "roofline": roofline_block(rr["prof"], sub_K, rr["ms_prof"], pk), "kernels": kernel_table(rr["prof"], sub_K)}

    def denoise_b8():
        m, tr, hp2, _ = make_model(ctx, args, "distilbert", args.layers, "bf16", 8, 1, args.chunk_rows, dp=False)
        rr = denoise_bench(ctx, args, m, 8, 5, 20, 5, profile=False)
        free(m, tr)
        return {"what": "the reference's own evaluation shape (CLIP-DDPM.py:613-617): B = 8 images, 5 denoise steps, ids copied back; latency per batch",
                "ms_per_batch": rr["ms"] / 20, "ms_per_batch_e2e": rr["ms_e2e"] / 20, "value": 8 * 20 / (rr["ms_e2e"] / 1e3), "unit": UNIT,
                "launches_per_batch": rr["launches"] / 20, "cuda_graph": bool(getattr(clipdlm, "SAMPLE_USES_CUDA_GRAPH", False))}

    def layers12():
        a2 = argparse.Namespace(**vars(args)); a2.layers = 12
        m, tr, hp2, _ = make_model(ctx, a2, "distilbert", 12, "bf16", args.batch, args.samples, args.chunk_rows, dp=True)
        rr = train_bench(ctx, a2, m, tr, hp2, sub_K, sub_W, legs=("resident", "e2e"))
        free(m, tr)
        v = args.batch * world * sub_K / (rr["ms"] / 1e3)
        alg = v * (args.samples + 1) * TRAIN_GFLOP_PER_ROW[12] / 1e3
        return {"what": "SURVEY M1's reading of 'bert-base': 12-layer encoder (DistilBertConfig(n_layers=12)), same train step and batch", "config": workload_config(a2, layers=12),
                "value": v, "unit": UNIT, "ms_per_step": rr["ms"] / sub_K, "steps": sub_K, "warmup": sub_W, "e2e": {"value": args.batch * world * sub_K / (rr["ms_e2e"] / 1e3), "unit": UNIT},
                "algorithmic_tflops": alg, "algorithmic_frac_of_sustained_peak": alg / pk["tf_sustained"] / world}

    def bert_large():
        a2 = argparse.Namespace(**vars(args)); a2.model, a2.batch, a2.chunk_rows = "bert-large", 64, 1024
        m, tr, hp2, mode = make_model(ctx, a2, "bert-large", 24, "bf16", 64, args.samples, 1024, dp=True)
        rr = train_bench(ctx, a2, m, tr, hp2, sub_K, sub_W, legs=("resident", "e2e"))
        chk = dp_check(ctx, m, tr, hp2) if world > 1 else None
        free(m, tr)
        v = 64 * world * sub_K / (rr["ms"] / 1e3)
        alg = v * (args.samples + 1) * TRAIN_GFLOP_PER_ROW["bert-large"] / 1e3
        return {"what": "BASELINE.json configs[4]", "config": workload_config(a2, model="bert-large"), "value": v, "unit": UNIT, "ms_per_step": rr["ms"] / sub_K,
                "steps": sub_K, "warmup": sub_W, "n_gpus": world, "dp_exchange": mode, "e2e": {"value": 64 * world * sub_K / (rr["ms_e2e"] / 1e3), "unit": UNIT},
                "algorithmic_tflops": alg, "algorithmic_frac_of_sustained_peak": alg / pk["tf_sustained"] / world, "dp_check": chk}

    def eager_b200():
        out = {"what": "the reference's eager PyTorch op sequence (oracle port of CLIP-DDPM.py train_func, AdamW included) on this same B200 - SURVEY 2.1's bar; "
                       f"its fp32 logits cap the batch: B = {REF_B} x S = {REF_S} (the reference's own step)"}
        for name, ac in (("f32", False), ("autocast_bf16", True)):
            out[name] = eager_on_gpu(args.layers, REF_B, REF_S, ac)
            torch.cuda.empty_cache()
        m, tr, hp2, _ = make_model(ctx, args, "distilbert", args.layers, "bf16", REF_B, REF_S, args.chunk_rows, dp=False)
        rr = train_bench(ctx, args, m, tr, hp2, 10, 3, legs=("resident",))
        free(m, tr)
        out["ours_same_batch"] = {"value": REF_B * 10 / (rr["ms"] / 1e3), "unit": UNIT, "ms_per_step": rr["ms"] / 10, "dtype": "bf16"}
        return out

    block("parity_mode", parity_mode)
    block("denoise", denoise)
    block("denoise_b8", denoise_b8)
    block("layers12", layers12)
    block("bert_large", bert_large)
    block("eager_b200", eager_b200)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and not args.profile_mode:
            wl = "train" if args.workload == "train" else "denoise"
            S_ref = REF_S if wl == "train" else args.denoise_steps
            times, rows, threads = cpu_reference_steps(args.layers, 2, 1, REF_B, S_ref, wl, budget_s=40.0, min_steps=1)
            med = statistics.median(times)
            line["cpu_baseline"] = {"value": REF_B / med, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": (f"median of {len(times)} steps after 1 warm-up; one step = {REF_B} captions x {S_ref}" +
                                               ("+1 noise levels (the reference's own step, BASELINE.json configs[0]), fp32 oracle port, dropout on, AdamW" if wl == "train"
                                                else " denoise steps, lm_head every step"))}
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
    ap.add_argument("--blocks", default="auto", help="sub-benchmarks appended to the line: auto | none | comma list of parity_mode,denoise,denoise_b8,layers12,"
                                                     "bert_large,eager_b200,dp_check")
    ap.add_argument("--train-embedding", action="store_true", help="TRAIN_EMBEDDING=True variant of the train step (use with --no-cpu-baseline)")
    ap.add_argument("--dp-exchange", default="auto", choices=["auto", "fused", "nccl"],
                    help="N > 1 gradient exchange: fused = reduce-scatter + AdamW + all-gather in one kernel over NVLink peer memory; auto = fused, NCCL if unavailable")
    ap.add_argument("--fused-softmax-grad", type=int, default=-1, choices=[-1, 0, 1],
                    help="factored softmax-CE gradient of the lm_head (default on in bf16; 0 = the in-place pass over stored logits)")
    ap.add_argument("--gelu-deriv-store", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="lin1 stores gelu'(u), the lin2 gradient GEMM multiplies by it (2, the bf16 default: that GEMM also sums the lin1 bias gradient); 0 = off")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"],
                    help="bf16 = speed mode (BASELINE.json configs[1]); bf16x3 = parity mode (split-bf16 operands, fp32-class results)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"], help="--impl reference: 'cuda' runs the eager PyTorch port on the GPU (extra comparison)")
    ap.add_argument("--ref-autocast", action="store_true", help="--ref-device cuda under torch.autocast(bfloat16)")
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--ref-samples", type=int, default=100)
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: wall-time bound; past it the loop stops once 5 timed steps exist")
    ap.add_argument("--gemm-debug-flags", type=int, default=0, help="triage: clipdlm_gemm_debug_flags bits (4096 = no band tile order in the lm_head passes, "
                                                                    "2048 = LSE_EXP output through row stores instead of TMA stores)")
    ap.add_argument("--attn-path", type=int, default=0, help="triage: clipdlm_attn_force_simt (0 default, 4 = packed attention without the pipelined backward, "
                                                             "3 = 32-row-slot tcgen05, 2 = mma.sync ring)")
    ap.add_argument("--profile-mode", action="store_true", help="for runs under ncu: no forced warm-up, no e2e leg, no sub-benchmarks, no CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
