#!/usr/bin/env python
"""A/B of the micro-chunk size of the train step (bench.py --chunk-rows) on one B200, with a correctness gate.

The step of bs = 512 x S = 100 runs as chunks of `chunk_rows` encoder rows; every GEMM of a chunk ends in a partial wave of the
74 CTA pairs, so fewer / larger chunks lose less to wave quantisation (model: 1.8 % of GEMM time at 8192 rows, 0.3 % at 25600).
Larger chunks also push byte offsets past 2^31, so phase 1 checks that the gradients do not depend on the chunking (dropout off,
identical weights and draws; only the fp32 summation order of the split-K weight gradients may differ) before phase 2 times it.

  python tools/chunk_ab.py [--chunks 8192,25600,51200] [--steps 4]        -> one JSON line per measurement on stdout
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def batch_for(B: int, dev):
    g = torch.Generator().manual_seed(0)
    return {"input_ids": torch.randint(0, 30522, (B, 16), generator=g).to(dev), "attention_mask": torch.ones(B, 16, dtype=torch.int64, device=dev),
            "image_clip": torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).to(dev),
            "text_clip": torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).to(dev)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", default="8192,25600,51200")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--samples", type=int, default=100)
    args = ap.parse_args()
    chunks = [int(c) for c in args.chunks.split(",")]
    import clipdlm
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B, S = args.batch, args.samples
    batch = batch_for(B, dev)
    g = torch.Generator().manual_seed(1)
    t = torch.randint(0, 1000, (S, 1, 1), generator=g)
    n_t, n_1 = torch.randn(B, 16, 768, generator=g), torch.randn(B, 16, 768, generator=g)

    # ---- phase 1: gradients must not depend on the chunking
    hp = clipdlm.default_hparams(BATCH_SIZE=B, SAMPLE_SIZE=S, DROPOUT=0.0, ATTENTION_DROPOUT=0.0)
    model = clipdlm.DistilBertModel(None, None, None, hp=hp, precision="bf16", seed=0, chunk_rows=chunks[0]).train()
    trainer = clipdlm.AdamW(model.parameters(), lr=1e-4)
    snap = {}

    def keep_grads_skip_update():
        snap["g"] = model.grad.clone()
        model.grad.zero_()
        model._grads_dirty = False

    trainer.step = keep_grads_skip_update
    ref_g = ref_l = None
    ok = {}
    for c in chunks:
        model.chunk_rows = c
        try:
            out = clipdlm.train_func(model, trainer, batch, t=t, noise_t=n_t, noise_1=n_1, dropout_seed=1)
            torch.cuda.synchronize()
        except Exception as ex:  # an out-of-range chunk size must show up as a line, not end the run
            print(json.dumps({"phase": "gradients", "chunk_rows": c, "error": f"{type(ex).__name__}: {ex}"}), flush=True)
            ok[c] = False
            break  # a CUDA fault poisons the context: nothing after it can be trusted
        losses = [float(x.item()) for x in out]
        gr = snap["g"].double()
        if ref_g is None:
            ref_g, ref_l = gr, losses
        rel = float((gr - ref_g).norm() / ref_g.norm())
        lrel = max(abs(a - b) / abs(b) for a, b in zip(losses, ref_l))
        ok[c] = bool(rel < 2e-4 and lrel < 1e-5 and torch.isfinite(gr).all())
        print(json.dumps({"phase": "gradients", "chunk_rows": c, "grad_rel_vs_first": rel, "loss_rel_vs_first": lrel, "grad_norm": float(gr.norm()),
                          "losses": losses, "ok": ok[c], "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    del model, trainer, snap, ref_g
    torch.cuda.empty_cache()

    # ---- phase 2: time the real step (dropout 0.1, AdamW) per chunk size; first size repeated last to expose drift of the box
    hp = clipdlm.default_hparams(BATCH_SIZE=B, SAMPLE_SIZE=S)
    model = clipdlm.DistilBertModel(None, None, None, hp=hp, precision="bf16", seed=0, chunk_rows=chunks[0]).train()
    trainer = clipdlm.AdamW(model.parameters(), lr=1e-4)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for c in [c for c in chunks if ok.get(c)] + ([chunks[0]] if ok.get(chunks[0]) else []):
        model.chunk_rows = c
        for _ in range(2):
            clipdlm.train_func(model, trainer, batch)
        evs = []
        torch.cuda.synchronize()
        for _ in range(args.steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            clipdlm.train_func(model, trainer, batch)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        print(json.dumps({"phase": "timing", "chunk_rows": c, "ms_per_step_mean": sum(ms) / len(ms), "ms_per_step_min": ms[0],
                          "captions_per_s": B / (sum(ms) / len(ms) / 1e3), "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


if __name__ == "__main__":
    main()
