"""Generates tests/golden/*.npz by running the REAL reference (CLIP-DDPM.py slices exec'd from /root/reference, driving HF
DistilBertForMaskedLM) on closed-form (generator-free) weights and inputs. Run in the build container:

    python tests/golden/make_golden.py

The fixtures travel with the repo; /root/reference does not. tests/test_oracle_golden.py replays them against the oracle
(CPU) and tests/test_parity_gpu.py against the CUDA path (B200).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import clipdlm_oracle as O  # noqa: E402
from oracle import reference_harness as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def golden_hp(**kw):
    hp = O.default_hparams()
    hp.update(BATCH_SIZE=3, SAMPLE_SIZE=4, N_LAYERS=2, VOCAB_SIZE=997, DROPOUT=0.0, ATTENTION_DROPOUT=0.0)
    hp.update(kw)
    return hp


GOLDEN_T = [0, 17, 400, 999]


class _TorchProxy:
    """`torch` as seen by the exec'd reference code, with the two random draws of train_func pinned."""

    def __init__(self, t, noises):
        self._t, self._noises = t, list(noises)

    def __getattr__(self, k):
        return getattr(torch, k)

    def randint(self, *a, **k):
        return self._t.clone()

    def normal(self, *a, **k):
        return self._noises.pop(0).clone()


def make_case(name: str, full: bool = True, **kw):
    hp = golden_hp(**kw)
    ns = H.build_namespace(hp)
    model = H.build_model(ns, hp, seed=0)
    P = O.init_params(hp, seed=0, closed_form=True)
    H.load_params(model, P)
    batch = O.closed_form_batch(hp, k=1, ragged=True)
    B, S, ML, D = hp["BATCH_SIZE"], hp["SAMPLE_SIZE"], hp["MAX_LENGTH"], hp["IN_CHANNEL"]  # == DIM unless TRAIN_EMBEDDING
    out = {}
    # forward (eval) on explicit rows
    model.eval()
    R = 2
    xin = O.closed_form_tensor((R, ML, D), 3, 0.5)
    img = O.closed_form_tensor((R, 1, hp["CLIP_DIM"]), 4, 0.05)
    txt = O.closed_form_tensor((R, 1, hp["CLIP_DIM"]), 5, 0.05)
    mask = torch.ones(R, ML, dtype=torch.int64); mask[1, 9:] = 0
    with torch.no_grad():
        logits, x_out = model(xin, img, txt, mask, torch.tensor([1, 0]).repeat(R, 1))
    out["fwd_x_out"] = x_out.numpy() if full else x_out[:, :, :8].numpy()
    out["fwd_argmax"] = logits.argmax(-1).numpy()
    out["fwd_logits_head"] = logits[:, :, :64].numpy()
    top2 = logits.topk(2, dim=-1).values
    out["fwd_top2_gap"] = (top2[..., 0] - top2[..., 1]).numpy()
    # denoise loop, 5 steps
    Lfull = x_out.shape[1]
    restored = O.closed_form_tensor((B, Lfull, D), 6, 1.0)
    r = restored.clone()
    with torch.no_grad():
        for _ in range(5):
            o, r = model(r[:, :ML, :], batch["image_clip"].unsqueeze(1), torch.zeros_like(batch["image_clip"]).unsqueeze(1),
                         torch.ones(B, ML), torch.tensor([1, 0]).repeat(B, 1))
    out["sample_ids"] = torch.softmax(o, -1).argmax(-1).numpy()
    t2 = o.topk(2, dim=-1).values
    out["sample_top2_gap"] = (t2[..., 0] - t2[..., 1]).numpy()
    out["sample_restored"] = r.numpy() if full else r[:, :, :8].numpy()
    # one train step (train mode; dropout p = 0), pinned t / noise
    model.train()
    t = torch.tensor(GOLDEN_T).reshape(S, 1, 1)
    n_t = O.closed_form_tensor((B, ML, D), 7, 1.0)
    n_1 = O.closed_form_tensor((B, ML, D), 8, 1.0)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    ns["torch"] = _TorchProxy(t, [n_t, n_1])
    l, a, b, c = ns["train_func"](model, opt, batch)
    ns["torch"] = torch
    out["train_losses"] = np.array([l.item(), a.item(), b.item(), c.item()], dtype=np.float64)
    grads = {n: p.grad.detach() for n, p in model.named_parameters() if p.grad is not None}
    names = [n for n in O.trainable_names(hp) if n in grads]
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([float(grads[n].double().norm()) for n in names])
    for n in names:  # full gradients of the small tensors + first rows of the matrices
        g = grads[n]
        out["grad::" + n] = (g if g.numel() <= 4096 else g.reshape(-1)[:(4096 if full else 256)]).numpy()
    after = H.export_params(model)
    out["after_norms"] = np.array([float(after[n].double().norm()) for n in names])
    for n in ("model.vocab_layer_norm.weight", "model.distilbert.transformer.layer.0.ffn.lin1.bias", "image_linear.bias"):
        out["after::" + n] = after[n].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "losses", out["train_losses"], "min top-2 gaps", out["fwd_top2_gap"].min(), out["sample_top2_gap"].min())


REAL_WIDTH_HP = dict(BATCH_SIZE=8, SAMPLE_SIZE=100, N_LAYERS=6, VOCAB_SIZE=30522, DROPOUT=0.0, ATTENTION_DROPOUT=0.0)


def real_width_t(S: int) -> torch.Tensor:
    """Pinned noise levels for the real-width case: spread over 0..999, includes t = 0 and t = 999."""
    t = (torch.arange(S, dtype=torch.int64) * 373 + 11) % 1000
    t[0], t[-1] = 0, 999
    return t.reshape(S, 1, 1)


def make_real_width(name: str = "real_width_6L", steps: int = 2):
    """BASELINE.json configs[0] dimensions: the reference's own defaults (CLIP-DDPM.py:55-114) - 6 layers, V = 30522, B = 8, S = 100
    (808 encoder rows per step) - with dropout 0 and pinned t / noise. Stores scalars, norms and slices only (the gradients are
    44 M elements): per-step losses of `steps` consecutive train_func calls (AdamW lr 1e-4 in between), every gradient's norm + its first
    256 elements after step 0, post-AdamW parameter norms, a strided sample of the step-0 x_out / logits via a 5-step denoise loop."""
    hp = O.default_hparams()
    hp.update(REAL_WIDTH_HP)
    ns = H.build_namespace(hp)
    model = H.build_model(ns, hp, seed=0)
    P = O.init_params(hp, seed=0, closed_form=True)
    H.load_params(model, P)
    batch = O.closed_form_batch(hp, k=1, ragged=True)
    B, S, ML, D = hp["BATCH_SIZE"], hp["SAMPLE_SIZE"], hp["MAX_LENGTH"], hp["IN_CHANNEL"]
    out = {}
    # 5-step denoise loop BEFORE training (the reference's eval shape, :613-621)
    model.eval()
    restored = O.closed_form_tensor((B, ML + 2, D), 6, 1.0)
    r = restored.clone()
    ids_steps, gaps = [], []
    with torch.no_grad():
        for _ in range(5):
            o, r = model(r[:, :ML, :], batch["image_clip"].unsqueeze(1), torch.zeros_like(batch["image_clip"]).unsqueeze(1),
                         torch.ones(B, ML), torch.tensor([1, 0]).repeat(B, 1))
            ids_steps.append(torch.softmax(o, -1).argmax(-1).numpy())
            t2 = o.topk(2, dim=-1).values
            gaps.append((t2[..., 0] - t2[..., 1]).numpy())
    # the reference ITSELF under torch.autocast(bfloat16) on the same inputs: how far plain-bf16 arithmetic moves its own arg-max ids
    # (the yardstick for the speed mode's agreement gate; this fixture's untrained, closed-form weights give nearly flat logits)
    r16 = restored.clone()
    ids16 = []
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        for _ in range(5):
            o16, r16 = model(r16[:, :ML, :].float(), batch["image_clip"].unsqueeze(1), torch.zeros_like(batch["image_clip"]).unsqueeze(1),
                             torch.ones(B, ML), torch.tensor([1, 0]).repeat(B, 1))
            r16 = r16.float()
            ids16.append(torch.softmax(o16.float(), -1).argmax(-1).numpy())
    out["autocast_bf16_ids_steps"] = np.stack(ids16)
    out["sample_ids_steps"] = np.stack(ids_steps)
    out["sample_top2_gap_steps"] = np.stack(gaps)
    out["sample_restored_slice"] = r[:, :, ::16].numpy()
    out["sample_logits_slice"] = o[:, :, ::509].numpy()
    # train steps
    model.train()
    t = real_width_t(S)
    out["t"] = t.reshape(-1).numpy()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    losses = []
    for k in range(steps):
        n_t = O.closed_form_tensor((B, ML, D), 7 + 10 * k, 1.0)
        n_1 = O.closed_form_tensor((B, ML, D), 8 + 10 * k, 1.0)
        ns["torch"] = _TorchProxy(t, [n_t, n_1])
        l, a, b, c = ns["train_func"](model, opt, batch)
        ns["torch"] = torch
        losses.append([l.item(), a.item(), b.item(), c.item()])
        print(name, "step", k, losses[-1], flush=True)
        if k == 0:
            grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
            names = [n for n in O.trainable_names(hp) if n in grads]
            out["grad_names"] = np.array(names)
            out["grad_norms"] = np.array([float(grads[n].double().norm()) for n in names])
            for n in names:
                out["grad::" + n] = grads[n].reshape(-1)[:256].numpy()
            after = H.export_params(model)
            out["after_norms"] = np.array([float(after[n].double().norm()) for n in names])
            out["after_delta_norms"] = np.array([float((after[n].double() - P[n].double()).norm()) for n in names])
    out["train_losses"] = np.array(losses, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "min top-2 gap", out["sample_top2_gap_steps"].min())


if __name__ == "__main__":
    if not H.available():
        sys.exit("reference unavailable")
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "real":  # BASELINE.json configs[0] dimensions (about 2 minutes of CPU)
        make_real_width()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "te":  # only the TRAIN_EMBEDDING cases (added later; the others are unchanged)
        make_case("te_concat_l1", TRAIN_EMBEDDING=True, IN_CHANNEL=16, CLIP_ADDING_METHOD="concat", LOSS_FUNC="series_sum_sample_mean")
        make_case("te_add_mse_mean", full=False, TRAIN_EMBEDDING=True, IN_CHANNEL=16, CLIP_ADDING_METHOD="add", LOSS_FUNC="mse_series_mean")
        sys.exit(0)
    make_case("concat_l1", CLIP_ADDING_METHOD="concat", LOSS_FUNC="series_sum_sample_mean")
    make_case("add_l1", CLIP_ADDING_METHOD="add", LOSS_FUNC="series_sum_sample_mean")
    make_case("concat_mse_mean", full=False, CLIP_ADDING_METHOD="concat", LOSS_FUNC="mse_series_mean")
    make_case("concat_series_sum", full=False, CLIP_ADDING_METHOD="concat", LOSS_FUNC="series_sum")
    make_case("concat_mse_sum", full=False, CLIP_ADDING_METHOD="concat", LOSS_FUNC="mse_series_sum")
    make_case("te_concat_l1", TRAIN_EMBEDDING=True, IN_CHANNEL=16, CLIP_ADDING_METHOD="concat", LOSS_FUNC="series_sum_sample_mean")
    make_case("te_add_mse_mean", full=False, TRAIN_EMBEDDING=True, IN_CHANNEL=16, CLIP_ADDING_METHOD="add", LOSS_FUNC="mse_series_mean")
