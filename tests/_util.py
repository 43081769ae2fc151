"""Shared helpers for the test-suite (tests/ may import oracle/; the product package never does)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import clipdlm_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_T = [0, 17, 400, 999]


def golden_hp(**kw):
    hp = O.default_hparams()
    hp.update(BATCH_SIZE=3, SAMPLE_SIZE=4, N_LAYERS=2, VOCAB_SIZE=997, DROPOUT=0.0, ATTENTION_DROPOUT=0.0)
    hp.update(kw)
    return hp


GOLDEN_CASES = {
    "concat_l1": dict(CLIP_ADDING_METHOD="concat", LOSS_FUNC="series_sum_sample_mean"),
    "add_l1": dict(CLIP_ADDING_METHOD="add", LOSS_FUNC="series_sum_sample_mean"),
    "concat_mse_mean": dict(CLIP_ADDING_METHOD="concat", LOSS_FUNC="mse_series_mean"),
    "concat_series_sum": dict(CLIP_ADDING_METHOD="concat", LOSS_FUNC="series_sum"),
    "concat_mse_sum": dict(CLIP_ADDING_METHOD="concat", LOSS_FUNC="mse_series_sum"),
    # TRAIN_EMBEDDING=True (CLIP-DDPM.py:238-243): 16-channel learned embedding, trainable lm_head and in/out projections
    "te_concat_l1": dict(TRAIN_EMBEDDING=True, IN_CHANNEL=16, CLIP_ADDING_METHOD="concat", LOSS_FUNC="series_sum_sample_mean"),
    "te_add_mse_mean": dict(TRAIN_EMBEDDING=True, IN_CHANNEL=16, CLIP_ADDING_METHOD="add", LOSS_FUNC="mse_series_mean"),
}
FULL_CASES = ("concat_l1", "add_l1", "te_concat_l1")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_inputs(hp):
    """The closed-form inputs tests/golden/make_golden.py fed to the reference."""
    B, S, ML, D = hp["BATCH_SIZE"], hp["SAMPLE_SIZE"], hp["MAX_LENGTH"], hp["IN_CHANNEL"]  # == DIM unless TRAIN_EMBEDDING
    R = 2
    mask = torch.ones(R, ML, dtype=torch.int64)
    mask[1, 9:] = 0
    Lfull = ML + (2 if hp["CLIP_ADDING_METHOD"] == "concat" else 0)
    return dict(
        batch=O.closed_form_batch(hp, k=1, ragged=True),
        fwd_x=O.closed_form_tensor((R, ML, D), 3, 0.5), fwd_img=O.closed_form_tensor((R, 1, hp["CLIP_DIM"]), 4, 0.05),
        fwd_txt=O.closed_form_tensor((R, 1, hp["CLIP_DIM"]), 5, 0.05), fwd_mask=mask,
        restored=O.closed_form_tensor((B, Lfull, D), 6, 1.0),
        t=torch.tensor(GOLDEN_T).reshape(S, 1, 1), noise_t=O.closed_form_tensor((B, ML, D), 7, 1.0),
        noise_1=O.closed_form_tensor((B, ML, D), 8, 1.0),
    )


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ---- BASELINE.json configs[0] dimensions (the reference's own defaults): tests/golden/real_width_6L.npz, make_golden.py::make_real_width
REAL_WIDTH_HP = dict(BATCH_SIZE=8, SAMPLE_SIZE=100, N_LAYERS=6, VOCAB_SIZE=30522, DROPOUT=0.0, ATTENTION_DROPOUT=0.0)


def real_width_inputs(step: int = 0):
    hp = O.default_hparams()
    hp.update(REAL_WIDTH_HP)
    B, S, ML, D = hp["BATCH_SIZE"], hp["SAMPLE_SIZE"], hp["MAX_LENGTH"], hp["IN_CHANNEL"]
    t = (torch.arange(S, dtype=torch.int64) * 373 + 11) % 1000
    t[0], t[-1] = 0, 999
    return hp, dict(batch=O.closed_form_batch(hp, k=1, ragged=True), t=t.reshape(S, 1, 1), restored=O.closed_form_tensor((B, ML + 2, D), 6, 1.0),
                    noise_t=O.closed_form_tensor((B, ML, D), 7 + 10 * step, 1.0), noise_1=O.closed_form_tensor((B, ML, D), 8 + 10 * step, 1.0))
