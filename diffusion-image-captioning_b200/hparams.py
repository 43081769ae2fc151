"""The reference's hyperparameter "dict": the ~35 module-level globals of CLIP-DDPM.py:55-114, same names and defaults
(SURVEY App. A), plus the implicit AdamW / DistilBertConfig defaults the reference inherits from torch / transformers."""
from __future__ import annotations

import math

import torch

LOSS_KIND = {"series_sum_sample_mean": 0, "series_sum": 1, "mse_series_mean": 2, "mse_series_sum": 3}  # CLIP-DDPM.py:77-87


def default_hparams(**overrides) -> dict:
    hp = dict(
        DEBUG=False, CONTINUE_TRAIN=False, BATCH_SIZE=8, MAX_LENGTH=16, LEARNING_RATE=1e-4, END_LEARNING_RATE=5e-5,  # :55-60
        SCHEDULER="linspace",  # torch.linspace :69 (alternatives "logspace" :68, "cosine" :63-67)
        TRAIN_SET_RATIO=0.8, EARLY_STOP_RATIO=1.05, EPOCH_NUM=5, DYNAMIC_ROUNDING_WEIGHT=-1, ROUNDING_WEIGHT=0.5,  # :71-75
        LOSS_FUNC="series_sum_sample_mean",  # :89
        CLIP_ADDING_METHOD="concat", CLASSIFIER_FREE_WEIGHT=0, CLASSIFIER_FREE_PROB=0.2,  # :94-97
        TRAIN_EMBEDDING=False, IN_CHANNEL=768,  # :98-102
        BETA_MIN=0.0001, BETA_MAX=0.02, STEP_TOT=1000, COSIN_SCHEDULE=True, SAMPLE_SIZE=100,  # :105-109
        X_0_PREDICTION=True, X_T_STEP_INTERVAL=100, USE_X_T_LOSS=True, USE_X_1_LOSS=True, USE_PROB_LOSS=True,  # :110-114
        VOCAB_SIZE=30522,  # tokenizer.vocab_size :206
        # torch.optim.AdamW(model.parameters(), lr) defaults (:335)
        ADAM_BETAS=(0.9, 0.999), ADAM_EPS=1e-8, WEIGHT_DECAY=0.01,
        # DistilBertConfig() defaults (HF configuration_distilbert.py:58-74), the model CLIP-DDPM.py:236,326 builds
        N_LAYERS=6, DIM=768, N_HEADS=12, HIDDEN_DIM=3072, DROPOUT=0.1, ATTENTION_DROPOUT=0.1, MAX_POSITION=512, CLIP_DIM=512,
    )
    unknown = set(overrides) - set(hp)
    if unknown:
        raise KeyError(f"unknown hyperparameters: {sorted(unknown)}")
    hp.update(overrides)
    if "IN_CHANNEL" not in overrides:
        hp["IN_CHANNEL"] = 16 if hp["TRAIN_EMBEDDING"] else hp["DIM"]  # :99-102
    return hp


# The reference keeps its hyperparameters (and `val_loader`) as MODULE GLOBALS that `diffuse_t(x, t)`, `generate_diffuse_pair(x_0, t, t_next)`,
# `validate(model)` read implicitly (CLIP-DDPM.py:55-114,221,347,364,488). GLOBALS plays that role for the verbatim call forms: functions that
# are not handed a model or an explicit `hp` read it. `bind()` (diffusion.py) sets it.
GLOBALS: dict = default_hparams()
ACTIVE = {"val_loader": None}


def set_globals(hp: dict = None, val_loader=None, **overrides) -> dict:
    """Make `hp` (+ overrides) the active hyperparameters / `val_loader` the active validation loader, like editing the constants at the
    top of CLIP-DDPM.py. Returns the active dict (the same object the functions read)."""
    if hp is not None:
        GLOBALS.clear()
        GLOBALS.update(hp)
    if overrides:
        unknown = set(overrides) - set(GLOBALS)
        if unknown:
            raise KeyError(f"unknown hyperparameters: {sorted(unknown)}")
        GLOBALS.update(overrides)
    if val_loader is not None:
        ACTIVE["val_loader"] = val_loader
    return GLOBALS


def model_name(hp: dict) -> str:
    """MODEL_NAME f-string of CLIP-DDPM.py:116-118 (file stem of the reference's logs / pickles)."""
    sched = {"linspace": "linspace", "logspace": "logspace", "cosine": "cosine_annealing"}[hp["SCHEDULER"]]
    e = lambda v: "%.0E" % v
    return (f"epoch{hp['EPOCH_NUM']}_loss{hp['LOSS_FUNC']}_lr{e(hp['LEARNING_RATE'])}-{e(hp['END_LEARNING_RATE'])}_scheduler{sched}"
            f"_round{e(hp['ROUNDING_WEIGHT'])}_dynamic{hp['DYNAMIC_ROUNDING_WEIGHT']}"
            f"_clip{hp['CLIP_ADDING_METHOD']}_class_weight{e(hp['CLASSIFIER_FREE_WEIGHT'])}_class_prob{e(hp['CLASSIFIER_FREE_PROB'])}"
            f"_train-embed{hp['TRAIN_EMBEDDING']}"
            f"_samplesize{hp['SAMPLE_SIZE']}_x_0_predict{hp['X_0_PREDICTION']}_X_INTERVAL{hp['X_T_STEP_INTERVAL']}"
            f"_use_x_t{hp['USE_X_T_LOSS']}_use_x_1{hp['USE_X_1_LOSS']}_use_prob{hp['USE_PROB_LOSS']}")


def learning_rates(hp: dict) -> list:
    """Per-epoch learning rates `lrs`, CLIP-DDPM.py:451-456 (cosine_annealing :63-67: 5-epoch cosine, repeated 3 times)."""
    n, lr0, lr1 = hp["EPOCH_NUM"], hp["LEARNING_RATE"], hp["END_LEARNING_RATE"]
    kind = hp["SCHEDULER"]
    if kind == "linspace":
        return torch.linspace(lr0, lr1, n).tolist()
    if kind == "logspace":  # the reference passes log10 endpoints (:453-454)
        return torch.logspace(torch.tensor([lr0]).log10().item(), torch.tensor([lr1]).log10().item(), n).tolist()
    if kind == "cosine":
        sub_epoch = 5
        x = torch.arange(0, sub_epoch)
        x = lr1 + (lr0 - lr1) * (1 + torch.cos(x / sub_epoch * math.pi)) / 2
        return x.repeat((3,)).tolist()
    raise NotImplementedError(kind)


def alpha_cumprod(hp: dict, device="cpu") -> torch.Tensor:
    """alpha-bar schedule, CLIP-DDPM.py:337-346: cosine (default) or linear-beta. fp32 [STEP_TOT]."""
    if hp["COSIN_SCHEDULE"]:
        def scheduler(t):
            s = 0.008
            return torch.cos(math.pi / 2 * (t / hp["STEP_TOT"] + s) / (1 + s)) ** 2
        ts = torch.arange(hp["STEP_TOT"]).to(device)
        return scheduler(ts) / scheduler(torch.zeros(1, device=device))
    betas = torch.hstack([torch.zeros(1), torch.linspace(hp["BETA_MIN"], hp["BETA_MAX"], hp["STEP_TOT"])]).to(device)
    return torch.cumprod((1 - betas)[:-1], 0)
