"""Runs the REAL reference (xu-shitong/diffusion-image-captioning, CLIP-DDPM.py) for the hot path by exec'ing its own source
slices from /root/reference (read-only, never copied) with the hyperparameter globals injected — SURVEY.md §8(c) recipe.

TEST INFRASTRUCTURE ONLY. Used (a) by oracle/validate_against_reference.py to pin the restatement in oracle/clipdlm_oracle.py,
(b) by tests/golden/make_golden.py to generate the committed fixtures. /root/reference does not exist on the GPU box: nothing
that runs there imports this file.
"""
from __future__ import annotations

import copy
import math
import os

import torch
from torch import nn

REF = os.environ.get("CLIPDLM_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "CLIP-DDPM.py")


def available() -> bool:
    try:
        import transformers  # noqa: F401
    except Exception:
        return False
    return os.path.exists(SRC)


def build_namespace(hp: dict, device="cpu") -> dict:
    """Namespace holding the reference's DistilBertModel / diffuse_t / loss / train_func, exec'd from its own file.
    0-indexed line slices: [226:323] class DistilBertModel, [336:446] schedule + diffuse_t + generate_diffuse_pair + loss,
    [457:487] train_func."""
    from transformers import DistilBertConfig, DistilBertForMaskedLM
    lines = open(SRC).read().split("\n")
    ns = dict(torch=torch, nn=nn, math=math, copy=copy, DistilBertForMaskedLM=DistilBertForMaskedLM, DistilBertConfig=DistilBertConfig,
              device=torch.device(device))
    for k in ("DEBUG", "CONTINUE_TRAIN", "BATCH_SIZE", "MAX_LENGTH", "LEARNING_RATE", "END_LEARNING_RATE", "TRAIN_SET_RATIO",
              "EARLY_STOP_RATIO", "EPOCH_NUM", "DYNAMIC_ROUNDING_WEIGHT", "ROUNDING_WEIGHT", "CLIP_ADDING_METHOD", "CLASSIFIER_FREE_WEIGHT",
              "CLASSIFIER_FREE_PROB", "TRAIN_EMBEDDING", "IN_CHANNEL", "BETA_MIN", "BETA_MAX", "STEP_TOT", "COSIN_SCHEDULE", "SAMPLE_SIZE",
              "X_0_PREDICTION", "X_T_STEP_INTERVAL", "USE_X_T_LOSS", "USE_X_1_LOSS", "USE_PROB_LOSS", "VOCAB_SIZE"):
        ns[k] = hp[k]
    exec("\n".join(lines[76:88]), ns)  # the four LOSS_FUNC definitions (:77-87)
    ns["LOSS_FUNC"] = ns[hp["LOSS_FUNC"]]
    cls_src = "\n".join(lines[226:323])
    if hp["DIM"] != 768:
        cls_src = cls_src.replace("768", str(hp["DIM"]))  # the class hard-codes 768 at :252-253,256
    exec(cls_src, ns)
    exec("\n".join(lines[336:446]), ns)
    exec("\n".join(lines[457:487]), ns)
    return ns


def build_model(ns: dict, hp: dict, seed: int = 0):
    """Reference construction order with the random-init stand-in for the pretrained checkpoint."""
    from transformers import DistilBertConfig, DistilBertForMaskedLM
    torch.manual_seed(seed)
    cfg = DistilBertConfig(n_layers=hp["N_LAYERS"], dim=hp["DIM"], n_heads=hp["N_HEADS"], hidden_dim=hp["HIDDEN_DIM"],
                           dropout=hp["DROPOUT"], attention_dropout=hp["ATTENTION_DROPOUT"], vocab_size=hp["VOCAB_SIZE"])
    if hp["TRAIN_EMBEDDING"]:  # CLIP-DDPM.py:325-327
        return ns["DistilBertModel"](config=cfg)
    origin = DistilBertForMaskedLM(cfg)
    model = ns["DistilBertModel"](origin.get_input_embeddings(), origin.get_output_embeddings(), cfg)
    return model


def export_params(model) -> dict:
    """Reference parameters under the names oracle/clipdlm_oracle.py uses."""
    P = {}
    for n, p in model.named_parameters():
        if n.startswith(("model.", "image_linear", "text_linear", "segment_embedding", "input_projection", "output_projection")) and "vocab_projector" not in n:
            P[n] = p.detach().clone()
    P["embedding.weight"] = model.embedding.weight.detach().clone()
    if hasattr(model, "input_projection"):  # TRAIN_EMBEDDING: the lm_head is its own trainable tensor (:239)
        P["lm_head.weight"] = model.lm_head.weight.detach().clone()
    return P


def load_params(model, P: dict):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n in P:
                p.copy_(P[n])
        model.embedding.weight.copy_(P["embedding.weight"])
        model.lm_head.weight.copy_(P.get("lm_head.weight", P["embedding.weight"]))
