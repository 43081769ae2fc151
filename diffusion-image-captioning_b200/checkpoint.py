"""Checkpoint interop with the reference.

The reference persists a trained model as a WHOLE-MODULE pickle — `torch.save(model.cpu(), f"{MODEL_NAME}.pickle")`
(CLIP-DDPM.py:551,560) — and reads it back with `torch.load(...)` (:506,570; COCO_BLEU.py:239), which only works inside a
script that defines `__main__.DistilBertModel`. This module moves weights both ways without that script:

  * `load_reference_checkpoint(path)`  any of {reference whole-module pickle, state_dict file, clipdlm checkpoint} -> state dict
    under the reference's parameter names (the names `DistilBertModel.state_dict()` of this package uses);
  * `to_reference_module(sd, hp)` / `save_reference_pickle(sd, hp, path)`  state dict -> a module tree pickled as
    `__main__.DistilBertModel`, which the reference's own `torch.load` call resolves to ITS class (methods come from the class,
    pickles carry only state), i.e. a model trained here can be evaluated by CLIP-DDPM.py / COCO_BLEU.py unchanged;
  * `save_checkpoint` / `load_checkpoint`  resume files of this package: model state dict + AdamW moments + step + hyperparameters.

Only tensors and containers move here; no compute. The HF version-skew hack the authors needed after loading an old pickle
(`model.model.add_module("activation", GELUActivation())`, COCO_BLEU.py:242) is irrelevant on the import side (only the state
dict is read) and unnecessary on the export side (the module tree is built with the transformers version that is installed).
"""
from __future__ import annotations

import io
import pickle
import sys
from typing import Dict, Optional, Tuple

import torch
from torch import nn

FORMAT = "clipdlm-checkpoint-v1"
_REF_CLASS = ("__main__", "DistilBertModel")


class ReferenceModuleShim(nn.Module):
    """Stands in for the reference's `class DistilBertModel(nn.Module)` while unpickling a whole-module checkpoint: it has the
    nn.Module state layout (so `_modules`, `_parameters` restore) and no behaviour."""

    def forward(self, *a, **kw):  # pragma: no cover - never called
        raise RuntimeError("ReferenceModuleShim only carries weights; load them into clipdlm.DistilBertModel")


class _RemapUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) == _REF_CLASS:
            main = sys.modules.get("__main__")
            real = getattr(main, name, None)
            if isinstance(real, type) and issubclass(real, nn.Module):
                return real  # running inside the reference script: keep its class
            return ReferenceModuleShim
        return super().find_class(module, name)


class _RemapPickle:
    """`pickle_module=` argument for torch.load: the stdlib pickle with the class remap above."""
    __name__ = "clipdlm_remap_pickle"
    Unpickler = _RemapUnpickler
    Pickler = pickle.Pickler
    load = staticmethod(lambda f, **kw: _RemapUnpickler(f, **kw).load())
    loads = staticmethod(lambda b, **kw: _RemapUnpickler(io.BytesIO(b), **kw).load())
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)


def _normalise(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Keep the tensors of the path (drop HF buffers such as position_ids and anything under the replaced in/out embeddings)."""
    out = {}
    for k, v in sd.items():
        if not torch.is_tensor(v):
            continue
        if k.endswith("position_ids") or "word_embeddings" in k or "vocab_projector" in k:
            continue
        out[k] = v.detach().to("cpu")
    return out


def _read(path, map_location="cpu", allow_pickle: bool = False):
    """torch.load with weights_only=True first: plain state dicts and this package's own resume files hold tensors and plain containers only
    and never need the unpickler. Only the reference's WHOLE-MODULE pickles do (arbitrary code runs on load, as with the reference's own
    `torch.load`, CLIP-DDPM.py:506,570) - that path is taken only with allow_pickle=True, i.e. for files the caller trusts."""
    try:
        return torch.load(path, map_location=map_location, weights_only=True)
    except Exception as ex:
        if not allow_pickle:
            raise pickle.UnpicklingError(
                f"{path}: not a weights-only file ({type(ex).__name__}). The reference's whole-module pickles execute code on load: "
                "pass allow_pickle=True for a file you trust") from ex
        if hasattr(path, "seek"):
            path.seek(0)
    return torch.load(path, map_location=map_location, pickle_module=_RemapPickle, weights_only=False)


def load_reference_checkpoint(path_or_obj, map_location="cpu", allow_pickle: bool = False) -> Dict[str, torch.Tensor]:
    """Reads a checkpoint and returns a CPU state dict under the reference's names:
    `model.distilbert...`, `model.vocab_transform.*`, `model.vocab_layer_norm.*`, `image_linear.*`, `text_linear.*`,
    `segment_embedding.weight` (concat fusion), `embedding.weight`, `lm_head.weight`, `lm_head.bias`
    (+ `input_projection.*` / `output_projection.*` when TRAIN_EMBEDDING, CLIP-DDPM.py:238-243).

    Accepted inputs: a path / file object holding (a) the reference's whole-module pickle (CLIP-DDPM.py:551,560),
    (b) a plain `state_dict` file, (c) a `save_checkpoint` file of this package; or the already-loaded object of any of those."""
    obj = path_or_obj
    if not isinstance(obj, (dict, nn.Module)):
        obj = _read(path_or_obj, map_location, allow_pickle)
    if isinstance(obj, nn.Module):
        return _normalise(obj.state_dict())
    if isinstance(obj, dict) and obj.get("format") == FORMAT:
        return _normalise(obj["model"])
    if isinstance(obj, dict):
        return _normalise(obj)
    raise TypeError(f"unsupported checkpoint object: {type(obj).__name__}")


def _hf_config(hp: dict):
    from transformers import DistilBertConfig
    return DistilBertConfig(n_layers=hp["N_LAYERS"], dim=hp["DIM"], n_heads=hp["N_HEADS"], hidden_dim=hp["HIDDEN_DIM"],
                            dropout=hp["DROPOUT"], attention_dropout=hp["ATTENTION_DROPOUT"], vocab_size=hp["VOCAB_SIZE"],
                            max_position_embeddings=hp["MAX_POSITION"])


def to_reference_module(sd: Dict[str, torch.Tensor], hp: dict, cls: Optional[type] = None) -> nn.Module:
    """Builds the module tree of the reference's `DistilBertModel.__init__` (CLIP-DDPM.py:227-256) on the CPU and fills it from
    `sd`. `cls` is the class the instance should have (default: `__main__.DistilBertModel` if the caller runs inside the
    reference script, else a behaviour-less stand-in that pickles under that name). Needs `transformers` (a dependency of the
    reference itself)."""
    from transformers import DistilBertForMaskedLM
    if cls is None:
        main_cls = getattr(sys.modules.get("__main__"), _REF_CLASS[1], None)
        cls = main_cls if isinstance(main_cls, type) and issubclass(main_cls, nn.Module) else _export_class()
    m = cls.__new__(cls)
    nn.Module.__init__(m)
    d, c, V, ch = hp["DIM"], hp["CLIP_DIM"], hp["VOCAB_SIZE"], hp["IN_CHANNEL"]
    m.model = DistilBertForMaskedLM(_hf_config(hp))
    if hp["TRAIN_EMBEDDING"]:  # :238-243
        m.embedding = nn.Embedding(V, ch)
        m.lm_head = nn.Linear(ch, V, bias=False)
        m.input_projection = nn.Linear(ch, d)
        m.output_projection = nn.Linear(d, ch)
    else:  # :245-247
        m.embedding = nn.Embedding(V, d).requires_grad_(False)
        m.lm_head = nn.Linear(d, V, bias=True).requires_grad_(False)
    m.model.set_input_embeddings(nn.Sequential())   # :249-250
    m.model.set_output_embeddings(nn.Sequential())
    m.image_linear = nn.Linear(c, d)
    m.text_linear = nn.Linear(c, d)
    if hp["CLIP_ADDING_METHOD"] == "concat":
        m.segment_embedding = nn.Embedding(2, d)
    own = m.state_dict()
    src = _normalise(sd)
    if "lm_head.bias" in own and "lm_head.bias" not in src:
        src["lm_head.bias"] = torch.zeros(V)  # :247
    missing = [k for k in own if k not in src and not k.endswith("position_ids")]
    if missing:
        raise KeyError(f"state dict lacks {missing[:4]}{'...' if len(missing) > 4 else ''}")
    with torch.no_grad():
        for k, v in own.items():
            if k in src:
                if tuple(v.shape) != tuple(src[k].shape):
                    raise ValueError(f"{k}: checkpoint shape {tuple(src[k].shape)} != model shape {tuple(v.shape)}")
                v.copy_(src[k].to(v.dtype))
    return m


_EXPORT_CLS = None


def _export_class() -> type:
    global _EXPORT_CLS
    if _EXPORT_CLS is None:
        _EXPORT_CLS = type(_REF_CLASS[1], (nn.Module,), {"__module__": _REF_CLASS[0], "__qualname__": _REF_CLASS[1],
                                                         "__doc__": "weights-only instance; behaviour comes from the reference's class on load"})
    return _EXPORT_CLS


def save_reference_pickle(sd: Dict[str, torch.Tensor], hp: dict, path) -> None:
    """Writes what the reference's `torch.save(model.cpu(), f"{MODEL_NAME}.pickle")` writes (CLIP-DDPM.py:551,560): a
    whole-module pickle whose class is recorded as `__main__.DistilBertModel`."""
    if hasattr(sd, "state_dict"):   # a model (e.g. `model.cpu()`, which returns the module itself) instead of its state dict
        sd = {k: v.detach().cpu() for k, v in sd.state_dict().items()}
    m = to_reference_module(sd, hp)
    cls = type(m)
    main = sys.modules["__main__"]
    had = hasattr(main, _REF_CLASS[1])
    prev = getattr(main, _REF_CLASS[1], None)
    setattr(main, _REF_CLASS[1], cls)  # pickle verifies that the recorded name resolves to this very class
    try:
        torch.save(m, path)
    finally:
        if had:
            setattr(main, _REF_CLASS[1], prev)
        else:
            delattr(main, _REF_CLASS[1])


def save_checkpoint(model, trainer, path, epoch: int = 0, extra: Optional[dict] = None) -> None:
    """Resume file of this package: reference-named model state dict, AdamW moments / step / param_groups, hyperparameters. The moments are
    always stored FULL-size with their `slice` and `n_params`. Under the fused data-parallel step (each rank holds 1/N of the moments) this
    is a collective: every rank calls it, the slices are gathered, rank 0 writes the file."""
    obj = dict(format=FORMAT, model={k: v.detach().cpu() for k, v in model.state_dict().items()}, hp=dict(model.hp), epoch=int(epoch),
               precision=model.precision, extra=extra or {})
    writer = True
    if trainer is not None:
        st = trainer.full_state_dict() if hasattr(trainer, "full_state_dict") else trainer.state_dict()
        obj["optimizer"] = dict(m=st["m"].cpu(), v=st["v"].cpu(), t=st["t"], param_groups=st["param_groups"])
        for k in ("slice", "n_params"):
            if k in st:
                obj["optimizer"][k] = st[k]
        if getattr(model, "dp_fused", None) is not None:
            import torch.distributed as dist
            writer = dist.get_rank(model.dp_group) == 0
    if writer:
        torch.save(obj, path)


def load_checkpoint(model, path_or_obj, trainer=None, strict: bool = True, allow_pickle: bool = False) -> Tuple[int, dict]:
    """Loads any supported checkpoint into `model` (and, for this package's own files, the optimizer state into `trainer`).
    Returns (epoch, extra). allow_pickle=True is needed (only) for the reference's whole-module pickles, see _read()."""
    obj = path_or_obj
    if not isinstance(obj, (dict, nn.Module)):
        obj = _read(path_or_obj, "cpu", allow_pickle)
    model.load_state_dict(load_reference_checkpoint(obj), strict=strict)
    if isinstance(obj, dict) and obj.get("format") == FORMAT:
        if trainer is not None and "optimizer" in obj:
            trainer.load_state_dict(obj["optimizer"])
        return int(obj.get("epoch", 0)), dict(obj.get("extra", {}))
    return 0, {}
