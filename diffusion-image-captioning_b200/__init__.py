"""clipdlm-b200: B200-native (sm_100a) implementation of the CLIP-Diffusion-LM training step and denoise loop
(xu-shitong/diffusion-image-captioning, CLIP-DDPM.py) behind the reference's own Python call surface.

    import importlib; clipdlm = importlib.import_module("diffusion-image-captioning_b200")   # or: import clipdlm
    hp = clipdlm.default_hparams(BATCH_SIZE=512)
    model = clipdlm.DistilBertModel(embedding, projection, config, hp=hp)
    trainer = clipdlm.AdamW(model.parameters(), lr=hp["LEARNING_RATE"])
    l, x_t_loss, x_1_loss, prob_loss = clipdlm.train_func(model, trainer, batch)
    ids, restored = clipdlm.sample(model, image_clip, n_steps=5)

All compute runs in libclipdlm.so (hand-written CUDA, C-ABI in include/clipdlm.h); importing this package on a machine
without the built library or without a B200 raises at first use — there is no CPU fallback.
"""
from .hparams import GLOBALS, LOSS_KIND, alpha_cumprod, default_hparams, learning_rates, model_name, set_globals
from ._lib import ClipdlmError, EXPORTED_SYMBOLS, LIB_PATH


def __getattr__(name):  # torch-dependent modules load lazily so that `build` works before torch is paged in
    if name in ("DistilBertModel", "DistilBertConfig", "AdamW"):
        from . import model as _m
        return getattr(_m, name)
    if name in ("diffuse_t", "generate_diffuse_pair", "loss", "train_func", "validate", "train", "sample", "bind", "SAMPLE_USES_CUDA_GRAPH"):
        from . import diffusion as _d
        return getattr(_d, name)
    if name in ("DeviceCaptionDataset", "CaptionSubset", "CaptionLoader", "synthetic_dataset"):
        from . import data as _da
        return getattr(_da, name)
    if name in ("postprocess", "decode", "bleu_score"):
        from . import metrics as _me
        return getattr(_me, name)
    if name in ("load_reference_checkpoint", "save_reference_pickle", "to_reference_module", "save_checkpoint", "load_checkpoint"):
        from . import checkpoint as _c
        return getattr(_c, name)
    if name in ("enable_data_parallel", "init_process_group_from_env", "shard_range"):
        from . import parallel as _p
        return getattr(_p, name)
    raise AttributeError(name)
