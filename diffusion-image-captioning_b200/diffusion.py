"""The reference's step functions over the native engine: `diffuse_t`, `generate_diffuse_pair`, `loss`, `train_func`,
`validate` (CLIP-DDPM.py:347-501) with the same names / argument meaning / return values, plus the thin `train()` (epoch loop,
:515-557) and `sample()` (denoise loop, :611-621) wrappers BASELINE.json's north star names.

Differences a maintainer should know (all documented in INTEGRATION.md):
  * the hyperparameters are a dict (`model.hp`) instead of module globals;
  * `loss()` is eager: the reference builds an autograd graph and `train_func` calls `l.backward()`; here the hand-written
    backward of each row chunk runs inside `loss()` (when `backward=True`) because activations are kept per chunk, and
    `train_func` only adds the optimizer step;
  * keyword-only `t=`, `noise_t=`, `noise_1=`, `dropout_seed=` pin the random draws for parity tests.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional

import torch

from . import _lib as L
from . import hparams as _hparams
from .hparams import LOSS_KIND, alpha_cumprod, learning_rates
from .model import AdamW, DistilBertModel

_ACP_CACHE: Dict[tuple, torch.Tensor] = {}


def _acp(hp: dict, device) -> torch.Tensor:
    key = (str(device), hp["COSIN_SCHEDULE"], hp["STEP_TOT"], hp["BETA_MIN"], hp["BETA_MAX"])
    if key not in _ACP_CACHE:
        _ACP_CACHE[key] = alpha_cumprod(hp, device).float().contiguous()
    return _ACP_CACHE[key]


def _coefs(hp: dict, t: torch.Tensor):
    a = _acp(hp, t.device)[t.reshape(-1)]
    return torch.sqrt(a).contiguous(), torch.sqrt(1 - a).contiguous()


def _need_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise L.ClipdlmError("clipdlm operates on CUDA tensors only (no CPU fallback)")


def diffuse_t(x: torch.Tensor, t: torch.Tensor, hp: Optional[dict] = None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q_sample, CLIP-DDPM.py:347-362, callable exactly as the reference's `diffuse_t(x, t)`. x [B, seq, C]; t [n] int64 (any shape with n
    elements) -> [n*B, seq, C], row = s*B + b; ONE noise draw of x.shape shared by the n samples (reference semantics). hp=None reads the
    active hyperparameters (hparams.GLOBALS - the reference's module globals); noise= pins the draw."""
    _need_cuda(x)
    hp = _hparams.GLOBALS if hp is None else hp
    B, seq, ch = x.shape
    S = t.numel()
    if noise is None:
        noise = torch.randn(x.shape, device=x.device)
    ca, cb = _coefs(hp, t.to(x.device))
    x32, n32 = x.float().contiguous(), noise.to(x.device, torch.float32).contiguous()
    assert tuple(n32.shape) == tuple(x32.shape), "diffuse_t: the noise draw has x's shape (CLIP-DDPM.py:359)"
    out = torch.empty(S * B, seq, ch, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().clipdlm_q_sample(L.ptr(x32), L.ptr(n32), L.ptr(ca), L.ptr(cb), x32.numel(), S, L.ptr(out),
                                          torch.cuda.current_stream(x.device).cuda_stream))
    return out


def generate_diffuse_pair(x_0, t, t_next=None, hp: Optional[dict] = None):
    """CLIP-DDPM.py:364-380, same positional form `generate_diffuse_pair(x_0, t, t_next)` (:467). x_0-prediction: (diffuse_t(x_0, t), x_0);
    otherwise a second, independently drawn noisy target at t_next."""
    if isinstance(t_next, dict):   # round-1 call form generate_diffuse_pair(x_0, t, hp, t_next)
        t_next, hp = (hp if not isinstance(hp, dict) else None), t_next
    hp = _hparams.GLOBALS if hp is None else hp
    if hp["X_0_PREDICTION"]:
        return diffuse_t(x_0, t, hp), x_0
    return diffuse_t(x_0, t, hp), diffuse_t(x_0, t_next, hp)


def bind(hp: Optional[dict] = None, val_loader=None, **overrides):
    """The reference's functions read module globals (hyperparameters :55-114, `val_loader` :221). `bind(hp, val_loader)` sets the active
    ones and returns the five step functions + the two wrappers, callable exactly as CLIP-DDPM.py:347,364,382,458,488 call them:

        F = clipdlm.bind(hp, val_loader)
        x_t = F.diffuse_t(x_0, t); pair = F.generate_diffuse_pair(x_0, t, t_next)
        l, x_t_loss, x_1_loss, prob_loss = F.train_func(model, trainer, x); val_x_t, val_x_1, val_prob = F.validate(model)
    """
    import types
    _hparams.set_globals(hp, val_loader, **overrides)
    return types.SimpleNamespace(hp=_hparams.GLOBALS, diffuse_t=diffuse_t, generate_diffuse_pair=generate_diffuse_pair, loss=loss,
                                 train_func=train_func, validate=validate, train=train, sample=sample)


def _next_seed() -> int:
    return int(torch.randint(0, 2 ** 62, (1,)).item())  # CPU generator: no device sync


def _loss_pass(model: DistilBertModel, eng, losses: torch.Tensor, slot: int, *, R: int, B: int, R_total: int, mode: int, ids32, mask32,
               image_clip, text_clip, backward: bool, seed: int, use_embed: bool, x_in=None, noise=None, coef_a=None, coef_b=None,
               target=None, target_rows: int = 0, guided: bool = False, cfg_rows: Optional[torch.Tensor] = None, eng_guided=None):
    """One encoder pass + its loss terms (+ backward). cfg_rows (int32 [R], 1 = classifier-free-guided row) switches on the
    guidance mix of CLIP-DDPM.py:313-317: a second, guided pass over the same rows in `eng_guided`, x_out mixed row-wise, the
    gradient of the mixed output split between the two passes ((1 + w) to the guided one, -w / 1 to the unguided one)."""
    hp = model.hp
    lib = L.load()
    if model.fused_softmax_grad and backward and slot == 0 and hp["USE_PROB_LOSS"]:
        model.refresh_exp_shift()   # the weights moved since the last step; every x_t chunk of this call reads the same device scalar
    fw = dict(R=R, B=B, mode=mode, train=model.training, image_clip=image_clip, text_clip=text_clip, attn_mask=mask32, x_in=x_in,
              ids=ids32, noise=noise, coef_a=coef_a, coef_b=coef_b)
    model._run_forward(eng, guided=guided, drop_seed=seed, **fw)
    s_self = s_exp = None
    if cfg_rows is not None:
        w = float(hp["CLASSIFIER_FREE_WEIGHT"])
        model._run_forward(eng_guided, guided=True, drop_seed=seed + 104729, **fw)   # the reference's second self.model(...) call draws its own dropout
        with torch.cuda.device(model.device):
            L.check(lib.clipdlm_engine_cfg_mix(eng, eng_guided, L.ptr(cfg_rows), w, model._stream()))
        if backward:
            gm = cfg_rows.to(torch.float32)
            s_self = (1.0 - (1.0 + w) * gm).contiguous()   # unguided pass: 1 on plain rows, -w on guided rows
            s_exp = ((1.0 + w) * gm).contiguous()          # guided pass: (1 + w) on guided rows, 0 elsewhere
    rw_host, rw_dev = model.rounding_weight()
    lc = L.LossCfg(LOSS_KIND[hp["LOSS_FUNC"]], 1 if use_embed else 0, 1 if hp["USE_PROB_LOSS"] else 0, hp["BATCH_SIZE"], R_total,
                   rw_host, 1 if backward else 0, L.ptr(target), target_rows,
                   L.ptr(s_self), L.ptr(s_exp), eng_guided if (cfg_rows is not None and backward) else None, L.ptr(rw_dev))
    with torch.cuda.device(model.device):
        L.check(lib.clipdlm_engine_loss_backward(eng, C.byref(lc), losses.data_ptr() + 8 * slot, model._stream()))
        if cfg_rows is not None and backward:
            L.check(lib.clipdlm_engine_backward(eng_guided, model._stream()))
    if backward:
        model._grads_dirty = True


def _finish(model, losses):
    hp = model.hp
    lf = losses.float()
    x_t_loss, x_1_loss = lf[0], lf[2]
    rw_host, rw_dev = model.rounding_weight()
    prob_loss = (rw_host if rw_dev is None else rw_dev.reshape(())) * (lf[1] + lf[3])  # CLIP-DDPM.py:445
    return x_t_loss, x_1_loss, prob_loss


def _prep_batch(model, image_clip, text_clip, mask, idx):
    dev = model.device
    return (image_clip.to(dev, torch.float32).contiguous(), text_clip.to(dev, torch.float32).contiguous(),
            (mask != 0).to(dev, torch.int32).contiguous(), idx.to(dev, torch.int32).contiguous())


def loss(model: DistilBertModel, x_t, x_1, x_tgt, x_0, image_clip, text_clip, mask, idx, loss_func=None, *, backward: Optional[bool] = None,
         dropout_seed: Optional[int] = None, classifier_mask: Optional[torch.Tensor] = None):
    """CLIP-DDPM.py:382-445. Same inputs / outputs: returns (x_t_loss, x_1_loss, ROUNDING_WEIGHT * (x_t_prob_loss + x_1_prob_loss))
    as 0-dim device tensors. `loss_func` is accepted for signature compatibility; the objective is `model.hp['LOSS_FUNC']`
    (a name) because the kernels implement the reference's four LOSS_FUNCs natively. backward=None => model.training and grad
    enabled: the gradients of `x_t_loss + x_1_loss + prob_loss` are accumulated into the model's flat grad buffer."""
    hp = model.hp
    S, B, ML, D = hp["SAMPLE_SIZE"], hp["BATCH_SIZE"], hp["MAX_LENGTH"], hp["IN_CHANNEL"]
    assert tuple(x_t.shape) == (S * B, ML, D)  # :396-400
    assert tuple(x_1.shape) == tuple(x_0.shape) == (B, ML, D)
    assert tuple(image_clip.shape) == tuple(text_clip.shape) == (B, hp["CLIP_DIM"])
    assert tuple(mask.shape) == (B, ML) and tuple(idx.shape) == (B, ML)
    if loss_func is not None and getattr(loss_func, "__name__", loss_func) != hp["LOSS_FUNC"]:
        raise ValueError(f"loss_func {getattr(loss_func, '__name__', loss_func)} differs from hp['LOSS_FUNC'] = {hp['LOSS_FUNC']}")
    _need_cuda(x_t)
    if hp["TRAIN_EMBEDDING"]:
        if backward is None:
            backward = model.training and torch.is_grad_enabled()
        return _loss_te(model, x_t, x_1, x_tgt, x_0, image_clip, text_clip, mask, idx, backward=backward, dropout_seed=dropout_seed,
                        classifier_mask=classifier_mask)
    cfg = None
    if hp["CLASSIFIER_FREE_WEIGHT"] > 0:  # :406-410: per-row guidance draw; rows 0 / 1 pinned so that both kinds always occur
        if classifier_mask is None:
            classifier_mask = (torch.rand((S * B, 1)) > hp["CLASSIFIER_FREE_PROB"]).to(torch.float32)
            classifier_mask[0] = 0
            classifier_mask[1] = 1
        cfg = (classifier_mask.reshape(-1) != 0).to(model.device, torch.int32).contiguous()
        assert cfg.numel() == S * B
    if backward is None:
        backward = model.training and torch.is_grad_enabled()
    img, txt, mask32, ids32 = _prep_batch(model, image_clip, text_clip, mask, idx)
    x_t32, x_132, x_032 = x_t.float().contiguous(), x_1.float().contiguous(), x_0.float().contiguous()
    tgt_t = x_032 if hp["X_0_PREDICTION"] else x_tgt.float().contiguous()
    if not hp["X_0_PREDICTION"]:
        assert tuple(x_tgt.shape) == tuple(x_t.shape)  # :420
    spc = max(1, model.chunk_rows // B)
    eng = model._engine(min(S, spc) * B, B, backward or model.training)
    eng_g = model._engine(min(S, spc) * B, B, backward or model.training, tag="g") if cfg is not None else None
    losses = torch.zeros(4, dtype=torch.float64, device=model.device)
    seed = _next_seed() if dropout_seed is None else int(dropout_seed)
    for ci, s0 in enumerate(range(0, S, spc)):
        s1 = min(S, s0 + spc)
        R = (s1 - s0) * B
        xin = x_t32[s0 * B:s1 * B]
        if hp["X_0_PREDICTION"]:
            target, trows = tgt_t, B
        else:
            target, trows = tgt_t[s0 * B:s1 * B], R
        _loss_pass(model, eng, losses, 0, R=R, B=B, R_total=S * B, mode=0, ids32=ids32, mask32=mask32, image_clip=img, text_clip=txt,
                   backward=backward, seed=seed + ci, use_embed=hp["USE_X_T_LOSS"], x_in=xin, target=target, target_rows=trows,
                   cfg_rows=cfg[s0 * B:s1 * B].contiguous() if cfg is not None else None, eng_guided=eng_g)
    _loss_pass(model, eng, losses, 2, R=B, B=B, R_total=B, mode=0, ids32=ids32, mask32=mask32, image_clip=img, text_clip=txt,
               backward=backward, seed=seed + 7919, use_embed=hp["USE_X_1_LOSS"], x_in=x_132, target=x_032, target_rows=B)
    return _finish(model, losses)


def _loss_te(model: DistilBertModel, x_t, x_1, x_tgt, x_0, image_clip, text_clip, mask, idx, *, backward: bool, dropout_seed=None,
             embed_coefs=None, classifier_mask=None):
    """loss() for TRAIN_EMBEDDING=True (train_embedding.py). The inputs x_t / x_1 / x_tgt / x_0 are functions of the trainable
    embedding; `embed_coefs = (ca_t [S], ca_1 [1], ca_tgt [S] or None)` (the sqrt(alpha_bar) factors of q_sample, CLIP-DDPM.py:360) lets
    the gradients that reach them be folded into d(embedding.weight) chunk by chunk. Called without it (a direct loss() call on
    explicit tensors) every other parameter gradient is still produced, exactly as autograd would treat detached inputs."""
    from . import train_embedding as TE
    hp = model.hp
    S, B, ML, ch = hp["SAMPLE_SIZE"], hp["BATCH_SIZE"], hp["MAX_LENGTH"], hp["IN_CHANNEL"]
    cfg = None
    if hp["CLASSIFIER_FREE_WEIGHT"] > 0:  # :406-410, as in loss()
        if classifier_mask is None:
            classifier_mask = (torch.rand((S * B, 1)) > hp["CLASSIFIER_FREE_PROB"]).to(torch.float32)
            classifier_mask[0] = 0
            classifier_mask[1] = 1
        cfg = (classifier_mask.reshape(-1) != 0).to(model.device, torch.int32).contiguous()
        assert cfg.numel() == S * B
    img, txt, mask32, ids32 = _prep_batch(model, image_clip, text_clip, mask, idx)
    x_t32, x_132, x_032 = x_t.float().contiguous(), x_1.float().contiguous(), x_0.float().contiguous()
    x0pred = bool(hp["X_0_PREDICTION"])
    tgt_t = x_032 if x0pred else x_tgt.float().contiguous()
    if not x0pred:
        assert tuple(x_tgt.shape) == tuple(x_t.shape)  # :420
    spc = max(1, model.chunk_rows // B)
    eng = model._engine(min(S, spc) * B, B, backward or model.training)
    eng_g = model._engine(min(S, spc) * B, B, backward or model.training, tag="g") if cfg is not None else None
    losses = torch.zeros(4, dtype=torch.float64, device=model.device)
    seed = _next_seed() if dropout_seed is None else int(dropout_seed)
    d_x0 = torch.zeros_like(x_032) if backward else None
    for ci, s0 in enumerate(range(0, S, spc)):
        s1 = min(S, s0 + spc)
        R = (s1 - s0) * B
        if x0pred:
            target, trows, d_tgt = tgt_t, B, d_x0
        else:
            target, trows = tgt_t[s0 * B:s1 * B], R
            d_tgt = torch.zeros(R, ML, ch, device=model.device) if backward else None
        dx = TE.loss_pass(model, eng, losses, 0, x16=x_t32[s0 * B:s1 * B], R=R, B=B, R_total=S * B, target=target, target_rows=trows, img=img,
                          txt=txt, mask32=mask32, ids32=ids32, seed=seed + ci, use_embed=hp["USE_X_T_LOSS"], backward=backward, d_target=d_tgt,
                          cfg_rows=cfg[s0 * B:s1 * B].contiguous() if cfg is not None else None, eng_guided=eng_g)
        if backward and embed_coefs is not None:
            TE.embedding_bwd(model, dx, embed_coefs[0][s0:s1].contiguous(), ids32, s1 - s0)
            if not x0pred and hp["USE_X_T_LOSS"]:
                TE.embedding_bwd(model, d_tgt, embed_coefs[2][s0:s1].contiguous(), ids32, s1 - s0)
    dx = TE.loss_pass(model, eng, losses, 2, x16=x_132, R=B, B=B, R_total=B, target=x_032, target_rows=B, img=img, txt=txt, mask32=mask32,
                      ids32=ids32, seed=seed + 7919, use_embed=hp["USE_X_1_LOSS"], backward=backward, d_target=d_x0)
    if backward and embed_coefs is not None:
        TE.embedding_bwd(model, dx, embed_coefs[1].contiguous(), ids32, 1)
        TE.embedding_bwd(model, d_x0, None, ids32, 1)  # the loss targets x_0.repeat(...) / x_0 (:418,428) are the embedding rows themselves
    return _finish(model, losses)


def _train_func_te(model: DistilBertModel, trainer, x: dict, train: bool, t, noise_t, noise_1, noise_tgt, dropout_seed, classifier_mask=None):
    """train_func (CLIP-DDPM.py:458-486) for TRAIN_EMBEDDING=True: x_0 = model.embedding(ids) carries gradient."""
    hp = model.hp
    dev = model.device
    S, B, ML = hp["SAMPLE_SIZE"], hp["BATCH_SIZE"], hp["MAX_LENGTH"]
    ids = x["input_ids"].to(dev)
    x_0 = model.embedding(ids).contiguous()  # :459
    ca, _ = _coefs(hp, t)
    one = torch.ones(1, dtype=torch.int64, device=dev)
    ca1, _ = _coefs(hp, one)
    x_t = diffuse_t(x_0, t, hp, noise_t)  # :464
    x_tgt, ca_n = None, None
    if not hp["X_0_PREDICTION"]:  # :466-467
        t_next = torch.max(t - hp["X_T_STEP_INTERVAL"], torch.zeros_like(t))
        x_tgt = diffuse_t(x_0, t_next, hp, noise_tgt)
        ca_n, _ = _coefs(hp, t_next)
    x_1 = diffuse_t(x_0, one, hp, noise_1)  # :468
    return _loss_te(model, x_t, x_1, x_tgt, x_0, x["image_clip"], x["text_clip"], x["attention_mask"], ids, backward=bool(train),
                    dropout_seed=dropout_seed, embed_coefs=(ca, ca1, ca_n), classifier_mask=classifier_mask)


def train_func(model: DistilBertModel, trainer: Optional[AdamW], x: dict, train: bool = True, *, t: Optional[torch.Tensor] = None,
               noise_t: Optional[torch.Tensor] = None, noise_1: Optional[torch.Tensor] = None, dropout_seed: Optional[int] = None,
               noise_tgt: Optional[torch.Tensor] = None, classifier_mask: Optional[torch.Tensor] = None):
    """CLIP-DDPM.py:458-486: embed -> draw t -> q_sample x2 -> zero_grad -> loss -> backward -> AdamW step.
    Returns (l, x_t_loss, x_1_loss, prob_loss) as 0-dim device tensors (no host sync, like the reference's loop :530-533).

    With X_0_PREDICTION (the default) nothing of shape [S*B, 16, 768] is materialised: the fused prologue kernel gathers
    E[ids], applies sqrt(abar_t) / sqrt(1 - abar_t) per sample and the CLIP fusion in one pass, chunk by chunk."""
    hp = model.hp
    dev = model.device
    S, B, ML, D = hp["SAMPLE_SIZE"], hp["BATCH_SIZE"], hp["MAX_LENGTH"], hp["IN_CHANNEL"]
    ids = x["input_ids"].to(dev)
    assert tuple(ids.shape) == (B, ML), f"batch must be exactly BATCH_SIZE={B} x MAX_LENGTH={ML} (reference asserts :396-400)"
    if t is None:
        t = torch.randint(0, hp["STEP_TOT"], (S, 1, 1), device=dev)  # :461
        if model.dp_group is not None:  # the reference shares t across the whole batch; keep that across ranks
            torch.distributed.broadcast(t, src=0, group=model.dp_group)
    t = t.to(dev)
    assert t.numel() == S
    backward = bool(train)
    if train:
        trainer.zero_grad()  # :471
    if hp["TRAIN_EMBEDDING"]:
        x_t_loss, x_1_loss, prob_loss = _train_func_te(model, trainer, x, train, t, noise_t, noise_1, noise_tgt, dropout_seed, classifier_mask)
    elif not hp["X_0_PREDICTION"] or hp["CLASSIFIER_FREE_WEIGHT"] > 0:
        # x_{t-1}-prediction objective (:466-467): explicit tensors through loss()
        x_0 = model.embedding(ids)
        t_next = torch.max(t - hp["X_T_STEP_INTERVAL"], torch.zeros_like(t))
        x_t = diffuse_t(x_0, t, hp, noise_t)
        x_tgt = diffuse_t(x_0, t_next, hp, noise_tgt)
        x_1 = diffuse_t(x_0, torch.ones(1, dtype=torch.int64, device=dev), hp, noise_1)
        x_t_loss, x_1_loss, prob_loss = loss(model, x_t, x_1, x_tgt, x_0, x["image_clip"], x["text_clip"], x["attention_mask"], ids,
                                             backward=backward, dropout_seed=dropout_seed, classifier_mask=classifier_mask)
    else:
        img, txt, mask32, ids32 = _prep_batch(model, x["image_clip"], x["text_clip"], x["attention_mask"], ids)
        if noise_t is None:
            noise_t = torch.randn(B, ML, D, device=dev)  # one draw shared by the S samples (:359)
        if noise_1 is None:
            noise_1 = torch.randn(B, ML, D, device=dev)
        noise_t, noise_1 = noise_t.to(dev, torch.float32).contiguous(), noise_1.to(dev, torch.float32).contiguous()
        ca, cb = _coefs(hp, t)
        ca1, cb1 = _coefs(hp, torch.ones(1, dtype=torch.int64, device=dev))  # x_1 = diffuse_t(x_0, ones(1)) :468
        spc = max(1, model.chunk_rows // B)
        eng = model._engine(min(S, spc) * B, B, backward or model.training)
        losses = torch.zeros(4, dtype=torch.float64, device=dev)
        seed = _next_seed() if dropout_seed is None else int(dropout_seed)
        for ci, s0 in enumerate(range(0, S, spc)):
            s1 = min(S, s0 + spc)
            _loss_pass(model, eng, losses, 0, R=(s1 - s0) * B, B=B, R_total=S * B, mode=1, ids32=ids32, mask32=mask32, image_clip=img,
                       text_clip=txt, backward=backward, seed=seed + ci, use_embed=hp["USE_X_T_LOSS"], noise=noise_t, coef_a=ca[s0:s1],
                       coef_b=cb[s0:s1])
        _loss_pass(model, eng, losses, 2, R=B, B=B, R_total=B, mode=1, ids32=ids32, mask32=mask32, image_clip=img, text_clip=txt,
                   backward=backward, seed=seed + 7919, use_embed=hp["USE_X_1_LOSS"], noise=noise_1, coef_a=ca1, coef_b=cb1)
        x_t_loss, x_1_loss, prob_loss = _finish(model, losses)
    l = x_t_loss + x_1_loss + prob_loss  # :481
    if train:
        if model.dp_group is not None and model.dp_fused is None:
            torch.distributed.all_reduce(model.grad, group=model.dp_group)  # sum; AdamW applies 1/world
        # (fused mode: trainer.step() is reduce-scatter + AdamW + all-gather in one kernel over NVLink peer memory, csrc/dp_fused.cu)
        trainer.step()  # :484
    return l, x_t_loss, x_1_loss, prob_loss


def validate(model: DistilBertModel, val_loader: Optional[Iterable[dict]] = None):
    """CLIP-DDPM.py:488-501: eval mode, no grad, mean of the three loss terms over the validation loader. Callable as the reference's
    `validate(model)`: the loader then is `model.val_loader` if set, else the one registered with `bind(hp, val_loader)` (the reference reads
    the module global `val_loader`, :221,495)."""
    if val_loader is None:
        val_loader = getattr(model, "val_loader", None) or _hparams.ACTIVE["val_loader"]
    if val_loader is None:
        raise ValueError("validate(model): no validation loader - pass one, set model.val_loader, or register it with clipdlm.bind(hp, val_loader)")
    was_training = model.training
    model.eval()
    x_t_loss = x_1_loss = prob_loss = 0
    n = 0
    with torch.no_grad():
        for x in val_loader:
            _, a, b, c = train_func(model, None, x, train=False)
            x_t_loss, x_1_loss, prob_loss = x_t_loss + a, x_1_loss + b, prob_loss + c
            n += 1
    model.train(was_training)
    n = max(n, 1)
    return x_t_loss / n, x_1_loss / n, prob_loss / n


def train(model: DistilBertModel, trainer: AdamW, train_loader, hp: Optional[dict] = None, val_loader=None, summary=None, on_early_stop=None):
    """The epoch loop of CLIP-DDPM.py:515-557 as a function: per-epoch learning rate from `lrs` (:451-456,520-522), the
    dynamic rounding weight (:535-536), validation + the early-stop hook (:546-553), one summary line per epoch (:554).
    Returns a list of per-epoch dicts. Loss accumulators stay on the device (no per-step sync)."""
    hp = model.hp if hp is None else hp
    lrs = learning_rates(hp)
    early_stopped = False
    model.train()
    history = []
    dp = getattr(model, "dp_group", None) is not None

    def _global_mean(*xs):
        """Under data parallelism every rank must take the same decisions (rounding weight, early stop): mean of the rank-local values."""
        if not dp:
            return xs
        v = torch.stack([torch.as_tensor(x, dtype=torch.float32, device=model.device) for x in xs])
        torch.distributed.all_reduce(v, group=model.dp_group)
        return tuple(v / model.dp_world)

    for epoch in range(hp["EPOCH_NUM"]):
        acc_x_t = acc_x_1 = acc_prob = acc_l = 0
        if not hp["END_LEARNING_RATE"] == hp["LEARNING_RATE"]:
            for g in trainer.param_groups:
                g["lr"] = lrs[epoch]
        n_batches = 0
        for x in train_loader:
            l, x_t_loss, x_1_loss, prob_loss = train_func(model, trainer, x)
            acc_x_t, acc_x_1, acc_prob, acc_l = acc_x_t + x_t_loss, acc_x_1 + x_1_loss, acc_prob + prob_loss, acc_l + l
            n_batches += 1
            if hp["DYNAMIC_ROUNDING_WEIGHT"] > 0:  # :535-536: a device tensor in the reference, a device tensor here - no read-back, the kernels multiply by it
                g_x_t, g_x_1, g_prob = _global_mean(acc_x_t, acc_x_1, acc_prob)
                rw = ((g_x_t + g_x_1) / g_prob).detach() * hp["DYNAMIC_ROUNDING_WEIGHT"]
                if torch.is_tensor(rw) and rw.is_cuda:
                    buf = model.__dict__.get("_rw_dev")
                    if buf is None:
                        buf = model.__dict__["_rw_dev"] = torch.empty(1, device=rw.device)   # one address for the life of the model
                    buf.copy_(rw.float().reshape(1))
                    model.hp["ROUNDING_WEIGHT"] = buf
                else:   # host-side stand-ins (tests of the loop logic)
                    model.hp["ROUNDING_WEIGHT"] = float(rw)
            if hp["DEBUG"]:
                break
        # the reference divides by len(train_loader) (:547,554), also when DEBUG cut the epoch after one batch
        n_batches = max(len(train_loader) if hasattr(train_loader, "__len__") else n_batches, 1)
        acc_x_t, acc_x_1, acc_prob, acc_l = _global_mean(acc_x_t, acc_x_1, acc_prob, acc_l)
        rec = dict(epoch=epoch, x_t_loss=acc_x_t / n_batches, x_1_loss=acc_x_1 / n_batches, prob_loss=acc_prob / n_batches, lr=trainer.param_groups[0]["lr"])
        if val_loader is not None:
            val_x_t, val_x_1, val_prob = _global_mean(*validate(model, val_loader))
            rec.update(val_x_t=val_x_t, val_x_1=val_x_1, val_prob=val_prob)
            if val_x_t + val_x_1 + val_prob > hp["EARLY_STOP_RATIO"] * acc_l / n_batches:
                if not early_stopped:
                    if summary is not None:
                        summary.write("early stop! \n")  # :549
                    if on_early_stop is not None:
                        on_early_stop(model, epoch)  # the reference saves the whole-module pickle here (:550-551)
                early_stopped = True
        rec["early_stopped"] = early_stopped
        if summary is not None:
            summary.write(f"epoch {epoch} average x_t_loss, x_1_loss, prob_loss, val losses: {float(rec['x_t_loss'])}, {float(rec['x_1_loss'])}, "
                          f"{float(rec['prob_loss'])}, {float(rec.get('val_x_t', float('nan')))}, {float(rec.get('val_x_1', float('nan')))}, "
                          f"{float(rec.get('val_prob', float('nan')))}\n")
        history.append(rec)
        if hp["DEBUG"]:
            break
    return history


SAMPLE_USES_CUDA_GRAPH = True
_GRAPH_MAX_BATCH = 256     # above this the loop is long kernels back to back and launch latency is hidden anyway


def _denoise_steps(model, eng, cur, nxt, img, txt, n_steps: int, return_all: bool):
    """n_steps x { encoder pass on restored[:, :MAX_LENGTH] } (+ fused lm_head / arg-max where ids are needed). Pure launches on the current
    stream over caller-owned buffers: this is the body the CUDA graph of the small-batch path captures."""
    hp = model.hp
    B = cur.shape[0]
    Lfull, D = cur.shape[1], cur.shape[2]
    outs = []
    for i in range(n_steps):
        model._run_forward(eng, R=B, B=B, mode=0, guided=False, train=False, image_clip=img, text_clip=txt, attn_mask=None, x_in=cur,
                           x_in_stride=Lfull * D, x_out=nxt, reuse_proj=i > 0)  # the CLIP projections do not change across steps
        cur, nxt = nxt, cur
        if return_all:
            outs.append(model.argmax_last(B))
    indexes = outs[-1] if return_all and outs else model.argmax_last(B)  # softmax is monotone: argmax(softmax(x)) == argmax(x) (:620)
    return indexes, cur, outs


@torch.no_grad()
def sample(model: DistilBertModel, image_clip: torch.Tensor, n_steps: int = 5, return_all: bool = False,
           restored: Optional[torch.Tensor] = None, unique_consecutive: bool = False, use_graph: Optional[bool] = None):
    """The reference's reverse "denoise" loop (CLIP-DDPM.py:611-621, COCO_BLEU.py:249-257): restored ~ N(0, I) [B, L, C];
    n_steps x { out, restored = model(restored[:, :MAX_LENGTH], image_clip, 0, ones, [1, 0]) }; ids = argmax over the vocabulary.

    Returns ids int64 [B, MAX_LENGTH] (and the final `restored`, and per-step ids if return_all). The lm_head GEMM runs only
    where ids are needed (last step; every step if return_all) with a fused running-argmax epilogue: logits never reach HBM.
    unique_consecutive=True applies the reference's `indexes.unique_consecutive(dim=-1)` post-processing (:621).

    Small batches (the reference evaluates B = 8 x 5 steps, :613-617) are launch-bound - ~45 kernels of a few microseconds per step - so for
    B <= 256 the whole loop is captured ONCE per (B, n_steps) into a CUDA graph over static buffers and replayed (use_graph=False opts out;
    None = on unless profiling)."""
    hp = model.hp
    dev = model.device
    _need_cuda(image_clip)
    B = image_clip.shape[0]
    ML, D = hp["MAX_LENGTH"], hp["IN_CHANNEL"]
    Lfull = ML + (2 if hp["CLIP_ADDING_METHOD"] == "concat" else 0)
    if restored is None:
        restored = torch.randn(B, Lfull, D, device=dev)  # :613
    img = image_clip.to(dev, torch.float32).contiguous()
    eng = model._engine(B, B, False)
    model._last_eng = eng
    outs = []
    if hp["TRAIN_EMBEDDING"]:  # the loop runs in the IN_CHANNEL-wide space: input_projection -> encoder -> output_projection per step
        from . import train_embedding as TE
        cur = restored.to(dev, torch.float32).contiguous().clone()
        txt = torch.zeros_like(img)  # text_clip = zeros (:617)
        xo = torch.empty(B, Lfull, hp["DIM"], device=dev)
        for i in range(n_steps):
            u = TE.in_proj(model, cur[:, :ML].contiguous())
            model._run_forward(eng, R=B, B=B, mode=0, guided=False, train=False, image_clip=img, text_clip=txt, attn_mask=None, x_in=u,
                               x_out=xo, reuse_proj=i > 0)
            cur = TE.out_proj(model, xo)
            if return_all:
                outs.append(TE.argmax(model, cur))
        indexes = outs[-1] if return_all and outs else TE.argmax(model, cur)
        if unique_consecutive:
            indexes = indexes.unique_consecutive(dim=-1)
        cur = cur.clone()
        return (indexes, cur, outs) if return_all else (indexes, cur)
    if use_graph is None:
        use_graph = SAMPLE_USES_CUDA_GRAPH and B <= _GRAPH_MAX_BATCH and not getattr(model, "_profiling", False)
    if use_graph:
        indexes, cur, outs = _sample_graphed(model, eng, restored, img, n_steps, return_all)
    else:
        cur = restored.to(dev, torch.float32).contiguous().clone()
        indexes, cur, outs = _denoise_steps(model, eng, cur, torch.empty_like(cur), img, torch.zeros_like(img), n_steps, return_all)
    if unique_consecutive:
        indexes = indexes.unique_consecutive(dim=-1)
    return (indexes, cur, outs) if return_all else (indexes, cur)


def _sample_graphed(model, eng, restored, img, n_steps: int, return_all: bool):
    """Replay (capture on first use) of the CUDA graph of the denoise loop for this (engine, B, n_steps, return_all)."""
    dev = model.device
    B = img.shape[0]
    cache = model.__dict__.setdefault("_sample_graphs", {})
    key = (int(eng), B, int(n_steps), bool(return_all), tuple(restored.shape))
    ent = cache.get(key)
    if ent is None:
        for k in [k for k in cache if k[0] != int(eng)]:   # the engine was regrown: its workspace moved, older graphs are stale
            del cache[k]
        st_cur = restored.to(dev, torch.float32).contiguous().clone()
        st_img = img.clone()
        st_txt = torch.zeros_like(st_img)
        st_nxt = torch.empty_like(st_cur)
        keep = st_cur.clone()
        _denoise_steps(model, eng, st_cur, st_nxt, st_img, st_txt, min(n_steps, 2), return_all)   # un-captured warm-up: lazy one-time initialisation
        torch.cuda.synchronize(dev)
        st_cur.copy_(keep)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ids, cur, outs = _denoise_steps(model, eng, st_cur, st_nxt, st_img, st_txt, n_steps, return_all)
        ent = cache[key] = dict(graph=graph, cur_in=st_cur, img=st_img, ids=ids, cur_out=cur, outs=outs, keep=(st_nxt, st_txt))
    ent["cur_in"].copy_(restored.to(dev, torch.float32))
    ent["img"].copy_(img)
    ent["graph"].replay()
    return ent["ids"].clone(), ent["cur_out"].clone(), [o.clone() for o in ent["outs"]]
