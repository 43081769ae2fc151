"""TRAIN_EMBEDDING=True branch (CLIP-DDPM.py:98-102,238-243,292-293,319-320): the diffusion runs in a learned IN_CHANNEL-wide (16)
embedding space; `input_projection` (16 -> 768) and `output_projection` (768 -> 16) wrap the encoder, and `embedding`, `lm_head`
(nn.Linear(16, V, bias=False)) and both projections are trainable.

Host-side composition over the C-ABI (include/clipdlm.h):
  * the encoder is the native engine (clipdlm_engine_forward / clipdlm_engine_backward_from);
  * the K = 16 / N = 16 projections are the fp32 small-linear kernels (forward, and backward via the transposed weight);
  * the lm_head (logits, LSE / argmax, softmax-CE gradient, dgrad, wgrad) is the tcgen05 GEMM on operands zero-padded to CH_PAD = 64
    channels (one 128-byte K block). The lm_head weight lives in the flat parameter buffer already padded ([V rounded up to 256, 64];
    the padding has zero gradient, so AdamW keeps it zero) — its bf16 shadow is therefore a ready GEMM operand after every step;
  * narrow-channel glue kernels: clipdlm_feature_loss_f32, clipdlm_pack_rows_bf16, clipdlm_embedding_bwd.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .hparams import LOSS_KIND

CH_PAD = 64


def extra_params(hp: dict):
    """(reference name, stored shape, logical shape) of the tensors TRAIN_EMBEDDING adds, in the order of the reference's
    parameters() (CLIP-DDPM.py:260-262)."""
    ch, V, D = hp["IN_CHANNEL"], hp["VOCAB_SIZE"], hp["DIM"]
    if ch % 4 != 0 or ch > CH_PAD:
        raise ValueError(f"IN_CHANNEL must be a multiple of 4, <= {CH_PAD} (got {ch})")
    vpad = (V + 255) // 256 * 256
    return [("embedding.weight", (V, ch), (V, ch)), ("lm_head.weight", (vpad, CH_PAD), (V, ch)),
            ("input_projection.weight", (D, ch), (D, ch)), ("input_projection.bias", (D,), (D,)),
            ("output_projection.weight", (ch, D), (ch, D)), ("output_projection.bias", (ch,), (ch,))]


def _dims(model):
    hp = model.hp
    ML = hp["MAX_LENGTH"]
    return ML, ML + (2 if hp["CLIP_ADDING_METHOD"] == "concat" else 0), hp["IN_CHANNEL"], hp["DIM"], hp["VOCAB_SIZE"]


def _lin_fwd(model, x, w, b, rows, K, N, out):
    with torch.cuda.device(model.device):
        L.check(L.load().clipdlm_small_linear_fwd(L.ptr(x), L.ptr(w), L.ptr(b), rows, K, N, L.ptr(out), model._stream()))
    return out


def _lin_bwd(model, x, dy, rows, K, N, dw, db):
    with torch.cuda.device(model.device):
        L.check(L.load().clipdlm_small_linear_bwd(L.ptr(x), L.ptr(dy), rows, K, N, L.ptr(dw), L.ptr(db), model._stream()))


def in_proj(model, x16: torch.Tensor) -> torch.Tensor:
    """input_projection (CLIP-DDPM.py:292-293): fp32 [R, ML, ch] -> [R, ML, DIM]."""
    ML, _, ch, D, _ = _dims(model)
    R = x16.shape[0]
    out = model._scratch("te_u", (R, ML, D), torch.float32)
    return _lin_fwd(model, x16, model._views["input_projection.weight"], model._views["input_projection.bias"], R * ML, ch, D, out)


def out_proj(model, xo: torch.Tensor, tag: str = "te_y") -> torch.Tensor:
    """output_projection (CLIP-DDPM.py:319-320): fp32 [R, L, DIM] -> [R, L, ch]."""
    _, Lf, ch, D, _ = _dims(model)
    R = xo.shape[0]
    out = model._scratch(tag, (R, Lf, ch), torch.float32)
    return _lin_fwd(model, xo, model._views["output_projection.weight"], model._views["output_projection.bias"], R * Lf, D, ch, out)


def _pack(model, y: torch.Tensor):
    """y[:, :ML] -> compact bf16 (pair) [R * ML, CH_PAD] GEMM operand."""
    ML, Lf, ch, _, _ = _dims(model)
    M = y.shape[0] * ML
    hi = model._scratch("te_a_hi", (M, CH_PAD), torch.bfloat16)
    lo = model._scratch("te_a_lo", (M, CH_PAD), torch.bfloat16) if model.precision == "bf16x3" else None
    with torch.cuda.device(model.device):
        L.check(L.load().clipdlm_pack_rows_bf16(L.ptr(y), M, ML, Lf, ch, CH_PAD, L.ptr(hi), L.ptr(lo), model._stream()))
    return hi, lo


def _lm_shadow(model):
    off = model._te_off["lm_head.weight"]
    hi = model.shadow_hi.data_ptr() + 2 * off
    lo = model.shadow_lo.data_ptr() + 2 * off if model.shadow_lo is not None else None
    return hi, lo


def _gemm(model, **kw):
    g = L.Gemm()
    for k, v in kw.items():
        setattr(g, k, v.data_ptr() if hasattr(v, "data_ptr") else v)
    with torch.cuda.device(model.device):
        L.check(L.load().clipdlm_gemm(C.byref(g), model._stream()))


def _lse(model, a_hi, a_lo, M, targets, tgt_period, want_argmax, loss_acc_ptr, scale):
    """lm_head + online log-sum-exp (+ running argmax) without logits in HBM. Returns (lse, argmax or None)."""
    V = model.hp["VOCAB_SIZE"]
    nt = (V + 255) // 256
    pm = model._scratch("te_pm", (2 * nt, M), torch.float32)
    ps = model._scratch("te_ps", (2 * nt, M), torch.float32)
    pa = model._scratch("te_pa", (2 * nt, M), torch.int32) if want_argmax else None
    tl = model._scratch("te_tl", (M,), torch.float32)
    lse = model._scratch("te_lse", (M,), torch.float32)
    am = torch.empty(M, device=model.device, dtype=torch.int32) if want_argmax else None
    w_hi, w_lo = _lm_shadow(model)
    _gemm(model, a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, lda=CH_PAD, ldb=CH_PAD, M=M, N=V, K=CH_PAD, epilogue=L.EPI_LSE,
          part_max=pm, part_sum=ps, part_arg=pa, tgt_logit=tl, targets=targets, tgt_period=tgt_period)
    with torch.cuda.device(model.device):
        L.check(L.load().clipdlm_lse_combine(L.ptr(pm), L.ptr(ps), L.ptr(pa), 2 * nt, M, L.ptr(tl) if targets is not None else None,
                                             L.ptr(lse), L.ptr(am), loss_acc_ptr, scale, model._stream()))
    return lse, am


def logits(model, y: torch.Tensor) -> torch.Tensor:
    """self.lm_head(x_out[:, :MAX_LENGTH, :]) (CLIP-DDPM.py:323) for y fp32 [R, L, ch]: fp32 [R, ML, V]."""
    ML, _, _, _, V = _dims(model)
    R = y.shape[0]
    a_hi, a_lo = _pack(model, y)
    n32 = (V + 31) // 32 * 32
    out = torch.empty(R * ML, n32, device=model.device)
    w_hi, w_lo = _lm_shadow(model)
    _gemm(model, a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, lda=CH_PAD, ldb=CH_PAD, M=R * ML, N=n32, K=CH_PAD, epilogue=L.EPI_STORE,
          out_f32=out, ldo=n32)
    return out.view(R, ML, n32)[:, :, :V]


def logits_dense(model, x: torch.Tensor) -> torch.Tensor:
    """model.lm_head(x) for an arbitrary [..., ch] tensor (demos / the CFG branch of forward()); not on the hot path."""
    _, _, ch, _, V = _dims(model)
    lead = x.shape[:-1]
    flat = x.detach().to(model.device, torch.float32).reshape(-1, ch)
    n = flat.shape[0]
    pad = torch.zeros(n, CH_PAD, device=model.device)
    pad[:, :ch] = flat
    hi = pad.to(torch.bfloat16)
    lo = (pad - hi.float()).to(torch.bfloat16) if model.precision == "bf16x3" else None
    n32 = (V + 31) // 32 * 32
    out = torch.empty(n, n32, device=model.device)
    w_hi, w_lo = _lm_shadow(model)
    _gemm(model, a_hi=hi, a_lo=lo, b_hi=w_hi, b_lo=w_lo, lda=CH_PAD, ldb=CH_PAD, M=n, N=n32, K=CH_PAD, epilogue=L.EPI_STORE, out_f32=out, ldo=n32)
    return out[:, :V].reshape(*lead, V)


def argmax(model, y: torch.Tensor) -> torch.Tensor:
    """argmax over the vocabulary of lm_head(y[:, :ML]) via the fused running-argmax epilogue: int64 [R, ML]."""
    ML = model.hp["MAX_LENGTH"]
    R = y.shape[0]
    a_hi, a_lo = _pack(model, y)
    _, am = _lse(model, a_hi, a_lo, R * ML, None, 1, True, None, 0.0)
    return am.view(R, ML).to(torch.int64)


def loss_pass(model, eng, losses: torch.Tensor, slot: int, *, x16, R, B, R_total, target, target_rows, img, txt, mask32, ids32, seed,
              use_embed: bool, backward: bool, d_target: Optional[torch.Tensor] = None, cfg_rows: Optional[torch.Tensor] = None,
              eng_guided=None):
    """One pass of loss() (CLIP-DDPM.py:415-437) over R rows in TRAIN_EMBEDDING mode: forward, the two loss terms into
    losses[slot] / losses[slot + 1] (float64 accumulators), and - if backward - every parameter gradient except the embedding's,
    whose input-side gradient d(x16) [R, ML, ch] is returned (the caller folds it through q_sample, see clipdlm_embedding_bwd).
    d_target accumulates -d(loss)/d(target rows). cfg_rows (int32 [R], 1 = classifier-free-guided row) switches on the guidance mix of
    CLIP-DDPM.py:313-317: a second, guided encoder pass over the same projected input in `eng_guided`, the two fp32 encoder outputs
    mixed row-wise BEFORE the output projection (:319), the gradient split (1 + w) / -w between the passes on the way back."""
    hp = model.hp
    lib = L.load()
    ML, Lf, ch, D, V = _dims(model)
    T16, T = R * ML, R * Lf
    kind = LOSS_KIND[hp["LOSS_FUNC"]]
    x16 = x16.contiguous()
    u = in_proj(model, x16)
    xo = model._scratch("te_xo", (R, Lf, D), torch.float32)
    model._run_forward(eng, R=R, B=B, mode=0, guided=False, train=model.training, image_clip=img, text_clip=txt, attn_mask=mask32, x_in=u,
                       x_out=xo, drop_seed=seed)
    w_cfg = float(hp["CLASSIFIER_FREE_WEIGHT"])
    if cfg_rows is not None:
        xg = model._scratch("te_xo_g", (R, Lf, D), torch.float32)
        model._run_forward(eng_guided, R=R, B=B, mode=0, guided=True, train=model.training, image_clip=img, text_clip=txt, attn_mask=mask32,
                           x_in=u, x_out=xg, drop_seed=seed + 104729)   # the reference's second self.model(...) call draws its own dropout
        with torch.cuda.device(model.device):
            L.check(lib.clipdlm_row_mix_f32(L.ptr(xo), L.ptr(xg), L.ptr(cfg_rows), w_cfg, R, Lf * D, model._stream()))
    y = out_proj(model, xo)
    ce_scale = 1.0 / R_total if kind in (0, 2) else 1.0 / hp["BATCH_SIZE"]  # CLIP-DDPM.py:437 vs :439-440
    dce = None
    st = model._stream()
    if hp["USE_PROB_LOSS"]:
        a_hi, a_lo = _pack(model, y)
        lse, _ = _lse(model, a_hi, a_lo, T16, ids32, B * ML, False, losses.data_ptr() + 8 * (slot + 1), ce_scale)
        if backward:
            ldl = (V + 255) // 256 * 256
            d_hi = model._scratch("te_dlog_hi", (T16, ldl), torch.bfloat16)
            d_lo = model._scratch("te_dlog_lo", (T16, ldl), torch.bfloat16) if model.precision == "bf16x3" else None
            w_hi, w_lo = _lm_shadow(model)
            rw_host, rw_dev = model.rounding_weight()   # (a device-resident dynamic weight multiplies grad_scale inside the epilogue)
            _gemm(model, a_hi=a_hi, a_lo=a_lo, b_hi=w_hi, b_lo=w_lo, lda=CH_PAD, ldb=CH_PAD, M=T16, N=V, K=CH_PAD, epilogue=L.EPI_SMGRAD,
                  out_hi=d_hi, out_lo=d_lo, ldo=ldl, lse=lse, targets=ids32, tgt_period=B * ML, grad_scale=rw_host * ce_scale,
                  **({"part_max": rw_dev} if rw_dev is not None else {}))
            dce = model._scratch("te_dce", (T16, CH_PAD), torch.float32)
            # d y[:, :ML] = dlogits [T16, V] W [V, ch]   (W read in place as an MN-major operand)
            _gemm(model, a_hi=d_hi, a_lo=d_lo, b_hi=w_hi, b_lo=w_lo, lda=ldl, ldb=CH_PAD, M=T16, N=CH_PAD, K=V, a_major=0, b_major=1,
                  epilogue=L.EPI_STORE, out_f32=dce, ldo=CH_PAD)
            # d W [V, ch] += dlogits^T y[:, :ML]   (straight into the padded gradient slot of lm_head.weight)
            g_lm = model.grad.data_ptr() + 4 * model._te_off["lm_head.weight"]
            _gemm(model, a_hi=d_hi, a_lo=d_lo, b_hi=a_hi, b_lo=a_lo, lda=ldl, ldb=CH_PAD, M=V, N=CH_PAD, K=T16, a_major=1, b_major=1,
                  epilogue=L.EPI_WGRAD, acc_f32=g_lm, ldo=CH_PAD)
    dy = model._scratch("te_dy", (R, Lf, ch), torch.float32) if backward else None
    if use_embed or backward:
        with torch.cuda.device(model.device):
            L.check(lib.clipdlm_feature_loss_f32(L.ptr(y), L.ptr(target), target_rows, R, ML, Lf, ch, kind, R_total, hp["BATCH_SIZE"],
                                                 1.0 if use_embed else 0.0, losses.data_ptr() + 8 * slot if use_embed else None,
                                                 L.ptr(dce), CH_PAD, L.ptr(dy), L.ptr(d_target) if (backward and use_embed) else None, st))
    if not backward:
        return None
    gv = model._gviews
    _lin_bwd(model, xo, dy, T, D, ch, gv["output_projection.weight"], gv["output_projection.bias"])
    w_out_t = model._views["output_projection.weight"].t().contiguous()  # [D, ch]
    dxo = _lin_fwd(model, dy, w_out_t, None, T, ch, D, model._scratch("te_dxo", (R, Lf, D), torch.float32))
    du = model._scratch("te_du", (R, ML, D), torch.float32)
    with torch.cuda.device(model.device):
        if cfg_rows is None:
            L.check(lib.clipdlm_engine_backward_from(eng, L.ptr(dxo), L.ptr(du), st))
        else:
            dxg = model._scratch("te_dxo_g", (R, Lf, D), torch.float32)
            du_g = model._scratch("te_du_g", (R, ML, D), torch.float32)
            L.check(lib.clipdlm_row_split_f32(L.ptr(dxo), L.ptr(dxg), L.ptr(cfg_rows), w_cfg, R, Lf * D, st))
            L.check(lib.clipdlm_engine_backward_from(eng, L.ptr(dxo), L.ptr(du), st))
            L.check(lib.clipdlm_engine_backward_from(eng_guided, L.ptr(dxg), L.ptr(du_g), st))
            L.check(lib.clipdlm_add_f32(L.ptr(du), L.ptr(du_g), R * ML * D, st))   # both passes consumed the same projected input
    _lin_bwd(model, x16, du, T16, ch, D, gv["input_projection.weight"], gv["input_projection.bias"])
    w_in_t = model._views["input_projection.weight"].t().contiguous()  # [ch, D]
    dx16 = _lin_fwd(model, du, w_in_t, None, T16, D, ch, model._scratch("te_dx16", (R, ML, ch), torch.float32))
    model._grads_dirty = True
    return dx16


def embedding_bwd(model, dx: torch.Tensor, scale: Optional[torch.Tensor], ids32: torch.Tensor, S: int):
    """d embedding.weight[ids] += sum_s scale[s] * dx[s]  (dx fp32 [S * B, ML, ch])."""
    ch = model.hp["IN_CHANNEL"]
    with torch.cuda.device(model.device):
        L.check(L.load().clipdlm_embedding_bwd(L.ptr(dx), L.ptr(scale), L.ptr(ids32), S, ids32.numel(), ch,
                                               L.ptr(model._gviews["embedding.weight"]), model._stream()))
