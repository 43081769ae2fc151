"""ctypes binding of libclipdlm.so (include/clipdlm.h). This is the only place the product touches native code.

There is deliberately NO fallback: if the shared object is missing or a symbol cannot be resolved, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libclipdlm.so")

c_p = C.c_void_p
i32, i64, u32, u64, f32, f64 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float, C.c_double


class Bf(C.Structure):
    _fields_ = [("hi", c_p), ("lo", c_p)]


class Gemm(C.Structure):
    _fields_ = [
        ("a_hi", c_p), ("a_lo", c_p), ("b_hi", c_p), ("b_lo", c_p),
        ("lda", i64), ("ldb", i64),
        ("M", i32), ("N", i32), ("K", i32),
        ("a_major", i32), ("b_major", i32),
        ("gather_len", i32), ("gather_stride", i32),
        ("epilogue", i32), ("k_splits", i32),
        ("out_hi", c_p), ("out_lo", c_p), ("out_f32", c_p), ("ldo", i64),
        ("out2_hi", c_p), ("out2_lo", c_p),
        ("bias", c_p),
        ("res_hi", c_p), ("res_lo", c_p), ("ldr", i64),
        ("u_hi", c_p), ("u_lo", c_p), ("ldu", i64),
        ("scatter_len", i32), ("scatter_stride", i32),
        ("drop_seed", u64), ("drop_site", u32), ("drop_p", f32),
        ("acc_f32", c_p),
        ("part_max", c_p), ("part_sum", c_p), ("part_arg", c_p),
        ("tgt_logit", c_p),
        ("targets", c_p), ("tgt_period", i32),
        ("lse", c_p),
        ("grad_scale", f32),
        ("exp_shift", c_p), ("row_scale", c_p),
    ]


class Embed(C.Structure):
    _fields_ = [
        ("R", i32), ("B", i32), ("Ltxt", i32), ("L", i32), ("D", i32), ("fusion", i32), ("mode", i32), ("guided", i32),
        ("x_in", c_p), ("x_in_stride", i64),
        ("emb_table", c_p), ("ids", c_p), ("noise", c_p), ("coef_a", c_p), ("coef_b", c_p),
        ("img_proj", c_p), ("txt_proj", c_p),
        ("seg", c_p), ("pos", c_p),
        ("ln_w", c_p), ("ln_b", c_p), ("ln_eps", f32),
        ("z", Bf), ("h", Bf),
        ("drop_seed", u64), ("drop_site", u32), ("drop_p", f32),
    ]


class Config(C.Structure):
    _fields_ = [
        ("n_layers", i32), ("dim", i32), ("n_heads", i32), ("hidden_dim", i32), ("vocab", i32), ("max_len", i32),
        ("clip_dim", i32), ("max_pos", i32),
        ("fusion", i32), ("precision", i32),
        ("ln_eps", f32), ("dropout", f32), ("attn_dropout", f32),
    ]


class Buffers(C.Structure):
    _fields_ = [
        ("params", c_p), ("grads", c_p), ("shadow_hi", c_p), ("shadow_lo", c_p),
        ("emb_table", c_p), ("emb_hi", c_p), ("emb_lo", c_p),
        ("workspace", c_p), ("workspace_bytes", C.c_size_t),
    ]


class Pass(C.Structure):
    _fields_ = [
        ("R", i32), ("B", i32), ("mode", i32), ("guided", i32), ("train", i32), ("reuse_proj", i32),
        ("x_in", c_p), ("x_in_stride", i64),
        ("ids", c_p), ("noise", c_p), ("coef_a", c_p), ("coef_b", c_p),
        ("image_clip", c_p), ("text_clip", c_p), ("attn_mask", c_p),
        ("drop_seed", u64),
        ("x_out", c_p),
    ]


class LossCfg(C.Structure):
    _fields_ = [
        ("loss_kind", i32), ("use_embed_loss", i32), ("use_prob_loss", i32), ("batch_size", i32),
        ("R_total", i64),
        ("rounding_weight", f32), ("backward", i32),
        ("target", c_p), ("target_rows", i32),
        ("row_scale_self", c_p), ("row_scale_export", c_p), ("export_engine", c_p),
        ("rounding_weight_dev", c_p),
    ]


MAX_PEERS = 16


class DpBuffers(C.Structure):
    _fields_ = [("rank", i32), ("world", i32), ("p", c_p * MAX_PEERS), ("g", c_p * MAX_PEERS), ("shadow_hi", c_p * MAX_PEERS),
                ("shadow_lo", c_p * MAX_PEERS), ("p_mc", c_p), ("g_mc", c_p), ("shadow_hi_mc", c_p), ("shadow_lo_mc", c_p)]


EPI_STORE, EPI_WGRAD, EPI_LSE, EPI_SMGRAD, EPI_LSE_EXP, EPI_STORE_ROWSCALE, EPI_STORE_GELU_DERIV, EPI_STORE_MULAUX = 0, 1, 2, 3, 4, 5, 6, 7
OPT_FUSED_SOFTMAX_GRAD, OPT_EXP_SHIFT_PTR, OPT_GELU_DERIV_STORE = 1, 2, 3

PROF_CATEGORIES = ("gemm_fwd", "gemm_dgrad", "gemm_wgrad", "gemm_lse", "gemm_smgrad", "attn_fwd", "attn_bwd", "ln_fwd", "ln_bwd", "embed",
                   "loss", "colsum", "other")


class Prof(C.Structure):
    _fields_ = [("ms", f64), ("flops", f64), ("bytes", f64), ("launches", i64)]

# parameter slots (enum in clipdlm.h)
(P_POS, P_EMB_LN_W, P_EMB_LN_B, P_VT_W, P_VT_B, P_VLN_W, P_VLN_B, P_IMG_W, P_IMG_B, P_TXT_W, P_TXT_B, P_SEG,
 P_LAYER0) = range(13)
(PL_QKV_W, PL_QKV_B, PL_O_W, PL_O_B, PL_LN1_W, PL_LN1_B, PL_FF1_W, PL_FF1_B, PL_FF2_W, PL_FF2_B, PL_LN2_W, PL_LN2_B,
 P_PER_LAYER) = range(13)

_SIGS = {
    "clipdlm_last_error": (C.c_char_p, []),
    "clipdlm_version": (C.c_int, []),
    "clipdlm_device_ok": (C.c_int, []),
    "clipdlm_gemm": (C.c_int, [C.POINTER(Gemm), c_p]),
    "clipdlm_gemm_debug_mn_desc": (None, [u32, u32]),
    "clipdlm_gemm_debug_flags": (None, [u32]),
    "clipdlm_lse_combine": (C.c_int, [c_p, c_p, c_p, i32, i32, c_p, c_p, c_p, c_p, f64, c_p]),
    "clipdlm_ce_row_terms": (C.c_int, [c_p, c_p, c_p, i32, f32, i32, c_p, i64, c_p, i64, i32, i32, i32, c_p, c_p]),
    "clipdlm_embed_fwd": (C.c_int, [C.POINTER(Embed), c_p]),
    "clipdlm_embed_bwd": (C.c_int, [C.POINTER(Bf), i32, i32, i32, i32, i32, i32, i32, c_p, c_p, c_p, c_p, c_p]),
    "clipdlm_layernorm_fwd": (C.c_int, [C.POINTER(Bf), c_p, c_p, f32, i64, i32, C.POINTER(Bf), c_p, u64, u32, f32, c_p]),
    "clipdlm_layernorm_bwd": (C.c_int, [C.POINTER(Bf), C.POINTER(Bf), c_p, f32, i64, i32, C.POINTER(Bf), c_p, c_p,
                                        u64, u32, f32, C.POINTER(Bf), u32, f32, C.POINTER(Bf), c_p, c_p]),
    "clipdlm_attn_fwd": (C.c_int, [C.POINTER(Bf), c_p, i32, i32, i32, i32, C.POINTER(Bf), u64, u32, f32, c_p]),
    "clipdlm_attn_bwd": (C.c_int, [C.POINTER(Bf), c_p, C.POINTER(Bf), i32, i32, i32, i32, C.POINTER(Bf), u64, u32, f32, c_p]),
    "clipdlm_attn_bwd_bias": (C.c_int, [C.POINTER(Bf), c_p, C.POINTER(Bf), i32, i32, i32, i32, C.POINTER(Bf), u64, u32, f32, c_p, c_p, c_p]),
    "clipdlm_attn_force_simt": (None, [i32]),
    "clipdlm_colsum": (C.c_int, [C.POINTER(Bf), i64, i32, c_p, c_p]),
    "clipdlm_embed_loss": (C.c_int, [C.POINTER(Bf), c_p, c_p, c_p, i32, i32, i32, i32, i32, i32, i32, i64, i32, f32, c_p,
                                     C.POINTER(Bf), c_p]),
    "clipdlm_small_linear_fwd": (C.c_int, [c_p, c_p, c_p, i32, i32, i32, c_p, c_p]),
    "clipdlm_small_linear_bwd": (C.c_int, [c_p, c_p, i32, i32, i32, c_p, c_p, c_p]),
    "clipdlm_adamw": (C.c_int, [c_p, c_p, c_p, c_p, c_p, c_p, i64, f32, f32, f32, f32, f32, i32, f32, i32, c_p]),
    "clipdlm_dp_slice": (C.c_int, [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]),
    "clipdlm_adamw_dp": (C.c_int, [C.POINTER(DpBuffers), c_p, c_p, i64, f32, f32, f32, f32, f32, i32, f32, c_p]),
    "clipdlm_to_bf16": (C.c_int, [c_p, c_p, c_p, i64, c_p]),
    "clipdlm_to_f32": (C.c_int, [c_p, c_p, c_p, i64, c_p]),
    "clipdlm_gather_rows_f32": (C.c_int, [C.POINTER(Bf), i64, i32, i32, i32, c_p, c_p]),
    "clipdlm_q_sample": (C.c_int, [c_p, c_p, c_p, c_p, i64, i32, c_p, c_p]),
    "clipdlm_keymask": (C.c_int, [c_p, i32, i32, i32, i32, i32, i32, c_p, c_p]),
    "clipdlm_param_count": (i64, [C.POINTER(Config)]),
    "clipdlm_param_offset": (i64, [C.POINTER(Config), i32]),
    "clipdlm_param_size": (i64, [C.POINTER(Config), i32]),
    "clipdlm_workspace_bytes": (C.c_size_t, [C.POINTER(Config), i32, i32, i32]),
    "clipdlm_engine_create": (c_p, [C.POINTER(Config), C.POINTER(Buffers), i32, i32, i32]),
    "clipdlm_engine_destroy": (None, [c_p]),
    "clipdlm_engine_forward": (C.c_int, [c_p, C.POINTER(Pass), c_p]),
    "clipdlm_engine_lm_head": (C.c_int, [c_p, c_p, i64, c_p, c_p]),
    "clipdlm_engine_loss_backward": (C.c_int, [c_p, C.POINTER(LossCfg), c_p, c_p]),
    "clipdlm_engine_cfg_mix": (C.c_int, [c_p, c_p, c_p, f32, c_p]),
    "clipdlm_engine_backward": (C.c_int, [c_p, c_p]),
    "clipdlm_engine_backward_from": (C.c_int, [c_p, c_p, c_p, c_p]),
    "clipdlm_feature_loss_f32": (C.c_int, [c_p, c_p, i32, i32, i32, i32, i32, i32, i64, i32, f32, c_p, c_p, i32, c_p, c_p, c_p]),
    "clipdlm_pack_rows_bf16": (C.c_int, [c_p, i64, i32, i32, i32, i32, c_p, c_p, c_p]),
    "clipdlm_embedding_bwd": (C.c_int, [c_p, c_p, c_p, i32, i64, i32, c_p, c_p]),
    "clipdlm_row_mix_f32": (C.c_int, [c_p, c_p, c_p, f32, i32, i64, c_p]),
    "clipdlm_row_split_f32": (C.c_int, [c_p, c_p, c_p, f32, i32, i64, c_p]),
    "clipdlm_add_f32": (C.c_int, [c_p, c_p, i64, c_p]),
    "clipdlm_engine_set_option": (C.c_int, [c_p, i32, i64]),
    "clipdlm_engine_launch_count": (i64, [c_p]),
    "clipdlm_engine_profile": (C.c_int, [c_p, i32]),
    "clipdlm_engine_profile_read": (C.c_int, [c_p, c_p, i32]),
}

EXPORTED_SYMBOLS = tuple(_SIGS.keys())

_lib = None


class ClipdlmError(RuntimeError):
    pass


def load():
    """Load libclipdlm.so (building is the caller's / __graft_entry__.build()'s job). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ClipdlmError(
            f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise ClipdlmError(f"libclipdlm error {rc}: {load().clipdlm_last_error().decode()}")


def bf(hi, lo=None) -> Bf:
    """Bf pair from torch tensors (lo may be None)."""
    return Bf(hi.data_ptr() if hi is not None else None, lo.data_ptr() if lo is not None else None)


def ptr(t):
    return None if t is None else t.data_ptr()
