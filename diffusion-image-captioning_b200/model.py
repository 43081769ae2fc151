"""Host-side mirror of the reference's model / optimizer objects (CLIP-DDPM.py:227-335) over the native engine.

`DistilBertModel` keeps the reference's constructor, `forward(x, image_clip, text_clip, mask, concat_mask)` signature,
`.embedding`, `.lm_head`, `.parameters()`, `.train()/.eval()`; underneath it owns ONE flat fp32 parameter buffer (+ flat grad
buffer + bf16 shadow) laid out by libclipdlm (include/clipdlm.h, clipdlm_param_offset) and calls the C-ABI engine.
`AdamW` mirrors `torch.optim.AdamW(model.parameters(), lr)` (param_groups[0]['lr'], zero_grad(), step()) over the flat buffer.

PyTorch is used for device memory, streams and torch.distributed only. There is no CPU / eager fallback: every entry point
raises if libclipdlm.so or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L
from .hparams import default_hparams


class _Cfg:
    """Duck-typed stand-in for transformers.DistilBertConfig (only the fields the path reads)."""

    def __init__(self, n_layers=6, dim=768, n_heads=12, hidden_dim=3072, dropout=0.1, attention_dropout=0.1,
                 max_position_embeddings=512, vocab_size=30522):
        self.n_layers, self.dim, self.n_heads, self.hidden_dim = n_layers, dim, n_heads, hidden_dim
        self.dropout, self.attention_dropout = dropout, attention_dropout
        self.max_position_embeddings, self.vocab_size = max_position_embeddings, vocab_size


DistilBertConfig = _Cfg


class ParamList(list):
    """`model.parameters()` result: the reference returns a plain list (CLIP-DDPM.py:258-269); this one remembers its owner
    so that `AdamW(model.parameters(), lr)` can find the flat buffers."""
    owner: "DistilBertModel" = None


class _Embedding:
    """`model.embedding(ids)` (CLIP-DDPM.py:459,584): frozen lookup into the fp32 table."""

    def __init__(self, weight: torch.Tensor):
        self.weight = weight

    def __call__(self, ids: torch.Tensor) -> torch.Tensor:
        return self.weight[ids]


class _LmHead:
    """`model.lm_head` (CLIP-DDPM.py:246-247): frozen projection, bias zeroed. Calling it runs the tcgen05 GEMM."""

    def __init__(self, owner: "DistilBertModel"):
        self._owner = owner
        self.weight = owner.lm_head_weight
        # TRAIN_EMBEDDING: nn.Linear(IN_CHANNEL, VOCAB_SIZE, bias=False), trainable (CLIP-DDPM.py:239)
        self.bias = None if owner.hp["TRAIN_EMBEDDING"] else torch.zeros(owner.hp["VOCAB_SIZE"], device=owner.device)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self._owner.hp["TRAIN_EMBEDDING"]:
            from . import train_embedding as TE
            return TE.logits_dense(self._owner, x)
        return self._owner._lm_head_dense(x)


def _slot_names(hp: dict) -> List[Tuple[str, int, Tuple[int, ...], int]]:
    """(reference parameter name, slot, shape, row offset inside the slot) in the reference's parameters() order."""
    d, f, c = hp["DIM"], hp["HIDDEN_DIM"], hp["CLIP_DIM"]
    out = [("model.distilbert.embeddings.position_embeddings.weight", L.P_POS, (hp["MAX_POSITION"], d), 0),
           ("model.distilbert.embeddings.LayerNorm.weight", L.P_EMB_LN_W, (d,), 0),
           ("model.distilbert.embeddings.LayerNorm.bias", L.P_EMB_LN_B, (d,), 0)]
    for i in range(hp["N_LAYERS"]):
        base = L.P_LAYER0 + i * L.P_PER_LAYER
        p = f"model.distilbert.transformer.layer.{i}."
        for k, lin in enumerate(("q_lin", "k_lin", "v_lin")):
            out += [(p + f"attention.{lin}.weight", base + L.PL_QKV_W, (d, d), k * d * d),
                    (p + f"attention.{lin}.bias", base + L.PL_QKV_B, (d,), k * d)]
        out += [(p + "attention.out_lin.weight", base + L.PL_O_W, (d, d), 0), (p + "attention.out_lin.bias", base + L.PL_O_B, (d,), 0),
                (p + "sa_layer_norm.weight", base + L.PL_LN1_W, (d,), 0), (p + "sa_layer_norm.bias", base + L.PL_LN1_B, (d,), 0),
                (p + "ffn.lin1.weight", base + L.PL_FF1_W, (f, d), 0), (p + "ffn.lin1.bias", base + L.PL_FF1_B, (f,), 0),
                (p + "ffn.lin2.weight", base + L.PL_FF2_W, (d, f), 0), (p + "ffn.lin2.bias", base + L.PL_FF2_B, (d,), 0),
                (p + "output_layer_norm.weight", base + L.PL_LN2_W, (d,), 0), (p + "output_layer_norm.bias", base + L.PL_LN2_B, (d,), 0)]
    out += [("model.vocab_transform.weight", L.P_VT_W, (d, d), 0), ("model.vocab_transform.bias", L.P_VT_B, (d,), 0),
            ("model.vocab_layer_norm.weight", L.P_VLN_W, (d,), 0), ("model.vocab_layer_norm.bias", L.P_VLN_B, (d,), 0),
            ("image_linear.weight", L.P_IMG_W, (d, c), 0), ("image_linear.bias", L.P_IMG_B, (d,), 0),
            ("text_linear.weight", L.P_TXT_W, (d, c), 0), ("text_linear.bias", L.P_TXT_B, (d,), 0)]
    if hp["CLIP_ADDING_METHOD"] == "concat":
        out += [("segment_embedding.weight", L.P_SEG, (2, d), 0)]
    return out


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class DistilBertModel(torch.nn.Module):
    """Drop-in for the reference's `class DistilBertModel(nn.Module)` (CLIP-DDPM.py:227-323). An `nn.Module` like the reference's: `model(x, ...)`,
    `.train()/.eval()`, `.parameters()` (the reference overrides it to return a list, :258-269; so does this one - `nn.Parameter`s that alias the
    flat fp32 buffer, `.grad` aliasing the flat gradient buffer), `torch.save(model.cpu(), path)` / `torch.load(path).to(device)` (:551,560,570:
    the pickle holds the host state dict + hyperparameters; the device buffers are rebuilt on load).

    embedding / projection: the pretrained word-embedding and vocab-projector (modules with `.weight`, or tensors
    [VOCAB_SIZE, DIM]); frozen copies are taken (:245-247). None => random N(0, 0.02) tied table (what
    `DistilBertForMaskedLM(DistilBertConfig())` gives, the stand-in used where the checkpoint is unavailable).
    config: DistilBertConfig-like object (n_layers, dim, n_heads, hidden_dim, dropout, attention_dropout, ...).
    hp: hyperparameter dict (hparams.default_hparams). precision: "bf16" (speed mode: bf16 tensor-core passes) or "bf16x3"
    (parity mode: split-bf16 storage, three tensor-core passes per GEMM, fp32-class results).
    """

    def __init__(self, embedding=None, projection=None, config=None, hp: Optional[dict] = None, precision: str = "bf16",
                 device="cuda", seed: Optional[int] = None, chunk_rows: int = 8192, fused_softmax_grad: Optional[bool] = None,
                 gelu_deriv_store: Optional[bool] = None):
        super().__init__()
        lib = L.load()
        if not torch.cuda.is_available():
            raise L.ClipdlmError("clipdlm needs a CUDA device (sm_100a); there is no CPU fallback")
        self._ctor = dict(precision=precision, chunk_rows=int(chunk_rows), fused_softmax_grad=fused_softmax_grad, gelu_deriv_store=gelu_deriv_store)
        self.device = torch.device(device if device != "cuda" else f"cuda:{torch.cuda.current_device()}")
        with torch.cuda.device(self.device):
            if lib.clipdlm_device_ok() != 1:
                raise L.ClipdlmError("clipdlm needs a compute-capability 10.x device (B200, tcgen05/TMEM)")
        hp = dict(hp) if hp is not None else default_hparams()
        if config is not None:
            hp.update(N_LAYERS=config.n_layers, DIM=config.dim, N_HEADS=config.n_heads, HIDDEN_DIM=config.hidden_dim,
                      DROPOUT=float(config.dropout), ATTENTION_DROPOUT=float(config.attention_dropout),
                      MAX_POSITION=config.max_position_embeddings)
            if not hp["TRAIN_EMBEDDING"]:
                hp["IN_CHANNEL"] = config.dim
        if hp["TRAIN_EMBEDDING"] and hp["IN_CHANNEL"] == hp["DIM"]:
            hp["IN_CHANNEL"] = 16  # CLIP-DDPM.py:99-100
        if hp["CLIP_ADDING_METHOD"] not in ("concat", "add"):
            raise NotImplementedError(hp["CLIP_ADDING_METHOD"])  # CLIP-DDPM.py:269
        if precision not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3'")
        self.hp = hp
        self.precision = precision
        self.training = True
        self.chunk_rows = int(chunk_rows)
        # Factored softmax-CE gradient of the lm_head in precision="bf16" (clipdlm.h CLIPDLM_OPT_FUSED_SOFTMAX_GRAD): the lm_head pass stores
        # exp(s - c) instead of the logits and the gradient GEMM applies 1 / sum as a row factor - no in-place pass over the stored
        # logits. Default ON since round 2 (validated on B200: tests/test_fused_paths_gpu.py, -13 ms per step); None = the
        # CLIPDLM_FUSED_SOFTMAX_GRAD environment switch (0 selects the in-place path, which also is the automatic fallback when the
        # logit bound of refresh_exp_shift() leaves the range the stored exponentials cover).
        if fused_softmax_grad is None:
            fused_softmax_grad = os.environ.get("CLIPDLM_FUSED_SOFTMAX_GRAD", "1") == "1"
        self.fused_softmax_grad = bool(fused_softmax_grad) and precision == "bf16" and not hp["TRAIN_EMBEDDING"]
        self._exp_shift = None
        self._exp_bound_host = None     # pinned copy of the logit bound, checked one step late (no host sync on the step path)
        self._exp_bound_event = None
        # lin1 stores gelu'(u) instead of u and the lin2 gradient GEMM multiplies by it (clipdlm.h CLIPDLM_OPT_GELU_DERIV_STORE);
        # 2 = additionally the lin1 bias gradient is summed in that GEMM's epilogue (no colsum pass over the [tokens, 3072] gradient).
        # Default 2 since round 2 (-9 ms per step); None = the CLIPDLM_GELU_DERIV_STORE environment switch (0 / 1 / 2).
        if gelu_deriv_store is None:
            gelu_deriv_store = int(os.environ.get("CLIPDLM_GELU_DERIV_STORE", "2") or 0)
        self.gelu_deriv_store = int(gelu_deriv_store) if precision == "bf16" else 0
        if self.gelu_deriv_store not in (0, 1, 2):
            raise ValueError("gelu_deriv_store must be 0 / False, 1 / True or 2")
        self.dp_group = None  # set by parallel.enable_data_parallel
        self.dp_world = 1
        self._cfg = L.Config(hp["N_LAYERS"], hp["DIM"], hp["N_HEADS"], hp["HIDDEN_DIM"], hp["VOCAB_SIZE"], hp["MAX_LENGTH"], hp["CLIP_DIM"],
                             hp["MAX_POSITION"], 0 if hp["CLIP_ADDING_METHOD"] == "concat" else 1, 1 if precision == "bf16x3" else 0,
                             1e-12, hp["DROPOUT"], hp["ATTENTION_DROPOUT"])
        n = lib.clipdlm_param_count(C.byref(self._cfg))
        if n <= 0:
            raise L.ClipdlmError("bad model configuration: " + lib.clipdlm_last_error().decode())
        self.n_engine = int(n)  # the encoder's slots (laid out by libclipdlm)
        # TRAIN_EMBEDDING: the extra trainable tensors follow in the same flat buffer, so the one fused AdamW / all-reduce covers them
        self._te_off: Dict[str, int] = {}
        te_extra = []
        off = _round_up(self.n_engine, 64)
        if hp["TRAIN_EMBEDDING"]:
            from . import train_embedding as TE
            te_extra = TE.extra_params(hp)
            for name, stored, _ in te_extra:
                self._te_off[name] = off
                off += _round_up(int(math.prod(stored)), 64)
        self.n_params = off if te_extra else self.n_engine
        dev = self.device
        self.flat = torch.zeros(self.n_params, device=dev)
        self.grad = torch.zeros(self.n_params, device=dev)
        self.shadow_hi = torch.zeros(self.n_params, device=dev, dtype=torch.bfloat16)
        self.shadow_lo = torch.zeros(self.n_params, device=dev, dtype=torch.bfloat16) if precision == "bf16x3" else None
        self._scr: Dict[str, torch.Tensor] = {}
        self._te_extra = te_extra
        self.dp_fused = None   # set by parallel.enable_data_parallel when the fused NVLink optimizer step is active
        self._build_views()
        self._init_parameters(seed)
        self._engines: Dict[tuple, tuple] = {}
        self._launches_retired = 0
        self._grads_dirty = False
        V, D = hp["VOCAB_SIZE"], hp["DIM"]
        if hp["TRAIN_EMBEDDING"]:  # CLIP-DDPM.py:238-239: both live in the trainable flat buffer; `embedding` / `projection` are ignored
            self.embedding_weight = self._views["embedding.weight"]
            self.lm_head_weight = self._views["lm_head.weight"]
            self._vpad = (V + 255) // 256 * 256
            self.emb_hi = self.emb_lo = None
            self.embedding = _Embedding(self.embedding_weight)
            self.lm_head = _LmHead(self)
            self.sync_shadow()
            return
        if embedding is None:
            g = torch.Generator(device="cpu")
            g.manual_seed(0 if seed is None else seed + 1)
            emb = torch.randn(V, D, generator=g) * 0.02
        else:
            emb = getattr(embedding, "weight", embedding).detach().float()
        if tuple(emb.shape) != (V, D):
            raise ValueError(f"embedding must be [{V}, {D}], got {tuple(emb.shape)}")
        self.embedding_weight = emb.to(dev).contiguous().clone()
        proj = emb if projection is None else getattr(projection, "weight", projection).detach().float()
        if tuple(proj.shape) != (V, D):
            raise ValueError(f"projection must be [{V}, {D}], got {tuple(proj.shape)}")
        self.lm_head_weight = proj.to(dev).contiguous().clone()
        self._vpad = (V + 255) // 256 * 256  # zero rows so that TMA boxes of the vocab tiles never leave the allocation
        self.emb_hi = torch.zeros(self._vpad, D, device=dev, dtype=torch.bfloat16)
        self.emb_lo = torch.zeros(self._vpad, D, device=dev, dtype=torch.bfloat16) if precision == "bf16x3" else None
        self.embedding = _Embedding(self.embedding_weight)
        self.lm_head = _LmHead(self)
        self.sync_shadow()

    # ---------------------------------------------------------------------------------------------------------- params
    def _build_views(self):
        """Reference-named views into the flat parameter / gradient buffers."""
        lib, hp = L.load(), self.hp
        self._views: Dict[str, torch.Tensor] = {}
        self._gviews: Dict[str, torch.Tensor] = {}
        for name, slot, shape, rel in _slot_names(hp):
            off = int(lib.clipdlm_param_offset(C.byref(self._cfg), slot)) + rel
            cnt = int(math.prod(shape))
            self._views[name] = self.flat[off:off + cnt].view(shape)
            self._gviews[name] = self.grad[off:off + cnt].view(shape)
        if sum(v.numel() for v in self._views.values()) != self.n_engine:
            raise L.ClipdlmError("parameter name map does not cover the flat buffer")
        if self._te_extra:  # the reference lists them before segment_embedding (CLIP-DDPM.py:259-265)
            seg = [(k, self._views.pop(k), self._gviews.pop(k)) for k in ("segment_embedding.weight",) if k in self._views]
            for name, stored, logical in self._te_extra:
                o, cnt = self._te_off[name], int(math.prod(stored))
                sl = tuple(slice(0, n) for n in logical)
                self._views[name] = self.flat[o:o + cnt].view(stored)[sl]    # lm_head.weight: strided view into its zero-padded slot
                self._gviews[name] = self.grad[o:o + cnt].view(stored)[sl]
            for k, v, g in seg:
                self._views[k], self._gviews[k] = v, g
        # what parameters() / named_parameters() hand out: nn.Parameters aliasing the views, .grad aliasing the gradient views
        self._params: Dict[str, torch.nn.Parameter] = {}
        for k, v in self._views.items():
            prm = torch.nn.Parameter(v, requires_grad=False)   # no autograd on this path: the backward is hand-written
            prm.grad = self._gviews[k]
            self._params[k] = prm

    def _rebind_buffers(self, flat, grad, shadow_hi, shadow_lo):
        """Move the flat parameter / gradient / shadow buffers into caller-provided storage of the same size (symmetric memory for
        the fused data-parallel step). Values are copied, the named views re-created, engines (which hold raw pointers) dropped."""
        lib = L.load()
        for e in self._engines.values():
            self._launches_retired += int(lib.clipdlm_engine_launch_count(e[0]))
            lib.clipdlm_engine_destroy(e[0])
        self._engines = {}
        self.__dict__.pop("_sample_graphs", None)
        flat.copy_(self.flat); grad.copy_(self.grad); shadow_hi.copy_(self.shadow_hi)
        if self.shadow_lo is not None:
            shadow_lo.copy_(self.shadow_lo)
        self.flat, self.grad, self.shadow_hi = flat, grad, shadow_hi
        self.shadow_lo = shadow_lo if self.shadow_lo is not None else None
        self._build_views()
        if self.hp["TRAIN_EMBEDDING"]:
            self.embedding_weight = self._views["embedding.weight"]
            self.lm_head_weight = self._views["lm_head.weight"]
            self.embedding = _Embedding(self.embedding_weight)
            self.lm_head = _LmHead(self)

    def _init_parameters(self, seed: Optional[int]):
        """HF DistilBERT init (N(0, 0.02) Linear / position weights, LayerNorm (1, 0), zero biases), nn.Linear default for the
        CLIP projections, N(0, 1) segment embedding (CLIP-DDPM.py:236,252-256)."""
        g = torch.Generator(device="cpu")
        g.manual_seed(torch.initial_seed() if seed is None else seed)
        c = self.hp["CLIP_DIM"]
        for name, v in self._views.items():
            if self.hp["TRAIN_EMBEDDING"] and name.startswith(("embedding.", "lm_head.", "input_projection.", "output_projection.")):
                if name == "embedding.weight":  # nn.Embedding default N(0, 1)
                    v.copy_(torch.randn(v.shape, generator=g))
                else:  # nn.Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
                    fan_in = self.hp["DIM"] if name.startswith("output_projection") else self.hp["IN_CHANNEL"]
                    v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) / math.sqrt(fan_in))
            elif "LayerNorm.weight" in name or "layer_norm.weight" in name:
                v.fill_(1.0)
            elif name.startswith(("image_linear", "text_linear")):
                bound = 1.0 / math.sqrt(c)
                v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) * bound)
            elif name.endswith(".bias"):
                v.zero_()
            elif name == "segment_embedding.weight":
                v.copy_(torch.randn(v.shape, generator=g))
            else:
                v.copy_(torch.randn(v.shape, generator=g) * 0.02)

    def parameters(self, recurse: bool = True) -> ParamList:
        out = ParamList(self._params.values())
        out.owner = self
        return out

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        return [(prefix + k, v) for k, v in self._params.items()]

    def zero_grad(self, set_to_none: bool = True):
        self.grad.zero_()
        self._grads_dirty = False

    def named_grads(self) -> Dict[str, torch.Tensor]:
        return dict(self._gviews)

    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False) -> Dict[str, torch.Tensor]:
        """Reference-compatible names: trainable tensors + embedding.weight + lm_head.{weight,bias}."""
        sd = {k: v.detach().clone() for k, v in self._views.items()}
        if self.hp["TRAIN_EMBEDDING"]:  # embedding.weight / lm_head.weight are trainable views already; no lm_head bias (:239)
            return sd
        sd["embedding.weight"] = self.embedding_weight.clone()
        sd["lm_head.weight"] = self.lm_head_weight.clone()
        sd["lm_head.bias"] = torch.zeros(self.hp["VOCAB_SIZE"], device=self.device)
        return sd

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True, assign: bool = False):
        """nn.Module semantics: strict => every key of state_dict() must be present (lm_head.* may be absent: tied to embedding.weight,
        bias 0) and no unknown key; shapes are checked before anything is copied."""
        hp = self.hp
        te = hp["TRAIN_EMBEDDING"]
        frozen = {} if te else {"embedding.weight": self.embedding_weight, "lm_head.weight": self.lm_head_weight}
        optional = set() if te else {"lm_head.weight", "lm_head.bias"}
        known = set(self._views) | set(frozen) | optional
        missing = [k for k in list(self._views) + list(frozen) if k not in sd and k not in optional]
        unexpected = [k for k in sd if k not in known]
        if strict and (missing or unexpected):
            raise KeyError(f"load_state_dict: missing keys {missing[:4]}{'...' if len(missing) > 4 else ''}, "
                           f"unexpected keys {unexpected[:4]}{'...' if len(unexpected) > 4 else ''}")
        for k, v in list(self._views.items()) + list(frozen.items()):
            if k in sd and tuple(sd[k].shape) != tuple(v.shape):
                raise ValueError(f"load_state_dict: {k} has shape {tuple(sd[k].shape)}, expected {tuple(v.shape)}")
        for k, v in self._views.items():
            if k in sd:
                v.copy_(sd[k].detach().to(self.device, torch.float32))
        if te:
            self.sync_shadow()
            return
        if "embedding.weight" in sd:
            self.embedding_weight.copy_(sd["embedding.weight"].detach().to(self.device, torch.float32))
        if "lm_head.weight" in sd:
            self.lm_head_weight.copy_(sd["lm_head.weight"].detach().to(self.device, torch.float32))
        elif "embedding.weight" in sd:
            self.lm_head_weight.copy_(self.embedding_weight)
        self.sync_shadow()

    def sync_shadow(self):
        """Refresh the bf16 (pair) copies the GEMMs read, after any out-of-band change of the fp32 master weights."""
        lib, st = L.load(), self._stream()
        self._lm_head_max_norm = None   # the lm_head weight may have been replaced (load_state_dict)
        with torch.cuda.device(self.device):
            L.check(lib.clipdlm_to_bf16(L.ptr(self.flat), L.ptr(self.shadow_hi), L.ptr(self.shadow_lo), self.n_params, st))
            if self.hp["TRAIN_EMBEDDING"]:
                return  # the lm_head operand is part of the flat shadow
            n = self.hp["VOCAB_SIZE"] * self.hp["DIM"]
            L.check(lib.clipdlm_to_bf16(L.ptr(self.lm_head_weight), L.ptr(self.emb_hi), L.ptr(self.emb_lo), n, st))

    EXP_SHIFT_MAX = 60.0    # with c <= 60 a row whose largest logit is >= -27 still has exp(s - c) far above underflow
    EXP_ARG_MAX = 69.0      # exp(s - c) stays below the kernel's 2^100 clamp while s - c <= 69

    def refresh_exp_shift(self):
        """Exponent shift c of the factored softmax gradient (clipdlm.h CLIPDLM_OPT_EXP_SHIFT_PTR), from a bound that needs no pass
        over the logits: x_out = LN_v(.) has |x_out| <= sqrt(D) max|w| + |b|, so every logit is <= B = that * max_v |W_v| (Cauchy-Schwarz);
        under classifier-free guidance the lm_head reads the mix (1 + w) guided - w unguided, whose norm is up to (1 + 2w) times that.
        c = clamp(B - 69, 0, 60): exp(s - c) can then not reach the 2^100 clamp (s - c <= 69), and for B <= 69 - the usual case, B ~ 17
        at random init, ~ 40-70 with pretrained BERT embeddings - c = 0 and the stored values are plain exp(s). A few tiny torch
        kernels on the current stream, no host sync; called once per train_func / loss call. B > 129 would need c > 60: the bound
        is copied to pinned memory and looked at one call later (an event query, still no sync) - past it the model switches itself
        to the in-place softmax-gradient path for good instead of saturating silently."""
        if self._exp_shift is None or not self.fused_softmax_grad:
            return
        if self._exp_bound_event is not None and self._exp_bound_event.query():
            self._check_exp_bound(float(self._exp_bound_host[0]))
            if not self.fused_softmax_grad:
                return
        bound = self._logit_bound()
        self._exp_shift.copy_(self._shift_from_bound(bound).reshape(1))
        if self._exp_bound_host is None:    # first call (engine creation, off the step path): look at the bound right away
            self._exp_bound_host = torch.empty(1, pin_memory=True)
            self._exp_bound_event = torch.cuda.Event()
            self._check_exp_bound(float(bound.item()))
            if not self.fused_softmax_grad:
                return
        self._exp_bound_host.copy_(bound.reshape(1), non_blocking=True)
        self._exp_bound_event.record(torch.cuda.current_stream(self.device))

    def _logit_bound(self) -> torch.Tensor:
        """(sqrt(D) max|w_LN| + |b_LN|) max_v |W_v| (1 + 2 CLASSIFIER_FREE_WEIGHT): 0-dim tensor, no host sync."""
        if getattr(self, "_lm_head_max_norm", None) is None:
            self._lm_head_max_norm = self.lm_head_weight.norm(dim=1).max()   # frozen (CLIP-DDPM.py:246-247): once
        w, b = self._views["model.vocab_layer_norm.weight"], self._views["model.vocab_layer_norm.bias"]
        cfg_w = max(float(self.hp.get("CLASSIFIER_FREE_WEIGHT", 0.0)), 0.0)
        return (math.sqrt(self.hp["DIM"]) * w.abs().max() + b.norm()) * self._lm_head_max_norm * (1.0 + 2.0 * cfg_w)

    @classmethod
    def _shift_from_bound(cls, bound: torch.Tensor) -> torch.Tensor:
        return (bound - cls.EXP_ARG_MAX).clamp(0.0, cls.EXP_SHIFT_MAX)

    def _check_exp_bound(self, bound: float):
        if bound <= self.EXP_ARG_MAX + self.EXP_SHIFT_MAX and math.isfinite(bound):
            return
        import warnings
        warnings.warn(f"clipdlm: logit bound {bound:.1f} exceeds {self.EXP_ARG_MAX + self.EXP_SHIFT_MAX:.0f}; the factored softmax gradient "
                      "is switched off for this model (in-place softmax-gradient path from here on)")
        self.fused_softmax_grad = False
        for e in self._engines.values():
            L.check(L.load().clipdlm_engine_set_option(e[0], L.OPT_FUSED_SOFTMAX_GRAD, 0))

    # ---------------------------------------------------------------------------------------------------------- nn.Module-ish
    def train(self, mode: bool = True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def _apply(self, fn, recurse=True):
        return self   # the module owns device buffers laid out by libclipdlm: dtype / device casts do not apply

    def to(self, *args, **kwargs):
        """`model.to(device)` (CLIP-DDPM.py:506,552,561,572). The compute buffers never leave the GPU; a CUDA target returns self, a CPU target is
        what `.cpu()` does."""
        dev = kwargs.get("device", args[0] if args else None)
        if isinstance(dev, (str, torch.device)) and torch.device(dev).type == "cpu":
            return self.cpu()
        return self

    def cuda(self, device=None):
        return self

    def cpu(self):
        """The reference pickles `model.cpu()` and moves the model back with `.to(device)` (CLIP-DDPM.py:551-552,560-561). Pickling this module
        always serialises host copies (`__getstate__`), so `.cpu()` has nothing to move: it returns self, `torch.save(model.cpu(), path)` writes
        a file `torch.load(path).to(device)` turns back into a working model, and training continues on the same device buffers."""
        return self

    def __getstate__(self):
        hp = {k: (float(v) if torch.is_tensor(v) else v) for k, v in self.hp.items()}   # a device-resident dynamic ROUNDING_WEIGHT pickles as its value
        return {"clipdlm_module": 1, "hp": hp, "ctor": dict(self._ctor), "training": self.training,
                "state_dict": {k: v.cpu() for k, v in self.state_dict().items()}}

    def __setstate__(self, st):
        """Unpickling (torch.load of a whole-module pickle, CLIP-DDPM.py:506,570) rebuilds the device buffers on the current CUDA device."""
        sd = st["state_dict"]
        emb = None if st["hp"]["TRAIN_EMBEDDING"] else sd["embedding.weight"]
        proj = None if st["hp"]["TRAIN_EMBEDDING"] else sd.get("lm_head.weight", emb)
        DistilBertModel.__init__(self, emb, proj, None, hp=st["hp"], **st["ctor"])
        self.load_state_dict(sd)
        self.train(st.get("training", True))

    # ---------------------------------------------------------------------------------------------------------- engine plumbing
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _engine(self, rows: int, batch: int, training: bool, tag: str = ""):
        """Engine (+ its workspace) able to run passes of up to `rows` rows over `batch` captions. One engine is kept per
        (training, tag) flavour and regrown on demand (tag "g" = the guided pass of classifier-free-guidance training)."""
        key = (bool(training), tag)
        cur = self._engines.get(key)
        if cur is not None and cur[1] >= rows and cur[2] >= batch:
            return cur[0]
        lib = L.load()
        if cur is not None:
            rows, batch = max(rows, cur[1]), max(batch, cur[2])
            self._launches_retired += int(lib.clipdlm_engine_launch_count(cur[0]))
            lib.clipdlm_engine_destroy(cur[0])
            del self._engines[key]
            cur = None
        need = int(lib.clipdlm_workspace_bytes(C.byref(self._cfg), rows, batch, 1 if training else 0))
        if need == 0:
            raise L.ClipdlmError("workspace query failed: " + lib.clipdlm_last_error().decode())
        ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        if self.hp["TRAIN_EMBEDDING"]:
            # the engine's own embedding gather / frozen lm_head stay unused in this mode (x enters through input_projection, the
            # lm_head is composed in train_embedding.py); its non-null checks get the flat buffers as placeholders
            bufs = L.Buffers(L.ptr(self.flat), L.ptr(self.grad), L.ptr(self.shadow_hi), L.ptr(self.shadow_lo), L.ptr(self.flat),
                             L.ptr(self.shadow_hi), L.ptr(self.shadow_lo), L.ptr(ws), need)
        else:
            bufs = L.Buffers(L.ptr(self.flat), L.ptr(self.grad), L.ptr(self.shadow_hi), L.ptr(self.shadow_lo), L.ptr(self.embedding_weight),
                             L.ptr(self.emb_hi), L.ptr(self.emb_lo), L.ptr(ws), need)
        h = lib.clipdlm_engine_create(C.byref(self._cfg), C.byref(bufs), rows, batch, 1 if training else 0)
        if not h:
            raise L.ClipdlmError("engine_create failed: " + lib.clipdlm_last_error().decode())
        self.__dict__.pop("_sample_graphs", None)   # CUDA graphs of sample() hold raw pointers into engine workspaces: a new engine invalidates them
        self._engines[key] = (h, rows, batch, ws)
        if self.fused_softmax_grad and self._exp_shift is None:
            self._exp_shift = torch.zeros(1, device=self.device)
            self.refresh_exp_shift()   # may switch the option off (logit bound out of range)
        if self.fused_softmax_grad:
            L.check(lib.clipdlm_engine_set_option(h, L.OPT_EXP_SHIFT_PTR, self._exp_shift.data_ptr()))
            L.check(lib.clipdlm_engine_set_option(h, L.OPT_FUSED_SOFTMAX_GRAD, 1))
        if self.gelu_deriv_store:
            L.check(lib.clipdlm_engine_set_option(h, L.OPT_GELU_DERIV_STORE, self.gelu_deriv_store))
        if getattr(self, "_profiling", False):
            L.check(lib.clipdlm_engine_profile(h, 1))
        return h

    def _scratch(self, name: str, shape, dtype) -> torch.Tensor:
        """Named reusable device buffer (host-composed paths allocate nothing per step once shapes have been seen)."""
        n = int(math.prod(shape))
        t = self._scr.get(name)
        if t is None or t.dtype != dtype or t.numel() < n:
            t = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._scr[name] = t
        return t[:n].view(shape)

    def rounding_weight(self):
        """ROUNDING_WEIGHT as (host factor, device scalar or None). A float (the reference's default, CLIP-DDPM.py:75) goes to the kernels by value; a
        tensor - what the reference's dynamic rounding weight is (:535-536) - stays on the device: the kernels multiply by it, nothing is read back."""
        rw = self.hp["ROUNDING_WEIGHT"]
        if torch.is_tensor(rw):
            return 1.0, rw.detach().to(self.device, torch.float32).reshape(1)
        return float(rw), None

    def launch_count(self) -> int:
        """Kernel launches issued by this model's engines so far (bench `gpu_launches`)."""
        lib = L.load()
        return self._launches_retired + sum(int(lib.clipdlm_engine_launch_count(e[0])) for e in self._engines.values())

    def profile(self, enable: bool = True):
        """Bracket every engine launch with CUDA events (per-kernel-category device times for the roofline report)."""
        self._profiling = bool(enable)
        for e in self._engines.values():
            L.check(L.load().clipdlm_engine_profile(e[0], 1 if enable else 0))

    def profile_read(self, reset: bool = True) -> Dict[str, dict]:
        lib = L.load()
        tot = {k: dict(ms=0.0, flops=0.0, bytes=0.0, launches=0) for k in L.PROF_CATEGORIES}
        for e in self._engines.values():
            arr = (L.Prof * len(L.PROF_CATEGORIES))()
            L.check(lib.clipdlm_engine_profile_read(e[0], C.cast(arr, C.c_void_p), 1 if reset else 0))
            for k, r in zip(L.PROF_CATEGORIES, arr):
                tot[k]["ms"] += r.ms; tot[k]["flops"] += r.flops; tot[k]["bytes"] += r.bytes; tot[k]["launches"] += int(r.launches)
        return tot

    def __del__(self):
        try:
            lib = L.load()
            for e in self._engines.values():
                lib.clipdlm_engine_destroy(e[0])
        except Exception:
            pass

    def _run_forward(self, eng, *, R, B, mode, guided, train, image_clip, text_clip, attn_mask, x_in=None, x_in_stride=0, ids=None,
                     noise=None, coef_a=None, coef_b=None, drop_seed=0, x_out=None, reuse_proj=False):
        p = L.Pass(R, B, mode, 1 if guided else 0, 1 if train else 0, 1 if reuse_proj else 0, L.ptr(x_in), x_in_stride, L.ptr(ids), L.ptr(noise), L.ptr(coef_a),
                   L.ptr(coef_b), L.ptr(image_clip), L.ptr(text_clip), L.ptr(attn_mask), drop_seed, L.ptr(x_out))
        with torch.cuda.device(self.device):
            L.check(L.load().clipdlm_engine_forward(eng, C.byref(p), self._stream()))

    # ---------------------------------------------------------------------------------------------------------- forward
    @staticmethod
    def _f32(t: torch.Tensor) -> torch.Tensor:
        return t.detach().to(torch.float32).contiguous()

    def forward(self, x, image_clip, text_clip, mask, concat_mask):
        """CLIP-DDPM.py:271-323. x [R, MAX_LENGTH, IN_CHANNEL]; image_clip / text_clip [R, 1, 512]; mask [R, MAX_LENGTH];
        concat_mask [R, 2]. Returns (vocab_out [R, MAX_LENGTH, VOCAB_SIZE], feature_out [R, L, IN_CHANNEL]) in fp32.
        Forward only (inference / parity checks): training goes through loss() / train_func(), which run the hand-written
        backward chunk by chunk."""
        hp = self.hp
        R = x.shape[0]
        ML, D = hp["MAX_LENGTH"], hp["IN_CHANNEL"]
        assert tuple(x.shape) == (R, ML, D)  # :284-287
        assert tuple(image_clip.shape) == tuple(text_clip.shape) == (R, 1, hp["CLIP_DIM"])
        assert tuple(mask.shape) == (R, ML)
        assert tuple(concat_mask.shape) == (R, 2)
        guidance = concat_mask[:, 1] == 1  # :290
        w = hp["CLASSIFIER_FREE_WEIGHT"]
        te = hp["TRAIN_EMBEDDING"]
        if te:  # :292-293
            from . import train_embedding as TE
            x = TE.in_proj(self, self._f32(x)).clone()
        x_out = self._encode(x, image_clip[:, 0], text_clip[:, 0], mask, guided=False)
        mixed = w > 0
        if mixed:  # :313-317. The reference gathers the guided rows (a host sync on the row count); here the guided pass runs over every row
            # and the mix is selected per row on the device - no .item(), same values on the guided rows, untouched elsewhere.
            guided_out = self._encode(x, image_clip[:, 0], text_clip[:, 0], mask, guided=True)
            x_out = torch.where(guidance.to(self.device)[:, None, None], (1 + w) * guided_out - w * x_out, x_out)
        if te:  # :319-320,323
            y = TE.out_proj(self, x_out).clone()
            return TE.logits(self, y), y
        if mixed:
            return self.lm_head(x_out[:, :ML, :]), x_out
        return self._lm_head_last(R), x_out

    def _encode(self, x, image_clip, text_clip, mask, guided: bool, x_stride: int = 0):
        hp = self.hp
        R = x.shape[0]
        Lfull = hp["MAX_LENGTH"] + (2 if hp["CLIP_ADDING_METHOD"] == "concat" else 0)
        eng = self._engine(R, R, False)
        x_out = torch.empty(R, Lfull, hp["DIM"], device=self.device)
        xin = x if x_stride else self._f32(x)
        self._run_forward(eng, R=R, B=R, mode=0, guided=guided, train=False, image_clip=self._f32(image_clip),
                          text_clip=self._f32(text_clip), attn_mask=(mask != 0).to(torch.int32).contiguous(), x_in=xin,
                          x_in_stride=x_stride, x_out=x_out)
        self._last_eng = eng
        return x_out

    def _lm_head_last(self, R: int) -> torch.Tensor:
        """Dense logits of the rows of the last forward (fp32; padding columns sliced away)."""
        hp = self.hp
        V, ML = hp["VOCAB_SIZE"], hp["MAX_LENGTH"]
        n32 = (V + 31) // 32 * 32
        logits = torch.empty(R * ML, n32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.load().clipdlm_engine_lm_head(self._last_eng, L.ptr(logits), n32, None, self._stream()))
        return logits.view(R, ML, n32)[:, :, :V]

    def _lm_head_dense(self, x: torch.Tensor) -> torch.Tensor:
        """lm_head on an arbitrary [..., DIM] tensor through the tcgen05 GEMM (used by forward()'s CFG branch and demos)."""
        hp = self.hp
        V, D = hp["VOCAB_SIZE"], hp["DIM"]
        lead = x.shape[:-1]
        xf = self._f32(x).view(-1, D)
        M = xf.shape[0]
        hi = torch.empty(M, D, device=self.device, dtype=torch.bfloat16)
        lo = torch.empty(M, D, device=self.device, dtype=torch.bfloat16) if self.precision == "bf16x3" else None
        n32 = (V + 31) // 32 * 32
        out = torch.empty(M, n32, device=self.device)
        lib, st = L.load(), self._stream()
        with torch.cuda.device(self.device):
            L.check(lib.clipdlm_to_bf16(L.ptr(xf), L.ptr(hi), L.ptr(lo), M * D, st))
            g = L.Gemm()
            g.a_hi, g.a_lo, g.b_hi, g.b_lo = L.ptr(hi), L.ptr(lo), L.ptr(self.emb_hi), L.ptr(self.emb_lo)
            g.lda, g.ldb, g.M, g.N, g.K = D, D, M, n32, D
            g.epilogue, g.out_f32, g.ldo = L.EPI_STORE, L.ptr(out), n32
            L.check(lib.clipdlm_gemm(C.byref(g), st))
        return out.view(*lead, n32)[..., :V]

    def argmax_last(self, R: int) -> torch.Tensor:
        """argmax over the vocabulary of lm_head(x_out[:, :MAX_LENGTH]) of the last forward, via the fused GEMM + running
        argmax epilogue (softmax(...).argmax(-1), CLIP-DDPM.py:620). int64 [R, MAX_LENGTH]."""
        ML = self.hp["MAX_LENGTH"]
        out = torch.empty(R * ML, device=self.device, dtype=torch.int32)
        with torch.cuda.device(self.device):
            L.check(L.load().clipdlm_engine_lm_head(self._last_eng, None, 0, L.ptr(out), self._stream()))
        return out.view(R, ML).to(torch.int64)


class AdamW:
    """`torch.optim.AdamW(model.parameters(), lr=...)` (CLIP-DDPM.py:335) over the model's flat buffers: one fused kernel per
    step (decoupled weight decay on every element — the reference has a single param group, so biases / LayerNorm decay too),
    which also refreshes the bf16 shadow weights and clears the gradient buffer for the next step."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        owner = getattr(params, "owner", None)
        if owner is None:
            raise TypeError("AdamW expects model.parameters() of a clipdlm DistilBertModel")
        self.model = owner
        self.param_groups = [dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)]
        self._slice = (0, owner.n_params)
        if owner.dp_fused is not None:  # fused data-parallel step: this rank only holds the moments of its own slice (ZeRO-1)
            self._slice = owner.dp_fused["slice"]
        n = self._slice[1] - self._slice[0]
        self.m = torch.zeros(max(n, 4), device=owner.device)
        self.v = torch.zeros(max(n, 4), device=owner.device)
        self.t = 0

    def zero_grad(self, set_to_none: bool = True):
        if self.model._grads_dirty:  # step() already cleared them inside the AdamW kernel
            self.model.grad.zero_()
            self.model._grads_dirty = False

    def step(self):
        g = self.param_groups[0]
        m = self.model
        self.t += 1
        if m.dp_fused is not None:
            f = m.dp_fused
            if self._slice != f["slice"]:
                raise L.ClipdlmError("AdamW was created before enable_data_parallel(): create the optimizer after it")
            b0 = self._slice[0]
            with torch.cuda.device(m.device):
                f["barrier"](0)   # every rank's backward has finished writing its gradient buffer
                L.check(L.load().clipdlm_adamw_dp(C.byref(f["bufs"]), self.m.data_ptr() - 4 * b0, self.v.data_ptr() - 4 * b0, m.n_params, g["lr"],
                                                  g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self.t, 1.0 / m.dp_world, m._stream()))
                f["barrier"](1)   # every rank's slice of the new weights has landed in every copy
                m.grad.zero_()
            m._grads_dirty = False
            return
        with torch.cuda.device(m.device):
            L.check(L.load().clipdlm_adamw(L.ptr(m.flat), L.ptr(m.grad), L.ptr(self.m), L.ptr(self.v), L.ptr(m.shadow_hi), L.ptr(m.shadow_lo),
                                           m.n_params, g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self.t,
                                           1.0 / m.dp_world, 1, m._stream()))
        m._grads_dirty = False  # the kernel zeroed them

    def state_dict(self):
        """This rank's optimizer state. Under the fused data-parallel step the moments cover only `slice` (ZeRO-1): use full_state_dict()
        (collective) for a file every rank can resume from."""
        n = self._slice[1] - self._slice[0]
        return {"m": self.m[:n].clone(), "v": self.v[:n].clone(), "t": self.t, "param_groups": [dict(g) for g in self.param_groups],
                "slice": tuple(self._slice), "n_params": self.model.n_params}

    def full_state_dict(self):
        """Optimizer state with FULL-size moments. Single device / NCCL exchange: same as state_dict(). Fused data-parallel step: COLLECTIVE
        (every rank of the model's group must call it) - each rank's slice of m / v is broadcast into a full-size tensor."""
        sd = self.state_dict()
        mdl = self.model
        if self._slice == (0, mdl.n_params):
            return sd
        import torch.distributed as dist
        world = dist.get_world_size(mdl.dp_group)
        m, v = torch.zeros(mdl.n_params, device=mdl.device), torch.zeros(mdl.n_params, device=mdl.device)
        b, e = C.c_int64(), C.c_int64()
        for r in range(world):
            L.check(L.load().clipdlm_dp_slice(mdl.n_params, r, world, C.byref(b), C.byref(e)))
            lo, hi = int(b.value), int(e.value)
            if r == dist.get_rank(mdl.dp_group):
                assert (lo, hi) == tuple(self._slice)
                m[lo:hi].copy_(self.m[:hi - lo]); v[lo:hi].copy_(self.v[:hi - lo])
            if hi > lo:
                dist.broadcast(m[lo:hi], src=dist.get_global_rank(mdl.dp_group, r), group=mdl.dp_group)
                dist.broadcast(v[lo:hi], src=dist.get_global_rank(mdl.dp_group, r), group=mdl.dp_group)
        sd.update(m=m, v=v, slice=(0, mdl.n_params))
        return sd

    def load_state_dict(self, sd):
        """Accepts full-size moments (any rank layout: this rank's slice is cut out) or a state saved for exactly this rank's slice; anything
        else - e.g. rank 0's slice file loaded on rank 1 - raises instead of silently applying moments to the wrong parameter range."""
        n_params = self.model.n_params
        lo, hi = self._slice
        m, v = sd["m"], sd["v"]
        src = tuple(sd.get("slice", (0, m.numel())))
        if sd.get("n_params", n_params) != n_params:
            raise ValueError(f"optimizer state was saved for {sd['n_params']} parameters, this model has {n_params}")
        if src == (0, n_params) and m.numel() >= n_params:
            m, v = m[lo:hi], v[lo:hi]
        elif src != (lo, hi):
            raise ValueError(f"optimizer state covers parameter slice {src}, this rank owns {(lo, hi)}: save with full_state_dict() "
                             "(save_checkpoint does) to resume under a different rank layout")
        if m.numel() < hi - lo or v.numel() < hi - lo:
            raise ValueError(f"optimizer moments hold {m.numel()} elements, slice {(lo, hi)} needs {hi - lo}")
        self.m[:hi - lo].copy_(m[:hi - lo]); self.v[:hi - lo].copy_(v[:hi - lo]); self.t = int(sd["t"])
        self.param_groups = [dict(g) for g in sd["param_groups"]]
