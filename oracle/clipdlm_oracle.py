"""ORACLE — CPU restatement of the CLIP-Diffusion-LM hot path (xu-shitong/diffusion-image-captioning, CLIP-DDPM.py).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it, and only as the checker / the reported CPU baseline.  The product path
(diffusion-image-captioning_b200/) never imports anything from oracle/.

What it is: a self-contained fp32 torch restatement of the reference's algorithm for the path — the diffusion
schedule, q_sample, the CLIP concat/add fusion, the DistilBERT encoder arithmetic (third-party HF `transformers`,
un-vendored and unpinned in the reference; observed 4.21.1 by the authors, 5.5.0 in this image — same math for this
path), the frozen lm_head, the four LOSS_FUNCs, the rounding cross-entropy, AdamW and the fixed-point denoise loop.
The encoder is written out with plain tensor ops (no `transformers` import) so it runs on the GPU box where
/root/reference does not exist. Every function cites the reference file:line (or HF file:line) it follows.

Parity pinning: the reference has NO tests / golden vectors for this path.  The restatement is
pinned instead against outputs of the reference ITSELF executed in the build container (the exec'd source slices of
CLIP-DDPM.py driving the real HF DistilBertForMaskedLM): oracle/validate_against_reference.py checks it live, and
tests/golden/make_golden.py commits the resulting vectors as fixtures which tests/test_oracle_golden.py replays.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------------------------
# Hyperparameters: the reference's module globals (CLIP-DDPM.py:55-114), same names and defaults
# ------------------------------------------------------------------------------------------------------------------
def default_hparams() -> dict:
    return dict(
        DEBUG=False, CONTINUE_TRAIN=False, BATCH_SIZE=8, MAX_LENGTH=16, LEARNING_RATE=1e-4, END_LEARNING_RATE=5e-5,
        SCHEDULER="linspace", TRAIN_SET_RATIO=0.8, EARLY_STOP_RATIO=1.05, EPOCH_NUM=5, DYNAMIC_ROUNDING_WEIGHT=-1,
        ROUNDING_WEIGHT=0.5, LOSS_FUNC="series_sum_sample_mean", CLIP_ADDING_METHOD="concat", CLASSIFIER_FREE_WEIGHT=0,
        CLASSIFIER_FREE_PROB=0.2, TRAIN_EMBEDDING=False, IN_CHANNEL=768, BETA_MIN=0.0001, BETA_MAX=0.02, STEP_TOT=1000,
        COSIN_SCHEDULE=True, SAMPLE_SIZE=100, X_0_PREDICTION=True, X_T_STEP_INTERVAL=100, USE_X_T_LOSS=True,
        USE_X_1_LOSS=True, USE_PROB_LOSS=True, VOCAB_SIZE=30522,
        # implicit: torch.optim.AdamW defaults (CLIP-DDPM.py:335) and DistilBertConfig defaults (HF configuration_distilbert.py:58-74)
        ADAM_BETAS=(0.9, 0.999), ADAM_EPS=1e-8, WEIGHT_DECAY=0.01,
        N_LAYERS=6, DIM=768, N_HEADS=12, HIDDEN_DIM=3072, DROPOUT=0.1, ATTENTION_DROPOUT=0.1, MAX_POSITION=512, CLIP_DIM=512,
    )


# ------------------------------------------------------------------------------------------------------------------
# LOSS_FUNCs (CLIP-DDPM.py:77-87)
# ------------------------------------------------------------------------------------------------------------------
def series_sum_sample_mean(x_hat: Tensor, x: Tensor, hp: dict) -> Tensor:  # :77-78
    return (x_hat - x).abs().sum(dim=1).mean()


def series_sum(x_hat: Tensor, x: Tensor, hp: dict) -> Tensor:  # :80-81  (768 and 100 are literals in the reference)
    return (x_hat - x).abs().sum() / hp["BATCH_SIZE"] / 768 / 100


def mse_series_mean(x_hat: Tensor, x: Tensor, hp: dict) -> Tensor:  # :83-84
    return ((x_hat - x) ** 2).sum(dim=[-2, -1]).sqrt().mean()


def mse_series_sum(x_hat: Tensor, x: Tensor, hp: dict) -> Tensor:  # :86-87
    return ((x_hat - x) ** 2).sum(dim=[-2, -1]).sqrt().sum() / hp["BATCH_SIZE"]


LOSS_FUNCS: Dict[str, Callable] = {
    "series_sum_sample_mean": series_sum_sample_mean, "series_sum": series_sum,
    "mse_series_mean": mse_series_mean, "mse_series_sum": mse_series_sum,
}
LOSS_KIND = {"series_sum_sample_mean": 0, "series_sum": 1, "mse_series_mean": 2, "mse_series_sum": 3}


# ------------------------------------------------------------------------------------------------------------------
# Diffusion schedule and q_sample (CLIP-DDPM.py:337-362)
# ------------------------------------------------------------------------------------------------------------------
def alpha_cumprod(hp: dict, device="cpu") -> Tensor:
    if hp["COSIN_SCHEDULE"]:  # :337-342
        def scheduler(t):
            s = 0.008
            return torch.cos(math.pi / 2 * (t / hp["STEP_TOT"] + s) / (1 + s)) ** 2
        ts = torch.arange(hp["STEP_TOT"]).to(device)
        return scheduler(ts) / scheduler(torch.zeros(1, device=device))
    betas = torch.hstack([torch.zeros(1), torch.linspace(hp["BETA_MIN"], hp["BETA_MAX"], hp["STEP_TOT"])]).to(device)  # :344-346
    return torch.cumprod((1 - betas)[:-1], 0)


def diffuse_t(x: Tensor, t: Tensor, acp: Tensor, noise: Optional[Tensor] = None) -> Tensor:
    """q_sample, CLIP-DDPM.py:347-362. One noise draw of x.shape shared by all t.numel() samples; row = s*B + b."""
    batch_size, seq_len, ch = x.shape
    sample_shape = (t.numel(), *(1,) * len(x.shape))
    if noise is None:
        noise = torch.normal(0, 1, x.shape).to(x.device)  # :359 (CPU generator, then copied)
    mean = torch.sqrt(acp[t].reshape(sample_shape)) * x
    epsilon = noise * torch.sqrt(1 - acp[t]).reshape(sample_shape)
    return (mean + epsilon).reshape((t.numel() * batch_size, seq_len, ch))


# ------------------------------------------------------------------------------------------------------------------
# Parameters: names follow named_parameters() of the reference DistilBertModel wrapper
# ------------------------------------------------------------------------------------------------------------------
def param_names(hp: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    d, f, c = hp["DIM"], hp["HIDDEN_DIM"], hp["CLIP_DIM"]
    out = [("model.distilbert.embeddings.position_embeddings.weight", (hp["MAX_POSITION"], d)),
           ("model.distilbert.embeddings.LayerNorm.weight", (d,)), ("model.distilbert.embeddings.LayerNorm.bias", (d,))]
    for i in range(hp["N_LAYERS"]):
        p = f"model.distilbert.transformer.layer.{i}."
        for lin in ("q_lin", "k_lin", "v_lin", "out_lin"):
            out += [(p + f"attention.{lin}.weight", (d, d)), (p + f"attention.{lin}.bias", (d,))]
        out += [(p + "sa_layer_norm.weight", (d,)), (p + "sa_layer_norm.bias", (d,)),
                (p + "ffn.lin1.weight", (f, d)), (p + "ffn.lin1.bias", (f,)),
                (p + "ffn.lin2.weight", (d, f)), (p + "ffn.lin2.bias", (d,)),
                (p + "output_layer_norm.weight", (d,)), (p + "output_layer_norm.bias", (d,))]
    out += [("model.vocab_transform.weight", (d, d)), ("model.vocab_transform.bias", (d,)),
            ("model.vocab_layer_norm.weight", (d,)), ("model.vocab_layer_norm.bias", (d,)),
            ("image_linear.weight", (d, c)), ("image_linear.bias", (d,)),
            ("text_linear.weight", (d, c)), ("text_linear.bias", (d,))]
    if hp["TRAIN_EMBEDDING"]:  # :238-243,260-262: learned IN_CHANNEL-wide embedding, lm_head and in/out projections, all trainable
        ch, V = hp["IN_CHANNEL"], hp["VOCAB_SIZE"]
        out += [("embedding.weight", (V, ch)), ("lm_head.weight", (V, ch)),
                ("input_projection.weight", (d, ch)), ("input_projection.bias", (d,)),
                ("output_projection.weight", (ch, d)), ("output_projection.bias", (ch,))]
    if hp["CLIP_ADDING_METHOD"] == "concat":
        out += [("segment_embedding.weight", (2, d))]
    return out


def init_params(hp: dict, seed: int = 0, closed_form: bool = False) -> Dict[str, Tensor]:
    """Trainable parameters + the frozen `embedding.weight` (== lm_head.weight values, lm_head.bias = 0; CLIP-DDPM.py:245-247).

    closed_form=False: HF-style random init (N(0, 0.02) weights, LN = (1, 0), zero biases; CLIP linears Kaiming-uniform-like;
    segment N(0,1)) from a torch generator.  closed_form=True: a generator-free integer-hash formula, reproducible on any platform,
    used by the committed golden fixtures.
    """
    g = torch.Generator().manual_seed(seed)
    params: Dict[str, Tensor] = {}
    names = param_names(hp) + ([] if hp["TRAIN_EMBEDDING"] else [("embedding.weight", (hp["VOCAB_SIZE"], hp["DIM"]))])
    for k, (name, shape) in enumerate(names):
        n = int(math.prod(shape))
        if closed_form:
            base = hash_uniform(n, 1000 * seed + k).reshape(shape)  # unit variance, platform independent
            if "LayerNorm.weight" in name or "layer_norm.weight" in name:
                v = 1.0 + 0.05 * base
            elif name.endswith(".bias"):
                v = 0.02 * base
            elif name == "segment_embedding.weight":
                v = 0.7 * base
            elif name.startswith(("image_linear", "text_linear")):
                v = base / math.sqrt(hp["CLIP_DIM"])
            elif hp["TRAIN_EMBEDDING"] and name == "embedding.weight":
                v = base  # nn.Embedding default N(0, 1) (:238)
            elif hp["TRAIN_EMBEDDING"] and name.startswith(("lm_head", "input_projection", "output_projection")):
                v = 0.5 * base / math.sqrt(shape[-1])  # nn.Linear-like scale (fan-in = last dim)
            else:
                v = 0.03 * base
            params[name] = v.float()
        else:
            if "LayerNorm.weight" in name or "layer_norm.weight" in name:
                params[name] = torch.ones(shape)
            elif name.startswith(("image_linear", "text_linear")):
                bound = 1.0 / math.sqrt(hp["CLIP_DIM"])
                params[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
            elif hp["TRAIN_EMBEDDING"] and name.startswith(("lm_head", "input_projection", "output_projection")):
                bound = 1.0 / math.sqrt(shape[-1] if name.endswith("weight") else (hp["IN_CHANNEL"] if name.startswith("input") else hp["DIM"]))
                params[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound  # nn.Linear default init
            elif name.endswith(".bias"):
                params[name] = torch.zeros(shape)
            elif name == "segment_embedding.weight" or (hp["TRAIN_EMBEDDING"] and name == "embedding.weight"):
                params[name] = torch.randn(shape, generator=g)
            else:
                params[name] = torch.randn(shape, generator=g) * 0.02
    return params


def trainable_names(hp: dict) -> List[str]:
    return [n for n, _ in param_names(hp)]


# ------------------------------------------------------------------------------------------------------------------
# Encoder arithmetic (HF transformers/models/distilbert/modeling_distilbert.py)
# ------------------------------------------------------------------------------------------------------------------
def _ln(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-12)  # eps=1e-12: HF :90,241,244,467


def _drop(x: Tensor, p: float, train: bool) -> Tensor:
    return F.dropout(x, p, training=train)


def encoder(P: Dict[str, Tensor], x: Tensor, key_mask: Tensor, hp: dict, train: bool = False) -> Tensor:
    """HF DistilBertForMaskedLM.forward with identity word/vocab embeddings (CLIP-DDPM.py:249-250,312).

    x [R, L, D] float (passed as `input_ids`, identity embedding), key_mask [R, L] (non-zero = visible key).
    Embeddings HF:96-122; mask -> additive -inf on masked keys (masking_utils create_bidirectional_mask / :142-143);
    block HF:245-263; attention HF:126-151,177-207; FFN HF:223-228; MLM head HF:514-517 (projector = identity).
    """
    R, L, D = x.shape
    H = hp["N_HEADS"]
    dh = D // H
    pe = "model.distilbert.embeddings."
    h = x + P[pe + "position_embeddings.weight"][:L]
    h = _ln(h, P[pe + "LayerNorm.weight"], P[pe + "LayerNorm.bias"])
    h = _drop(h, hp["DROPOUT"], train)
    add_mask = torch.zeros(R, 1, 1, L, dtype=x.dtype, device=x.device).masked_fill(key_mask[:, None, None, :] == 0, float("-inf"))
    for i in range(hp["N_LAYERS"]):
        p = f"model.distilbert.transformer.layer.{i}."
        q = F.linear(h, P[p + "attention.q_lin.weight"], P[p + "attention.q_lin.bias"]).view(R, L, H, dh).transpose(1, 2)
        k = F.linear(h, P[p + "attention.k_lin.weight"], P[p + "attention.k_lin.bias"]).view(R, L, H, dh).transpose(1, 2)
        v = F.linear(h, P[p + "attention.v_lin.weight"], P[p + "attention.v_lin.bias"]).view(R, L, H, dh).transpose(1, 2)
        w = torch.matmul(q, k.transpose(2, 3)) * (dh ** -0.5) + add_mask
        w = _drop(torch.softmax(w, dim=-1), hp["ATTENTION_DROPOUT"], train)
        a = torch.matmul(w, v).transpose(1, 2).reshape(R, L, D)
        a = F.linear(a, P[p + "attention.out_lin.weight"], P[p + "attention.out_lin.bias"])
        h = _ln(a + h, P[p + "sa_layer_norm.weight"], P[p + "sa_layer_norm.bias"])
        f = F.gelu(F.linear(h, P[p + "ffn.lin1.weight"], P[p + "ffn.lin1.bias"]))  # exact-erf GELU (HF activations.py:83)
        f = _drop(F.linear(f, P[p + "ffn.lin2.weight"], P[p + "ffn.lin2.bias"]), hp["DROPOUT"], train)
        h = _ln(f + h, P[p + "output_layer_norm.weight"], P[p + "output_layer_norm.bias"])
    h = F.gelu(F.linear(h, P["model.vocab_transform.weight"], P["model.vocab_transform.bias"]))
    return _ln(h, P["model.vocab_layer_norm.weight"], P["model.vocab_layer_norm.bias"])


def model_forward(P: Dict[str, Tensor], x: Tensor, image_clip: Tensor, text_clip: Tensor, mask: Tensor, concat_mask: Tensor,
                  hp: dict, train: bool = False) -> Tuple[Tensor, Tensor]:
    """DistilBertModel.forward, CLIP-DDPM.py:271-323. Returns (vocab_out, feature_out)."""
    R = x.shape[0]
    ML = hp["MAX_LENGTH"]
    assert x.shape == (R, ML, hp["IN_CHANNEL"])  # :284-287
    assert image_clip.shape == text_clip.shape == (R, 1, hp["CLIP_DIM"])
    assert mask.shape == (R, ML)
    assert concat_mask.shape == (R, 2)
    guidance = concat_mask[:, 1] == 1  # :290
    if hp["TRAIN_EMBEDDING"]:  # :292-293
        x = F.linear(x, P["input_projection.weight"], P["input_projection.bias"])
    img = F.linear(image_clip, P["image_linear.weight"], P["image_linear.bias"])
    txt = F.linear(text_clip, P["text_linear.weight"], P["text_linear.bias"])
    if hp["CLIP_ADDING_METHOD"] == "concat":  # :295-302
        ones = torch.ones(R, 1, dtype=mask.dtype, device=mask.device)
        guided_mask = torch.hstack([mask, ones, ones])
        non_mask = torch.hstack([mask, ones, torch.zeros_like(ones)])
        xx = torch.hstack([x, img, txt])
        seg_idx = torch.tensor([0] * ML + [1] * 2, device=x.device)
        xx = xx + P["segment_embedding.weight"][seg_idx]
        guided_x = non_x = xx
    elif hp["CLIP_ADDING_METHOD"] == "add":  # :303-307
        guided_mask = non_mask = mask
        non_x = x + img
        guided_x = non_x + txt
    else:
        raise NotImplementedError(hp["CLIP_ADDING_METHOD"])
    x_out = encoder(P, non_x, non_mask, hp, train)  # :312
    w = hp["CLASSIFIER_FREE_WEIGHT"]
    if w > 0 and not guidance.sum() == 0:  # :313-317
        x_out = x_out.clone()
        x_out[guidance] = (1 + w) * encoder(P, guided_x[guidance], guided_mask[guidance], hp, train) - w * x_out[guidance]
    if hp["TRAIN_EMBEDDING"]:  # :319-320
        x_out = F.linear(x_out, P["output_projection.weight"], P["output_projection.bias"])
    assert x_out.shape == (R, non_mask.shape[-1], hp["IN_CHANNEL"])  # :322
    if hp["TRAIN_EMBEDDING"]:
        return F.linear(x_out[:, :ML, :], P["lm_head.weight"]), x_out  # trainable nn.Linear(IN_CHANNEL, VOCAB_SIZE, bias=False) (:239)
    return F.linear(x_out[:, :ML, :], P["embedding.weight"]), x_out  # lm_head: frozen, weight == embedding, bias 0 (:246-247,323)


# ------------------------------------------------------------------------------------------------------------------
# loss (CLIP-DDPM.py:382-445) and train_func (:458-486)
# ------------------------------------------------------------------------------------------------------------------
def loss(P, x_t, x_1, x_tgt, x_0, image_clip, text_clip, mask, idx, hp, train=False, classifier_mask: Optional[Tensor] = None):
    S, B, ML = hp["SAMPLE_SIZE"], hp["BATCH_SIZE"], hp["MAX_LENGTH"]
    assert x_t.shape == (S * B, ML, hp["IN_CHANNEL"])  # :396-400
    assert x_1.shape == x_0.shape == (B, ML, hp["IN_CHANNEL"])
    assert image_clip.shape == text_clip.shape == (B, hp["CLIP_DIM"])
    assert mask.shape == (B, ML) and idx.shape == (B, ML)
    loss_func = LOSS_FUNCS[hp["LOSS_FUNC"]]
    repeat_shape = (S, *(1,) * (len(x_t.shape) - 1))
    image_clip = image_clip.unsqueeze(1)
    text_clip = text_clip.unsqueeze(1)
    dev = x_t.device
    if hp["CLASSIFIER_FREE_WEIGHT"] > 0:  # :406-410
        if classifier_mask is None:
            classifier_mask = (torch.rand((S * B, 1)) > hp["CLASSIFIER_FREE_PROB"]).type(torch.float32).to(dev)
            classifier_mask[0] = 0
            classifier_mask[1] = 1
        concat_mask = torch.hstack([torch.ones((S * B, 1), device=dev), classifier_mask])
    else:
        concat_mask = torch.tensor([1, 0], device=dev).repeat((S * B, 1))  # :412
    x_t_prob, x_t_hidden = model_forward(P, x_t, image_clip.repeat(repeat_shape), text_clip.repeat(repeat_shape),
                                         mask.repeat((S, 1)), concat_mask, hp, train)  # :415
    if hp["USE_X_T_LOSS"]:
        if hp["X_0_PREDICTION"]:
            x_t_loss = loss_func(x_t_hidden[:, :ML, :], x_0.repeat(repeat_shape), hp)  # :418
        else:
            x_t_loss = loss_func(x_t_hidden[:, :ML, :], x_tgt, hp)  # :421
    else:
        x_t_loss = torch.zeros((), device=dev)
    x_1_prob, x_1_hidden = model_forward(P, x_1, image_clip, text_clip, mask, torch.tensor([1, 0], device=dev).repeat((B, 1)),
                                         hp, train)  # :426
    x_1_loss = loss_func(x_1_hidden[:, :ML, :], x_0, hp) if hp["USE_X_1_LOSS"] else torch.zeros((), device=dev)  # :428
    if hp["USE_PROB_LOSS"]:  # :432-440.  log(softmax(.)) written as log_softmax: identical unless the reference underflows to -inf
        idx = idx.unsqueeze(dim=-1)
        lt = -F.log_softmax(x_t_prob, dim=-1).gather(-1, idx.repeat(repeat_shape))
        l1 = -F.log_softmax(x_1_prob, dim=-1).gather(-1, idx)
        if hp["LOSS_FUNC"] in ("series_sum_sample_mean", "mse_series_mean"):
            x_t_prob_loss, x_1_prob_loss = lt.sum(dim=1).mean(), l1.sum(dim=1).mean()
        else:
            x_t_prob_loss, x_1_prob_loss = lt.sum() / B, l1.sum() / B
    else:
        x_t_prob_loss = x_1_prob_loss = torch.zeros((), device=dev)
    return x_t_loss, x_1_loss, hp["ROUNDING_WEIGHT"] * (x_t_prob_loss + x_1_prob_loss)  # :445


class AdamW:
    """torch.optim.AdamW(params, lr) restated (defaults betas (0.9, 0.999), eps 1e-8, weight_decay 0.01, one param group:
    decay applies to every tensor incl. biases / LayerNorm; CLIP-DDPM.py:335,484). A None grad is skipped (torch semantics);
    an all-zero grad still decays."""

    def __init__(self, params: List[Tensor], lr: float, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.params, self.lr, self.betas, self.eps, self.wd = params, lr, betas, eps, weight_decay
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        self.t += 1
        b1, b2 = self.betas
        for p, m, v in zip(self.params, self.m, self.v):
            if p.grad is None:
                continue
            g = p.grad
            p.mul_(1 - self.lr * self.wd)
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-self.lr / bc1)


def train_func(P: Dict[str, Tensor], trainer: Optional[AdamW], x: dict, hp: dict, acp: Tensor, train: bool = True, *,
               t: Optional[Tensor] = None, noise_t: Optional[Tensor] = None, noise_1: Optional[Tensor] = None,
               dropout: bool = True):
    """CLIP-DDPM.py:458-486. Extra keyword-only inputs pin the random draws (t, the two noise tensors) for parity tests."""
    S = hp["SAMPLE_SIZE"]
    x_0 = F.embedding(x["input_ids"], P["embedding.weight"])  # :459
    repeat_shape = (S, *(1,) * (len(x_0.shape) - 1))
    if t is None:
        t = torch.randint(0, hp["STEP_TOT"], repeat_shape, device=x_0.device)  # :461
    if hp["X_0_PREDICTION"]:
        x_t = diffuse_t(x_0, t, acp, noise_t)  # :464
        x_tgt = None
    else:
        t_next = torch.max(t - hp["X_T_STEP_INTERVAL"], torch.zeros(t.shape, device=x_0.device, dtype=torch.int64))
        x_t, x_tgt = diffuse_t(x_0, t, acp, noise_t), diffuse_t(x_0, t_next, acp)  # :467 via generate_diffuse_pair :364-380
    x_1 = diffuse_t(x_0, torch.ones(1, dtype=torch.int64, device=x_0.device), acp, noise_1)  # :468
    if train:
        trainer.zero_grad()
    x_t_loss, x_1_loss, prob_loss = loss(P, x_t, x_1, x_tgt, x_0, x["image_clip"], x["text_clip"], x["attention_mask"],
                                         x["input_ids"], hp, train=train and dropout)
    l = x_t_loss + x_1_loss + prob_loss  # :481
    if train:
        l.backward()
        trainer.step()
    return l, x_t_loss, x_1_loss, prob_loss


def make_trainable(P: Dict[str, Tensor], hp: dict) -> List[Tensor]:
    """requires_grad on exactly the tensors the reference's overridden parameters() returns (CLIP-DDPM.py:258-269)."""
    out = []
    for n in trainable_names(hp):
        P[n].requires_grad_(True)
        out.append(P[n])
    if not hp["TRAIN_EMBEDDING"]:
        P["embedding.weight"].requires_grad_(False)
    return out


# ------------------------------------------------------------------------------------------------------------------
# Denoise / sampling loop (CLIP-DDPM.py:611-621, COCO_BLEU.py:249-257)
# ------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def sample(P: Dict[str, Tensor], image_clip: Tensor, hp: dict, n_steps: int = 5, restored: Optional[Tensor] = None,
           return_all: bool = False):
    B = image_clip.shape[0]
    ML = hp["MAX_LENGTH"]
    L = ML + 2 if hp["CLIP_ADDING_METHOD"] == "concat" else ML
    dev = image_clip.device
    if restored is None:
        restored = torch.randn((B, L, hp["IN_CHANNEL"]), device=dev)  # :613
    outs = []
    out = None
    for _ in range(n_steps):  # :616-617
        out, restored = model_forward(P, restored[:, :ML, :], image_clip.unsqueeze(1), torch.zeros_like(image_clip).unsqueeze(1),
                                      torch.ones((B, ML), device=dev), torch.tensor([1, 0], device=dev).repeat(B, 1), hp, False)
        if return_all:
            outs.append(out.argmax(dim=-1))
    indexes = torch.softmax(out, dim=-1).argmax(dim=-1)  # :620
    return (indexes, restored, outs) if return_all else (indexes, restored)


def _s64(c: int) -> int:
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr(x: Tensor, s: int) -> Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def hash_uniform(n: int, k: int) -> Tensor:
    """n pseudo-random float64 values, uniform with zero mean and unit variance, from the splitmix64 integer hash of
    (index, stream k). Pure int64 arithmetic (wrapping multiply): bit-identical on every platform and torch version."""
    x = torch.arange(n, dtype=torch.int64) + _s64(((k + 1) * 0x9E3779B97F4A7C15) & ((1 << 64) - 1))
    x = (x ^ _lsr(x, 30)) * _s64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _s64(0x94D049BB133111EB)
    x = x ^ _lsr(x, 31)
    u = _lsr(x, 11).double() / float(1 << 53)
    return (u - 0.5) * math.sqrt(12.0)


def closed_form_tensor(shape, k: int, scale: float = 1.0) -> Tensor:
    """Generator-free pseudo-random tensor (integer hash): identical on every platform / torch version.
    Used by the golden fixtures for noise, CLIP features and `restored`."""
    return (scale * hash_uniform(int(math.prod(shape)), 7777 + k)).reshape(shape).float()


def closed_form_batch(hp: dict, k: int = 0, ragged: bool = True) -> dict:
    B, ML, V = hp["BATCH_SIZE"], hp["MAX_LENGTH"], hp["VOCAB_SIZE"]
    ids = ((torch.arange(B * ML, dtype=torch.int64) * 7919 + 104729 * (k + 1)) % V).reshape(B, ML)
    mask = torch.ones(B, ML, dtype=torch.int64)
    if ragged:
        lens = 6 + (torch.arange(B) * 5 + k) % (ML - 5)
        mask = (torch.arange(ML)[None, :] < lens[:, None]).to(torch.int64)
    img = F.normalize(closed_form_tensor((B, hp["CLIP_DIM"]), 11 + k), dim=-1)
    txt = F.normalize(closed_form_tensor((B, hp["CLIP_DIM"]), 23 + k), dim=-1)
    return {"input_ids": ids, "attention_mask": mask, "image_clip": img, "text_clip": txt}


def synthetic_batch(hp: dict, seed: int = 0, ragged: bool = False, device="cpu") -> dict:
    """Synthetic inputs of SURVEY §8(d): random ids, unit-norm random CLIP features, all-ones (or ragged) attention mask."""
    g = torch.Generator().manual_seed(seed)
    B, ML = hp["BATCH_SIZE"], hp["MAX_LENGTH"]
    ids = torch.randint(0, hp["VOCAB_SIZE"], (B, ML), generator=g)
    mask = torch.ones(B, ML, dtype=torch.int64)
    if ragged:
        lens = torch.randint(6, ML + 1, (B,), generator=g)
        mask = (torch.arange(ML)[None, :] < lens[:, None]).to(torch.int64)
    img = F.normalize(torch.randn(B, hp["CLIP_DIM"], generator=g), dim=-1)
    txt = F.normalize(torch.randn(B, hp["CLIP_DIM"], generator=g), dim=-1)
    return {"input_ids": ids.to(device), "attention_mask": mask.to(device), "image_clip": img.to(device), "text_clip": txt.to(device)}
