"""-m gpu: the reference's functions driven POSITIONALLY, the way CLIP-DDPM.py:458-486, :495, :551-552, :570-594 drive them."""
import numpy as np
import pytest
import torch

from _util import O, golden_hp, rel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import clipdlm
    assert torch.cuda.is_available(), "GPU tests selected on a box without CUDA"
    return clipdlm


def _model(pkg, hp, P, precision="bf16x3"):
    cfg = pkg.DistilBertConfig(n_layers=hp["N_LAYERS"], dim=hp["DIM"], n_heads=hp["N_HEADS"], hidden_dim=hp["HIDDEN_DIM"], dropout=hp["DROPOUT"],
                               attention_dropout=hp["ATTENTION_DROPOUT"])
    m = pkg.DistilBertModel(P["embedding.weight"], P["embedding.weight"], cfg, hp=hp, precision=precision)
    m.load_state_dict({k: v.detach() for k, v in P.items()})
    return m


def test_reference_call_forms_drive_a_train_step(pkg, tmp_path):
    """train_func's body (:458-486), validate(model) (:488-501), the checkpoint lines (:551-552, :570) and the demo (:584-594), verbatim."""
    hp = golden_hp(X_0_PREDICTION=False, LOSS_FUNC="mse_series_mean", BATCH_SIZE=3, SAMPLE_SIZE=4)
    P = O.init_params(hp, seed=0, closed_form=True)
    model = _model(pkg, hp, P)
    assert isinstance(model, torch.nn.Module)
    x = {k: v.to(DEV) for k, v in O.closed_form_batch(hp, 2).items()}
    val_loader = [x, x]
    F = pkg.bind(hp, val_loader)
    try:
        trainer = pkg.AdamW(model.parameters(), lr=1e-4)
        device = torch.device(DEV)
        SAMPLE_SIZE, STEP_TOT, X_T_STEP_INTERVAL, MAX_LENGTH = hp["SAMPLE_SIZE"], hp["STEP_TOT"], hp["X_T_STEP_INTERVAL"], hp["MAX_LENGTH"]
        # ---- :459-468
        x_0 = model.embedding(x["input_ids"])
        repeat_shape = (SAMPLE_SIZE, *(1, ) * (len(x_0.shape) - 1))
        t = torch.randint(0, STEP_TOT, repeat_shape, device=device)
        x_t, x_tgt = F.generate_diffuse_pair(x_0, t, torch.max(t - X_T_STEP_INTERVAL, torch.zeros(t.shape, device=device, dtype=torch.int64)))
        x_1 = F.diffuse_t(x_0, torch.ones(1, dtype=torch.int64, device=device))
        assert tuple(x_t.shape) == tuple(x_tgt.shape) == (SAMPLE_SIZE * 3, MAX_LENGTH, 768) and tuple(x_1.shape) == (3, MAX_LENGTH, 768)
        # ---- :470-484
        trainer.zero_grad()
        x_t_loss, x_1_loss, prob_loss = F.loss(model, x_t, x_1, x_tgt, x_0, x["image_clip"], x["text_clip"], x["attention_mask"], x["input_ids"],
                                               hp["LOSS_FUNC"])
        ref = O.loss(P, x_t.cpu(), x_1.cpu(), x_tgt.cpu(), x_0.cpu(), *(x[k].cpu() for k in ("image_clip", "text_clip", "attention_mask", "input_ids")), hp)
        for a, b in zip((x_t_loss, x_1_loss, prob_loss), ref):
            assert abs(a.item() - b.item()) < 1e-3 * abs(b.item())
        trainer.step()
        # ---- the loop's own calls (:526, :546)
        l, a, b, c = F.train_func(model, trainer, x)
        assert abs(l.item() - (a + b + c).item()) < 1e-4 * abs(l.item())
        val_x_t, val_x_1, val_prob = F.validate(model)
        assert all(torch.isfinite(v) for v in (val_x_t, val_x_1, val_prob)) and model.training
        # ---- :551-552 / :570: whole-module pickle round trip
        path = str(tmp_path / "model.pickle")
        torch.save(model.cpu(), path)
        model = model.to(device)
        back = torch.load(path, weights_only=False).to(device)
        assert type(back) is type(model) and back is not model
        for (n1, p1), (n2, p2) in zip(model.named_parameters(), back.named_parameters()):
            assert n1 == n2 and torch.equal(p1, p2), n1
        # ---- :584-594 demo: one diffusion at t = 999, a few denoise calls with the reference's argument list
        model.eval(); back.eval()
        with torch.no_grad():
            x_t = F.diffuse_t(x_0[:1], torch.tensor([999], dtype=torch.int64, device=device))
            mask = x["attention_mask"][:1]
            restored = x_t
            for _ in range(2):
                out, restored = model(restored[:, :MAX_LENGTH, :], x["image_clip"][:1, None, :], x["text_clip"][:1, None, :], mask,
                                      torch.tensor([1, 0], device=device).repeat(mask.shape[0], 1))
            out_b, restored_b = back(x_t[:, :MAX_LENGTH, :], x["image_clip"][:1, None, :], x["text_clip"][:1, None, :], mask,
                                     torch.tensor([1, 0], device=device).repeat(mask.shape[0], 1))
            out_a, _ = model(x_t[:, :MAX_LENGTH, :], x["image_clip"][:1, None, :], x["text_clip"][:1, None, :], mask,
                             torch.tensor([1, 0], device=device).repeat(mask.shape[0], 1))
        assert tuple(out.shape) == (1, MAX_LENGTH, hp["VOCAB_SIZE"]) and tuple(restored.shape) == (1, MAX_LENGTH + 2, 768)
        assert torch.equal(out_a, out_b)
    finally:
        pkg.set_globals(pkg.default_hparams())
        from clipdlm import hparams
        hparams.ACTIVE["val_loader"] = None


def test_diffuse_t_and_generate_diffuse_pair(pkg):
    """a2 / a3: reference call forms, against the schedule. t = 0 returns x exactly (alpha_bar[0] = 1); both halves of the pair at t = t_next = 0
    equal x_0; x_0-prediction returns x_0 itself as target; at t = 999 the sample is (almost) pure unit noise; the n samples share one draw."""
    hp = pkg.default_hparams(BATCH_SIZE=4, SAMPLE_SIZE=3)
    x_0 = torch.randn(4, 16, 768, device=DEV) * 0.05
    zero = torch.zeros(3, 1, 1, dtype=torch.int64, device=DEV)
    F = pkg.bind(hp)
    try:
        a, tgt = F.generate_diffuse_pair(x_0, zero)
        assert tgt is x_0 and torch.equal(a, x_0.repeat(3, 1, 1))
        F.hp["X_0_PREDICTION"] = False   # F.hp is the active dict the functions read (a module global in the reference)
        a, b = F.generate_diffuse_pair(x_0, zero, zero)
        assert torch.equal(a, x_0.repeat(3, 1, 1)) and torch.equal(b, a)
        t = torch.tensor([999, 500, 999], device=DEV).reshape(3, 1, 1)
        a, b = F.generate_diffuse_pair(x_0, t, torch.max(t - 100, torch.zeros_like(t)))
        assert tuple(a.shape) == tuple(b.shape) == (12, 16, 768)
        acp = O.alpha_cumprod(hp)
        assert abs(float(a[:4].std()) - float((1 - acp[999]).sqrt())) < 2e-2
        assert torch.equal(a[:4], a[8:])                       # same t, same (shared) noise draw -> identical samples (:359)
        eps = (a[:4] - acp[999].sqrt().item() * x_0) / (1 - acp[999]).sqrt().item()
        mid = acp[500].sqrt().item() * x_0 + (1 - acp[500]).sqrt().item() * eps
        assert rel(a[4:8], mid) < 1e-5
        assert not torch.equal(a[:4], b[:4])                   # the target is an independent draw (:380)
    finally:
        pkg.set_globals(pkg.default_hparams())


@pytest.mark.parametrize("fusion", ["concat", "add"])
def test_forward_classifier_free_branch_vs_oracle(pkg, fusion):
    """forward()'s CFG branch (:313-317): rows with concat_mask[:, 1] == 1 get (1 + w) guided - w unguided, the others stay unguided."""
    hp = golden_hp(CLASSIFIER_FREE_WEIGHT=0.3, CLIP_ADDING_METHOD=fusion)
    P = O.init_params(hp, seed=0, closed_form=True)
    model = _model(pkg, hp, P).eval()
    R = 5
    x = O.closed_form_tensor((R, 16, 768), 3, 0.5)
    img, txt = O.closed_form_tensor((R, 1, 512), 4, 0.05), O.closed_form_tensor((R, 1, 512), 5, 0.05)
    mask = torch.ones(R, 16, dtype=torch.int64); mask[1, 9:] = 0; mask[3, 5:] = 0
    cm = torch.tensor([[1, 0], [1, 1], [1, 1], [1, 0], [1, 1]])
    with torch.no_grad():
        ref_logits, ref_x = O.model_forward(P, x, img, txt, mask, cm, hp, False)
    logits, x_out = model(x.to(DEV), img.to(DEV), txt.to(DEV), mask.to(DEV), cm.to(DEV))
    assert rel(x_out, ref_x) < 1e-3 and rel(logits, ref_logits) < 1e-3
    assert np.array_equal(logits.argmax(-1).cpu().numpy(), ref_logits.argmax(-1).numpy())
    plain, _ = model(x.to(DEV), img.to(DEV), txt.to(DEV), mask.to(DEV), torch.tensor([[1, 0]], device=DEV).repeat(R, 1))
    # (unguided rows go through the dense lm_head GEMM in the mixed call, through the gathered one otherwise: equal to rounding)
    assert rel(plain[[0, 3]], logits[[0, 3]]) < 1e-5 and rel(plain[[1, 2, 4]], logits[[1, 2, 4]]) > 1e-3
