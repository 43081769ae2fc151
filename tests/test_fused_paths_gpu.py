"""-m gpu: the two fused paths that became the bf16 default in round 2 (validated on B200: profiles/r02_validate_fused_paths.log):
  * the factored softmax-CE gradient of the lm_head (clipdlm.h CLIPDLM_EPI_LSE_EXP / CLIPDLM_EPI_STORE_ROWSCALE / clipdlm_ce_row_terms /
    CLIPDLM_OPT_FUSED_SOFTMAX_GRAD) against fp64 torch and against the in-place softmax-gradient path it replaces,
  * gelu'(u) stored by lin1's epilogue + the multiplying lin2 gradient GEMM (+ lin1's bias gradient summed in that epilogue)
    (CLIPDLM_EPI_STORE_GELU_DERIV / CLIPDLM_EPI_STORE_MULAUX / CLIPDLM_OPT_GELU_DERIV_STORE)."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

from _util import O, rel

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def G():
    import _gpu
    assert torch.cuda.is_available(), "GPU tests selected on a box without CUDA"
    torch.manual_seed(0)
    return _gpu


@pytest.mark.parametrize("shift", [0.0, 3.5])
@pytest.mark.parametrize("R", [1, 9, 300])
def test_factored_softmax_gradient_kernels(G, R, shift):
    """LSE_EXP GEMM -> lse_combine -> ce_row_terms -> STORE_ROWSCALE GEMM against fp64 torch, at the lm_head geometry (gathered x_out rows,
    vocabulary tail 30522 = 119 * 256 + 58, scattered output rows)."""
    Ltxt, Lf, D, V = 16, 18, 768, 30522
    M = R * Ltxt
    xo = torch.randn(R * Lf, D, device=DEV)
    E = torch.randn(V, D, device=DEV) * 0.05
    Epad = torch.zeros((V + 255) // 256 * 256, D, device=DEV); Epad[:V] = E
    xh, eh = xo.to(torch.bfloat16), Epad.to(torch.bfloat16)
    B = max(1, R // 3) if R % 3 == 0 else R
    tgt = torch.randint(0, V, (B * Ltxt,), device=DEV, dtype=torch.int32)
    nt = (V + 255) // 256
    ldl = nt * 256
    pm = torch.zeros(2 * nt, M, device=DEV); ps = torch.zeros(2 * nt, M, device=DEV)
    tl = torch.zeros(M, device=DEV)
    ex = torch.full((M, ldl), float("nan"), device=DEV, dtype=torch.bfloat16)
    sh = torch.tensor([shift], device=DEV)
    G.gemm(a_hi=xh, b_hi=eh, lda=D, ldb=D, M=M, N=V, K=D, gather_len=Ltxt, gather_stride=Lf, epilogue=G.L.EPI_LSE_EXP, part_max=pm, part_sum=ps,
           tgt_logit=tl, targets=tgt, tgt_period=B * Ltxt, out_hi=ex, ldo=ldl, exp_shift=sh)
    xg = xh.float().view(R, Lf, D)[:, :Ltxt].reshape(M, D).double()
    W = eh.float()[:V].double()
    logits = xg @ W.t()
    assert rel(ex[:, :V], torch.exp(logits - shift)) < 6e-3                      # bf16 storage of exp(s - c)
    assert float(pm.min()) == float(pm.max()) == pytest.approx(shift)
    lse = torch.zeros(M, device=DEV); acc = torch.zeros(1, device=DEV, dtype=torch.float64)
    G.L.check(G.lib().clipdlm_lse_combine(pm.data_ptr(), ps.data_ptr(), None, 2 * nt, M, tl.data_ptr(), lse.data_ptr(), None, acc.data_ptr(), 1.0 / R,
                                          G.st()))
    tfull = tgt.long().repeat(M // (B * Ltxt))
    ref_lse = torch.logsumexp(logits, -1)
    assert rel(lse, ref_lse) < 1e-4
    ref_loss = (ref_lse - logits.gather(1, tfull[:, None])[:, 0]).sum() / R
    assert abs(acc.item() - ref_loss.item()) < 2e-3 * abs(ref_loss.item())
    # row terms: rs = scale * exp(c - lse); one-hot term folded into the (scattered) rows of d(x_out)
    scale = 0.37
    g0 = torch.randn(R * Lf, D, device=DEV).to(torch.bfloat16)
    before = g0.float().clone()
    rs = torch.zeros(M, device=DEV)
    G.L.check(G.lib().clipdlm_ce_row_terms(lse.data_ptr(), sh.data_ptr(), tgt.data_ptr(), B * Ltxt, scale, M, eh.data_ptr(), D, g0.data_ptr(), D, Ltxt, Lf, D,
                                           rs.data_ptr(), G.st()))
    torch.cuda.synchronize()
    assert rel(rs, scale * torch.exp(shift - ref_lse)) < 1e-4
    ref_mid = before.view(R, Lf, D).double().clone()
    ref_mid[:, :Ltxt] -= scale * W[tfull].view(R, Ltxt, D)
    assert rel(g0.float(), ref_mid.view(R * Lf, D)) < 4e-3
    assert torch.equal(g0.float().view(R, Lf, D)[:, Ltxt:], before.view(R, Lf, D)[:, Ltxt:])   # CLIP rows untouched
    # gradient GEMM: d x_out[:, :Ltxt] = rs[m] * (exp(s - c) @ W) + (d L1 - scale W[tgt])
    mid = g0.float().clone()
    G.gemm(a_hi=ex, b_hi=eh, lda=ldl, ldb=D, M=M, N=D, K=V, a_major=0, b_major=1, epilogue=G.L.EPI_STORE_ROWSCALE, out_hi=g0, ldo=D, res_hi=g0, ldr=D,
           scatter_len=Ltxt, scatter_stride=Lf, row_scale=rs)
    ref_g = before.view(R, Lf, D).double().clone()
    ref_g[:, :Ltxt] += (((torch.softmax(logits, -1) - F.one_hot(tfull, V).double()) * scale) @ W).view(R, Ltxt, D)
    assert rel(g0.float(), ref_g.view(R * Lf, D)) < 6e-3
    assert torch.equal(g0.float().view(R, Lf, D)[:, Ltxt:], mid.view(R, Lf, D)[:, Ltxt:])
    # the same GEMM without scatter takes the TMA-store side of the epilogue
    dense = torch.zeros(M, D, device=DEV, dtype=torch.bfloat16)
    res = torch.randn(M, D, device=DEV).to(torch.bfloat16)
    G.gemm(a_hi=ex, b_hi=eh, lda=ldl, ldb=D, M=M, N=D, K=V, a_major=0, b_major=1, epilogue=G.L.EPI_STORE_ROWSCALE, out_hi=dense, ldo=D, res_hi=res, ldr=D,
           row_scale=rs)
    ref_d = res.float().double() + rs.double()[:, None] * (ex[:, :V].float().double() @ W)
    assert rel(dense.float(), ref_d) < 5e-3


def test_rowscale_rejects_unsupported_operands(G):
    a = torch.zeros(256, 256, device=DEV, dtype=torch.bfloat16)
    rs = torch.ones(256, device=DEV)
    with pytest.raises(G.L.ClipdlmError):   # no residual
        G.gemm(a_hi=a, b_hi=a, lda=256, ldb=256, M=256, N=256, K=256, epilogue=G.L.EPI_STORE_ROWSCALE, out_hi=a, ldo=256, row_scale=rs)
    with pytest.raises(G.L.ClipdlmError):   # LSE_EXP without the output array
        G.gemm(a_hi=a, b_hi=a, lda=256, ldb=256, M=256, N=256, K=256, epilogue=G.L.EPI_LSE_EXP, part_max=rs, part_sum=rs)


@pytest.mark.parametrize("fusion", ["concat", "add"])
def test_train_step_gradients_match_the_default_path(fusion):
    """Same weights, same draws, dropout off: the factored path must reproduce the default bf16 path's losses (same LSE up to the bf16
    rounding of exp(s) not entering it) and gradients (direction and norm), and both must sit at the same distance from the fp32 oracle."""
    import clipdlm as pkg
    hp = pkg.default_hparams(BATCH_SIZE=6, SAMPLE_SIZE=5, N_LAYERS=2, DROPOUT=0.0, ATTENTION_DROPOUT=0.0, CLIP_ADDING_METHOD=fusion)
    g = torch.Generator().manual_seed(3)
    batch = {"input_ids": torch.randint(0, 30522, (6, 16), generator=g).to(DEV), "attention_mask": torch.ones(6, 16, dtype=torch.int64, device=DEV),
             "image_clip": F.normalize(torch.randn(6, 512, generator=g), dim=-1).to(DEV), "text_clip": F.normalize(torch.randn(6, 512, generator=g), dim=-1).to(DEV)}
    batch["attention_mask"][1, 9:] = 0
    t = torch.tensor([0, 17, 400, 999, 250]).reshape(5, 1, 1)
    n_t, n_1 = torch.randn(6, 16, 768, generator=g), torch.randn(6, 16, 768, generator=g)
    outs = {}
    for flag in (False, True):
        model = pkg.DistilBertModel(None, None, None, hp=hp, precision="bf16", seed=0, chunk_rows=12, fused_softmax_grad=flag).train()
        assert model.fused_softmax_grad is flag
        trainer = pkg.AdamW(model.parameters(), lr=1e-4)
        snap = {}
        trainer.step = lambda m=model, s=snap: s.update(g=m.grad.clone())
        losses = pkg.train_func(model, trainer, batch, t=t, noise_t=n_t, noise_1=n_1, dropout_seed=1)
        outs[flag] = ([x.item() for x in losses], snap["g"].double())
    (l0, g0), (l1, g1) = outs[False], outs[True]
    for a, b in zip(l0, l1):
        assert abs(a - b) < 2e-4 * abs(a), (l0, l1)
    cos = float((g0 * g1).sum() / (g0.norm() * g1.norm()))
    assert cos > 0.999 and abs(float(g1.norm() / g0.norm()) - 1) < 1e-2, (cos, float(g1.norm() / g0.norm()))


def test_exp_shift_bound_and_large_logits():
    """refresh_exp_shift(): c = clamp(bound - 69, 0, 60). Scaling the lm_head weight up pushes the Cauchy-Schwarz bound past 69: c > 0, the
    stored exp(s - c) stay finite and the loss still matches the default path."""
    import clipdlm as pkg
    hp = pkg.default_hparams(BATCH_SIZE=4, SAMPLE_SIZE=3, N_LAYERS=2, DROPOUT=0.0, ATTENTION_DROPOUT=0.0)
    g = torch.Generator().manual_seed(5)
    E = torch.randn(30522, 768, generator=g) * 0.02
    batch = {"input_ids": torch.randint(0, 30522, (4, 16), generator=g).to(DEV), "attention_mask": torch.ones(4, 16, dtype=torch.int64, device=DEV),
             "image_clip": F.normalize(torch.randn(4, 512, generator=g), dim=-1).to(DEV), "text_clip": F.normalize(torch.randn(4, 512, generator=g), dim=-1).to(DEV)}
    t = torch.tensor([3, 250, 900]).reshape(3, 1, 1)
    n_t, n_1 = torch.randn(4, 16, 768, generator=g), torch.randn(4, 16, 768, generator=g)
    res = {}
    for flag in (False, True):
        model = pkg.DistilBertModel(E, E * 6.0, None, hp=hp, precision="bf16", seed=0, fused_softmax_grad=flag).train()   # logits x 6
        trainer = pkg.AdamW(model.parameters(), lr=1e-4)
        snap = {}
        trainer.step = lambda m=model, s=snap: s.update(g=m.grad.clone())
        losses = pkg.train_func(model, trainer, batch, t=t, noise_t=n_t, noise_1=n_1, dropout_seed=1)
        res[flag] = ([x.item() for x in losses], snap["g"].double(), model)
    m1 = res[True][2]
    bound = (768 ** 0.5 * 1.0 + 0.0) * float((E * 6.0).norm(dim=1).max())
    assert float(m1._exp_shift.item()) == pytest.approx(min(max(bound - 69.0, 0.0), 60.0), rel=1e-5) and float(m1._exp_shift.item()) > 0
    for a, b in zip(res[False][0], res[True][0]):
        assert abs(a - b) < 5e-4 * abs(a)
    g0, g1 = res[False][1], res[True][1]
    assert torch.isfinite(g1).all() and float((g0 * g1).sum() / (g0.norm() * g1.norm())) > 0.999


def test_logit_bound_out_of_range_falls_back_to_the_inplace_path():
    """a bound past 69 + 60 must not saturate silently. lm_head weight x 12 => Cauchy-Schwarz bound ~ 184: the model warns, switches
    the factored path off (also under classifier-free guidance, whose mixed x_out widens the bound by 1 + 2w) and gives the in-place path's numbers."""
    import clipdlm as pkg
    hp = pkg.default_hparams(BATCH_SIZE=4, SAMPLE_SIZE=3, N_LAYERS=2, DROPOUT=0.0, ATTENTION_DROPOUT=0.0)
    g = torch.Generator().manual_seed(5)
    E = torch.randn(30522, 768, generator=g) * 0.02
    batch = {"input_ids": torch.randint(0, 30522, (4, 16), generator=g).to(DEV), "attention_mask": torch.ones(4, 16, dtype=torch.int64, device=DEV),
             "image_clip": F.normalize(torch.randn(4, 512, generator=g), dim=-1).to(DEV), "text_clip": F.normalize(torch.randn(4, 512, generator=g), dim=-1).to(DEV)}
    t = torch.tensor([3, 250, 900]).reshape(3, 1, 1)
    n_t, n_1 = torch.randn(4, 16, 768, generator=g), torch.randn(4, 16, 768, generator=g)
    res = {}
    for flag in (False, True):
        model = pkg.DistilBertModel(E, E * 12.0, None, hp=hp, precision="bf16", seed=0, fused_softmax_grad=flag).train()
        trainer = pkg.AdamW(model.parameters(), lr=1e-4)
        snap = {}
        trainer.step = lambda m=model, s=snap: s.update(g=m.grad.clone())
        if flag:
            with pytest.warns(UserWarning, match="logit bound"):
                losses = pkg.train_func(model, trainer, batch, t=t, noise_t=n_t, noise_1=n_1, dropout_seed=1)
            assert model.fused_softmax_grad is False
        else:
            losses = pkg.train_func(model, trainer, batch, t=t, noise_t=n_t, noise_1=n_1, dropout_seed=1)
        res[flag] = ([x.item() for x in losses], snap["g"].double())
    assert res[False][0] == res[True][0]                                      # the same kernels ran: losses bit-equal,
    assert rel(res[True][1], res[False][1]) < 1e-4                              # gradients equal up to the fp32 atomics of the split-K weight gradients
    # classifier-free guidance widens the bound: x 4 alone stays inside (bound ~ 66 < 69, c = 0), with w = 0.25 (factor 1.5) c > 0
    hp2 = dict(hp, CLASSIFIER_FREE_WEIGHT=0.25)
    m_plain = pkg.DistilBertModel(E, E * 4.0, None, hp=hp, precision="bf16", seed=0, fused_softmax_grad=True)
    m_cfg = pkg.DistilBertModel(E, E * 4.0, None, hp=hp2, precision="bf16", seed=0, fused_softmax_grad=True)
    m_plain._engine(12, 4, True); m_cfg._engine(12, 4, True)
    assert float(m_plain._exp_shift.item()) == 0.0 and float(m_cfg._exp_shift.item()) > 0.0


# ------------------------------------------------------------------------------------------------ gelu'(u) stored by the forward
@pytest.mark.parametrize("M", [144, 4096 + 48])
def test_gelu_deriv_store_and_mulaux_kernels(G, M):
    """STORE_GELU_DERIV (out = gelu'(xW^T + b), out2 = gelu(xW^T + b)) and STORE_MULAUX (out = (dy W) * u) against fp64 torch."""
    K, N = 768, 3072
    x = (torch.randn(M, K, device=DEV) * 0.7).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) * 0.04).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV) * 0.3
    d = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16); g = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    G.gemm(a_hi=x, b_hi=w, lda=K, ldb=K, M=M, N=N, K=K, epilogue=G.L.EPI_STORE_GELU_DERIV, out_hi=d, out2_hi=g, ldo=N, bias=bias)
    u = x.double() @ w.double().t() + bias.double()
    Phi = 0.5 * (1 + torch.erf(u / 2 ** 0.5)); pdf = torch.exp(-0.5 * u * u) / (2 * torch.pi) ** 0.5
    assert rel(g.float(), u * Phi) < 4e-3
    assert rel(d.float(), Phi + u * pdf) < 4e-3
    assert float((d.float().double() - (Phi + u * pdf)).abs().max()) < 6e-3     # bf16 rounding of values up to 1.13
    # backward: dU = (dY W2) * gelu'(u), W2 stored [N_out = K][F = N] and read as the MN-major B operand
    dy = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16)
    w2 = (torch.randn(K, N, device=DEV) * 0.04).to(torch.bfloat16)
    du = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    G.gemm(a_hi=dy, b_hi=w2, lda=K, ldb=N, M=M, N=N, K=K, a_major=0, b_major=1, epilogue=G.L.EPI_STORE_MULAUX, out_hi=du, ldo=N, u_hi=d, ldu=N)
    assert rel(du.float(), (dy.double() @ w2.double()) * d.float().double()) < 4e-3
    # the same launch with the fused bias gradient: acc_f32[n] += column sums of the (bf16-rounded) output
    du2 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    cs = torch.full((N,), 0.25, device=DEV)
    G.gemm(a_hi=dy, b_hi=w2, lda=K, ldb=N, M=M, N=N, K=K, a_major=0, b_major=1, epilogue=G.L.EPI_STORE_MULAUX, out_hi=du2, ldo=N, u_hi=d, ldu=N, acc_f32=cs)
    assert torch.equal(du2, du)
    assert rel(cs, 0.25 + du.float().double().sum(0)) < 1e-5
    # and it is the same gradient the default epilogue forms from the stored pre-activation
    ub = u.float().to(torch.bfloat16)
    du0 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    G.gemm(a_hi=dy, b_hi=w2, lda=K, ldb=N, M=M, N=N, K=K, a_major=0, b_major=1, epilogue=G.L.EPI_STORE, out_hi=du0, ldo=N, u_hi=ub, ldu=N)
    assert rel(du.float(), du0.float()) < 8e-3


def test_gelu_deriv_epilogues_reject_unsupported_operands(G):
    a = torch.zeros(256, 256, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(G.L.ClipdlmError):   # no bias
        G.gemm(a_hi=a, b_hi=a, lda=256, ldb=256, M=256, N=256, K=256, epilogue=G.L.EPI_STORE_GELU_DERIV, out_hi=a, out2_hi=a, ldo=256)
    with pytest.raises(G.L.ClipdlmError):   # K-major B
        G.gemm(a_hi=a, b_hi=a, lda=256, ldb=256, M=256, N=256, K=256, epilogue=G.L.EPI_STORE_MULAUX, out_hi=a, ldo=256, u_hi=a, ldu=256)


@pytest.mark.parametrize("both,level", [(False, 1), (False, 2), (True, 2)])
def test_train_step_with_gelu_deriv_store_matches_the_default_path(both, level):
    """Same weights / draws, dropout ON (the masks are counter-based, so both runs drop the same elements): losses equal (the forward
    values do not change), gradients equal up to the bf16 rounding of gelu'(u). both=True also switches the factored softmax gradient on."""
    import clipdlm as pkg
    hp = pkg.default_hparams(BATCH_SIZE=6, SAMPLE_SIZE=5, N_LAYERS=2)
    g = torch.Generator().manual_seed(4)
    batch = {"input_ids": torch.randint(0, 30522, (6, 16), generator=g).to(DEV), "attention_mask": torch.ones(6, 16, dtype=torch.int64, device=DEV),
             "image_clip": F.normalize(torch.randn(6, 512, generator=g), dim=-1).to(DEV), "text_clip": F.normalize(torch.randn(6, 512, generator=g), dim=-1).to(DEV)}
    t = torch.tensor([0, 17, 400, 999, 250]).reshape(5, 1, 1)
    n_t, n_1 = torch.randn(6, 16, 768, generator=g), torch.randn(6, 16, 768, generator=g)
    outs = {}
    for flag in (False, True):
        model = pkg.DistilBertModel(None, None, None, hp=hp, precision="bf16", seed=0, chunk_rows=12, gelu_deriv_store=level if flag else 0,
                                    fused_softmax_grad=flag and both).train()
        trainer = pkg.AdamW(model.parameters(), lr=1e-4)
        snap = {}
        trainer.step = lambda m=model, s=snap: s.update(g=m.grad.clone())
        losses = pkg.train_func(model, trainer, batch, t=t, noise_t=n_t, noise_1=n_1, dropout_seed=7)
        outs[flag] = ([x.item() for x in losses], snap["g"].double())
    (l0, g0), (l1, g1) = outs[False], outs[True]
    for a, b in zip(l0, l1):
        assert abs(a - b) < (2e-4 if both else 1e-6) * abs(a), (l0, l1)
    cos = float((g0 * g1).sum() / (g0.norm() * g1.norm()))
    assert cos > 0.999 and abs(float(g1.norm() / g0.norm()) - 1) < 1e-2, (cos, float(g1.norm() / g0.norm()))
