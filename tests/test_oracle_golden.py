"""CPU: the oracle restatement (oracle/clipdlm_oracle.py) replayed against the golden fixtures that the REAL reference produced
(tests/golden/make_golden.py). This is what pins the oracle on machines where /root/reference does not exist."""
import numpy as np
import pytest
import torch

from _util import FULL_CASES, GOLDEN_CASES, O, golden_hp, golden_inputs, load_golden, rel


@pytest.fixture(scope="module", autouse=True)
def _threads():
    torch.set_num_threads(8)


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_train_step_matches_reference(name):
    hp = golden_hp(**GOLDEN_CASES[name])
    g = load_golden(name)
    inp = golden_inputs(hp)
    P = O.init_params(hp, seed=0, closed_form=True)
    trainable = O.make_trainable(P, hp)
    opt = O.AdamW(trainable, lr=1e-3)
    l, a, b, c = O.train_func(P, opt, inp["batch"], hp, O.alpha_cumprod(hp), True, t=inp["t"], noise_t=inp["noise_t"], noise_1=inp["noise_1"])
    got = np.array([l.item(), a.item(), b.item(), c.item()])
    np.testing.assert_allclose(got, g["train_losses"], rtol=2e-5)
    names = [str(n) for n in g["grad_names"]]
    gscale = float(g["grad_norms"].max())
    for n, ref_norm in zip(names, g["grad_norms"]):
        grad = P[n].grad
        assert grad is not None, n
        assert abs(float(grad.double().norm()) - ref_norm) <= 2e-4 * max(ref_norm, 1e-4 * gscale), n
        ref = torch.from_numpy(g["grad::" + n])
        mine = grad.reshape(-1)[:ref.numel()].reshape(ref.shape) if grad.numel() > ref.numel() else grad
        assert float((mine.double() - ref.double()).norm()) <= 2e-4 * max(float(ref.double().norm()), 1e-4 * gscale), n
    for n, ref_norm, gn in zip(names, g["after_norms"], g["grad_norms"]):
        if 0.0 < gn < 1e-4 * gscale:
            continue  # analytically-zero gradient (k_lin.bias: softmax is shift-invariant): Adam's first step is sign(noise)
        assert abs(float(P[n].detach().double().norm()) - ref_norm) <= 1e-5 * max(ref_norm, 1e-3), n


@pytest.mark.parametrize("name", FULL_CASES)
def test_forward_and_denoise_match_reference(name):
    hp = golden_hp(**GOLDEN_CASES[name])
    g = load_golden(name)
    inp = golden_inputs(hp)
    P = O.init_params(hp, seed=0, closed_form=True)
    R = inp["fwd_x"].shape[0]
    with torch.no_grad():
        logits, x_out = O.model_forward(P, inp["fwd_x"], inp["fwd_img"], inp["fwd_txt"], inp["fwd_mask"],
                                        torch.tensor([1, 0]).repeat(R, 1), hp, False)
    assert float((x_out - torch.from_numpy(g["fwd_x_out"])).abs().max()) < 2e-5
    assert np.array_equal(logits.argmax(-1).numpy(), g["fwd_argmax"])  # bit-exact token indices
    assert float((logits[:, :, :64] - torch.from_numpy(g["fwd_logits_head"])).abs().max()) < 2e-5
    ids, restored = O.sample(P, inp["batch"]["image_clip"], hp, 5, inp["restored"].clone())
    assert np.array_equal(ids.numpy(), g["sample_ids"])
    assert rel(restored, g["sample_restored"]) < 1e-5


def test_schedule_known_answers():
    """SURVEY App. C.1 known-answer values of the cosine schedule (fp32)."""
    acp = O.alpha_cumprod(O.default_hparams())
    assert acp[0].item() == 1.0
    np.testing.assert_allclose(acp[[1, 500, 999]].numpy(), [0.9999587536, 0.4938435555, 2.4289217890e-06], rtol=2e-6)
    x = torch.randn(2, 16, 768)
    out = O.diffuse_t(x, torch.tensor([0, 3]).reshape(2, 1, 1), acp, torch.randn(2, 16, 768))
    assert torch.equal(out[:2], x)  # t = 0 => x_t == x_0 exactly (App. E-3)


def test_structural_invariants():
    """App. C.1: text_linear gets an exactly-zero gradient yet still decays; exactly 17 position rows receive gradient."""
    hp = golden_hp()
    P = O.init_params(hp, seed=0, closed_form=True)
    before = P["text_linear.weight"].clone()
    trainable = O.make_trainable(P, hp)
    opt = O.AdamW(trainable, lr=1e-3)
    inp = golden_inputs(hp)
    O.train_func(P, opt, inp["batch"], hp, O.alpha_cumprod(hp), True, t=inp["t"], noise_t=inp["noise_t"], noise_1=inp["noise_1"])
    assert float(P["text_linear.weight"].grad.abs().max()) == 0.0
    np.testing.assert_allclose(P["text_linear.weight"].detach().numpy(), (before * (1 - 1e-3 * 0.01)).numpy(), rtol=1e-6)
    pg = P["model.distilbert.embeddings.position_embeddings.weight"].grad
    assert int((pg.abs().sum(dim=1) > 0).sum()) == hp["MAX_LENGTH"] + 1  # position 17 (masked text-CLIP slot) only reaches x_out[:, 17]


def test_real_width_oracle_matches_reference():
    """The reference's real dimensions (6 layers, V = 30522, B = 8, S = 100; CLIP-DDPM.py:57,109 = BASELINE.json configs[0]): the oracle
    replays the fixture the REAL reference produced (tests/golden/make_golden.py real) - 5-step denoise ids at every step, one train
    step's losses / gradient norms + slices / AdamW deltas. About 25 s on 8 cores."""
    from _util import real_width_inputs
    g = load_golden("real_width_6L")
    hp, inp = real_width_inputs(0)
    assert np.array_equal(inp["t"].reshape(-1).numpy(), g["t"])
    P = O.init_params(hp, seed=0, closed_form=True)
    ids, restored, outs = O.sample(P, inp["batch"]["image_clip"], hp, 5, inp["restored"].clone(), return_all=True) \
        if "return_all" in O.sample.__code__.co_varnames else (*O.sample(P, inp["batch"]["image_clip"], hp, 5, inp["restored"].clone()), None)
    assert np.array_equal(ids.numpy(), g["sample_ids_steps"][-1])
    assert rel(restored[:, :, ::16], g["sample_restored_slice"]) < 1e-5
    P0 = {k: v.clone() for k, v in P.items()}
    opt = O.AdamW(O.make_trainable(P, hp), lr=1e-4)
    l = O.train_func(P, opt, inp["batch"], hp, O.alpha_cumprod(hp), True, t=inp["t"], noise_t=inp["noise_t"], noise_1=inp["noise_1"])
    np.testing.assert_allclose([x.item() for x in l], g["train_losses"][0], rtol=2e-5)
    names = [str(n) for n in g["grad_names"]]
    gscale = float(g["grad_norms"].max())
    for n, ref_norm, dn in zip(names, g["grad_norms"], g["after_delta_norms"]):
        grad = P[n].grad
        assert abs(float(grad.double().norm()) - ref_norm) <= 3e-4 * max(ref_norm, 1e-4 * gscale), n
        ref = torch.from_numpy(g["grad::" + n])
        assert float((grad.reshape(-1)[:256].double() - ref.double()).norm()) <= 3e-4 * max(float(ref.double().norm()), 1e-4 * gscale), n
        if not (0.0 < ref_norm < 1e-4 * gscale):
            assert abs(float((P[n].detach().double() - P0[n].double()).norm()) - dn) <= 1e-3 * dn + 1e-9, n
