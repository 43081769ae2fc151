"""CPU (-m "not gpu"): host logic, the C-ABI library surface, and the no-fallback rule. No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from _util import O, ROOT


def test_library_exports_every_declared_symbol():
    import clipdlm
    from clipdlm import _lib as L
    header = open(os.path.join(ROOT, "include", "clipdlm.h")).read()
    declared = set(re.findall(r"\b(clipdlm_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = L.load()  # resolves every symbol, raises AttributeError otherwise
    nm = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (clipdlm_[a-z0-9_]+)", nm))
    assert declared <= exported
    assert lib.clipdlm_version() >= 100


def test_sass_is_blackwell_native():
    """The GEMM must carry tcgen05 / TMA instructions (UTC*MMA, UTMALDG), not a legacy mma.sync path."""
    from clipdlm import _lib as L
    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass
    assert "UTMALDG" in sass
    assert "LDTM" in sass
    assert "sm_100a" in sass
    # Regression guards read off the same dump (second session of round 2):
    # (1) no non-coherent global load in a tcgen05 GEMM / attention kernel - their only scalar global loads besides streamed activations are
    #     trainable parameters (biases), and LDG.E.CONSTANT (ld.global.nc) served stale biases under programmatic dependent launch chains;
    # (2) the MMA issuers issue under elect.sync with the whole warp in the loop: a `lane == 0` loop wraps every UTCHMMA in an ELECT / R2UR /
    #     BRA.U.ANY waterfall (one ELECT per MMA and more), the converged loops keep a handful per kernel.
    fn, nc, elect, mma = None, {}, {}, {}
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif fn and ("11gemm_kernelI" in fn or "attn_packed" in fn or "attn_umma" in fn or "attn_ring" in fn or ("layernorm_" in fn and "fast" in fn)):
            # = the kernels launched with programmatic dependent launch (not small_gemm_kernel & co: plain launches)
            if "LDG" in line and "CONSTANT" in line:
                nc[fn] = nc.get(fn, 0) + 1
            if " ELECT " in line:
                elect[fn] = elect.get(fn, 0) + 1
            if "UTCHMMA" in line:
                mma[fn] = mma.get(fn, 0) + 1
    assert not nc, f"non-coherent loads in {sorted(nc)[:3]}"
    assert mma and all(elect.get(f, 0) <= 12 for f in mma), {f: (elect.get(f, 0), n) for f, n in mma.items() if elect.get(f, 0) > 12}


def test_param_layout_matches_reference_counts():
    from clipdlm import _lib as L
    from clipdlm import default_hparams
    lib = L.load()
    for layers, dim, heads, hid, expect in ((6, 768, 12, 3072, 44_303_616), (12, 768, 12, 3072, 86_830_848), (24, 1024, 16, 4096, None)):
        cfg = L.Config(layers, dim, heads, hid, 30522, 16, 512, 512, 0, 0, 1e-12, 0.1, 0.1)
        n = lib.clipdlm_param_count(C.byref(cfg))
        hp = O.default_hparams(); hp.update(N_LAYERS=layers, DIM=dim, N_HEADS=heads, HIDDEN_DIM=hid)
        want = sum(int(torch.tensor(s).prod()) for _, s in O.param_names(hp))
        assert n == want
        if expect is not None:
            assert n == expect  # SURVEY App. B / BASELINE.md section 3
        offs = [lib.clipdlm_param_offset(C.byref(cfg), s) for s in range(L.P_LAYER0 + layers * L.P_PER_LAYER)]
        assert offs == sorted(offs) and all(o % 8 == 0 for o in offs)  # 16-byte aligned bf16 shadows for TMA
        assert lib.clipdlm_param_offset(C.byref(cfg), 10_000) == -1
    cfg = L.Config(6, 768, 12, 3072, 30522, 16, 512, 512, 0, 0, 1e-12, 0.1, 0.1)
    train = lib.clipdlm_workspace_bytes(C.byref(cfg), 4096, 512, 1)
    infer = lib.clipdlm_workspace_bytes(C.byref(cfg), 4096, 512, 0)
    assert 0 < infer < train < 40 * 2 ** 30
    bad = L.Config(6, 700, 12, 3072, 30522, 16, 512, 512, 0, 0, 1e-12, 0.1, 0.1)
    assert lib.clipdlm_workspace_bytes(C.byref(bad), 16, 8, 1) == 0 and b"dim" in lib.clipdlm_last_error()


def test_hparams_follow_reference_defaults():
    import clipdlm
    hp = clipdlm.default_hparams()
    ref = O.default_hparams()
    for k, v in ref.items():
        assert hp[k] == v, k
    assert clipdlm.model_name(hp).startswith("epoch5_lossseries_sum_sample_mean_lr1E-04-5E-05_schedulerlinspace_round5E-01_dynamic-1_clipconcat")
    lrs = clipdlm.learning_rates(hp)
    assert len(lrs) == 5 and abs(lrs[0] - 1e-4) < 1e-10 and abs(lrs[-1] - 5e-5) < 1e-10
    assert len(clipdlm.learning_rates(clipdlm.default_hparams(SCHEDULER="cosine"))) == 15  # cosine_annealing(): 5 epochs x 3
    lg = clipdlm.learning_rates(clipdlm.default_hparams(SCHEDULER="logspace"))
    assert abs(lg[2] - (1e-4 * 5e-5) ** 0.5) < 1e-9
    with pytest.raises(KeyError):
        clipdlm.default_hparams(NOT_A_KEY=1)
    assert torch.equal(clipdlm.alpha_cumprod(hp), O.alpha_cumprod(ref))
    assert torch.equal(clipdlm.alpha_cumprod(clipdlm.default_hparams(COSIN_SCHEDULE=False)), O.alpha_cumprod(dict(ref, COSIN_SCHEDULE=False)))


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU (no oracle / eager fallback)."""
    import clipdlm
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(clipdlm.ClipdlmError):
        clipdlm.DistilBertModel(None, None, None)
    with pytest.raises(clipdlm.ClipdlmError):
        clipdlm.diffuse_t(torch.zeros(1, 16, 768), torch.zeros(1, dtype=torch.int64), clipdlm.default_hparams())
    src = ""
    pkg_dir = os.path.join(ROOT, "diffusion-image-captioning_b200")
    for f in os.listdir(pkg_dir):
        if f.endswith(".py"):
            src += open(os.path.join(pkg_dir, f)).read()
    assert "oracle" not in src.replace("no oracle", "")  # the product never imports the checker


def test_shard_range_covers_everything():
    from clipdlm import shard_range
    for n in (0, 1, 7, 512, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_device_caption_dataset_and_loader():
    """pre-tokenised dataset + loader with the reference's split / drop_last / batch-dict semantics."""
    import torch
    import clipdlm
    ds = clipdlm.synthetic_dataset(103, seed=1)
    tr, va = ds.random_split(0.8, torch.Generator().manual_seed(0))
    assert len(tr) == int(103 * 0.8) and len(tr) + len(va) == 103
    assert set(tr.indices.tolist()).isdisjoint(va.indices.tolist())
    loader = tr.loader(8, shuffle=True, generator=torch.Generator().manual_seed(3))
    batches = list(loader)
    assert len(batches) == len(loader) == len(tr) // 8  # drop_last
    b = batches[0]
    assert set(b) >= {"image_clip", "text_clip", "input_ids", "attention_mask", "image"}
    assert tuple(b["input_ids"].shape) == (8, 16) and b["input_ids"].dtype == torch.int64 and tuple(b["image_clip"].shape) == (8, 512)
    seen = torch.cat([x["input_ids"] for x in batches])
    assert seen.shape[0] == len(batches) * 8
    # data-parallel sharding: two ranks see disjoint batches of the same permutation, same count
    r0 = list(tr.loader(8, shuffle=True, generator=torch.Generator().manual_seed(5), rank=0, world=2))
    r1 = list(tr.loader(8, shuffle=True, generator=torch.Generator().manual_seed(5), rank=1, world=2))
    assert len(r0) == len(r1) == (len(tr) // 8) // 2
    a = {tuple(x.tolist()) for bb in r0 for x in bb["input_ids"]}
    c = {tuple(x.tolist()) for bb in r1 for x in bb["input_ids"]}
    assert a.isdisjoint(c)


def test_postprocess_and_bleu():
    """unique_consecutive quirk + BLEU-4 known answers (hand-computed / sacrebleu-style definitions)."""
    import torch
    import clipdlm
    ids = torch.tensor([[5, 5, 7, 7, 9], [1, 1, 2, 3, 3]])
    # batch semantics of the reference: a column goes only if it repeats for EVERY row -> columns 1 (5|1) and 3 (7|3)... col 3 = (7,3) vs col 2 = (7,2): kept
    cols = clipdlm.postprocess(ids)
    assert [r.tolist() for r in cols] == [[5, 7, 7, 9], [1, 2, 3, 3]]
    assert [r.tolist() for r in clipdlm.postprocess(ids, per_sequence=True)] == [[5, 7, 9], [1, 2, 3]]
    assert clipdlm.decode([torch.tensor([3, 4])], lambda i: f"w{i}") == ["w3 w4"]
    cand = ["the cat sat on the mat"]
    assert abs(clipdlm.bleu_score(cand, [["the cat sat on the mat"]]) - 1.0) < 1e-12
    assert clipdlm.bleu_score(["a b c d"], [["e f g h"]]) == 0.0
    # 5 of 6 unigrams, 3 of 5 bigrams, 1 of 4 trigrams, 0 of 3 4-grams -> 0 (no smoothing)
    assert clipdlm.bleu_score(["the cat the cat on mat"], [["the cat sat on the mat"]]) == 0.0
    # brevity penalty: candidate = first 5 tokens of a 6-token reference, all n-grams match: BLEU = exp(1 - 6/5)
    import math
    assert abs(clipdlm.bleu_score(["the cat sat on the"], [["the cat sat on the mat"]]) - math.exp(1 - 6 / 5)) < 1e-12
    # two references: clipping takes the max count over references, length = closest reference
    s = clipdlm.bleu_score(["a a b c d e"], [["a b c d e f", "a a b c d e"]])
    assert abs(s - 1.0) < 1e-12


def test_bleu_published_known_answers():
    """Pins bleu_score() on the doc-test vectors two independent BLEU implementations publish (neither package is in this image, the
    expected values are the ones printed in their docstrings): torchmetrics.functional.bleu_score - the function the reference calls
    through `BLEUScore()` (CLIP-DDPM.py:604-606,629) - and nltk.translate.bleu_score sentence_bleu / corpus_bleu (same definition:
    clipped counts summed over the corpus, uniform 4-gram weights, no smoothing, closest-reference brevity penalty)."""
    import clipdlm
    # torchmetrics/functional/text/bleu.py docstring: tensor(0.7598); analytically (5/6 * 4/5 * 3/4 * 2/3) ** 0.25 = 3 ** -0.25, BP = 1
    s = clipdlm.bleu_score(["the cat is on the mat"], [["there is a cat on the mat", "a cat is on the mat"]])
    assert round(s, 4) == 0.7598 and abs(s - 3.0 ** -0.25) < 1e-12
    h1 = "It is a guide to action which ensures that the military always obeys the commands of the party"
    r1a = "It is a guide to action that ensures that the military will forever heed Party commands"
    r1b = "It is the guiding principle which guarantees the military forces always being under the command of the Party"
    r1c = "It is the practical guide for the army always to heed the directions of the party"
    h2 = "he read the book because he was interested in world history"
    r2a = "he was interested in world history because he read the book"
    # nltk.translate.bleu_score.sentence_bleu docstring: 0.5045666840058485; corpus_bleu docstring: 0.5920778868801042
    assert abs(clipdlm.bleu_score([h1], [[r1a, r1b, r1c]]) - 0.5045666840058485) < 1e-12
    assert abs(clipdlm.bleu_score([h1, h2], [[r1a, r1b, r1c], [r2a]]) - 0.5920778868801042) < 1e-12
    # the reference averages per-batch corpus scores (CLIP-DDPM.py:629-631); a batch with no 4-gram match scores 0, not NaN
    assert clipdlm.bleu_score([h2], [[r1a]]) == 0.0


def _build_c_host(out):
    from clipdlm import _lib as L
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    cuda = "/usr/local/cuda"
    cmd = ["gcc", "-O2", "-std=c99", "-Wall", "-Werror", os.path.join(ROOT, "examples", "c_host.c"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(cuda, "include"), "-L" + os.path.dirname(L.LIB_PATH), "-lclipdlm", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + os.path.dirname(L.LIB_PATH), "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", out]
    return subprocess.run(cmd, capture_output=True, text=True)


def test_c_host_example_compiles_against_the_header(tmp_path):
    """include/clipdlm.h is plain C: a C99 host (examples/c_host.c, no Python / torch) compiles warning-free and links against the .so."""
    r = _build_c_host(str(tmp_path / "c_host"))
    assert r.returncode == 0, r.stderr


class _StubModel:
    """Just enough of DistilBertModel for the epoch loop: hp, training flag, train()/eval()."""

    def __init__(self, hp):
        self.hp, self.training = hp, False

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)


class _StubTrainer:
    def __init__(self, lr):
        self.param_groups = [{"lr": lr}]


def _run_epoch_loop(monkeypatch, hp, train_losses, val_losses, n_batches=3, with_val=True):
    """Drive clipdlm.train() (CLIP-DDPM.py:515-557) with a scripted train_func: (x_t, x_1, prob) per call."""
    import io
    import clipdlm
    from clipdlm import diffusion
    calls = {"train": [], "val": [], "weights": [], "early": []}
    it_train, it_val = iter(train_losses), iter(val_losses)

    def fake_train_func(model, trainer, x, train=True, **kw):
        a, b, c = (torch.tensor(float(v)) for v in (next(it_train) if train else next(it_val)))
        calls["train" if train else "val"].append((trainer.param_groups[0]["lr"] if train else None, model.training, torch.is_grad_enabled()))
        calls["weights"].append(model.hp["ROUNDING_WEIGHT"])
        return a + b + c, a, b, c

    monkeypatch.setattr(diffusion, "train_func", fake_train_func)
    model, trainer = _StubModel(hp), _StubTrainer(hp["LEARNING_RATE"])
    summary = io.StringIO()
    hist = clipdlm.train(model, trainer, [{}] * n_batches, hp, val_loader=[{}] * 2 if with_val else None, summary=summary,
                         on_early_stop=lambda m, e: calls["early"].append(e))
    return hist, calls, summary.getvalue(), model, trainer


def test_epoch_loop_learning_rates_and_summary(monkeypatch):
    import clipdlm
    hp = clipdlm.default_hparams(EPOCH_NUM=3)
    hist, calls, text, model, trainer = _run_epoch_loop(monkeypatch, hp, [(1, 2, 3)] * 9, [(1, 2, 3)] * 6)
    lrs = clipdlm.learning_rates(hp)
    assert [c[0] for c in calls["train"]] == [lr for lr in lrs for _ in range(3)]  # :520-522: lr set once per epoch
    assert all(c[1] and c[2] for c in calls["train"])           # train mode, grad enabled
    assert all(not c[1] and not c[2] for c in calls["val"])     # validate: eval mode under no_grad (:489-490)
    assert model.training                                        # ... and back to train mode (:500)
    assert [float(h["x_t_loss"]) for h in hist] == [1.0] * 3 and [float(h["prob_loss"]) for h in hist] == [3.0] * 3
    assert [float(h["val_x_1"]) for h in hist] == [2.0] * 3
    assert not any(h["early_stopped"] for h in hist) and calls["early"] == []
    lines = text.strip().split("\n")
    assert len(lines) == 3 and lines[1].startswith("epoch 1 average x_t_loss, x_1_loss, prob_loss, val losses: 1.0, 2.0, 3.0, 1.0, 2.0, 3.0")
    # equal endpoints: the reference leaves the optimizer's lr alone (:519)
    hp2 = clipdlm.default_hparams(EPOCH_NUM=2, LEARNING_RATE=5e-5, END_LEARNING_RATE=5e-5)
    _, calls2, _, _, tr2 = _run_epoch_loop(monkeypatch, hp2, [(1, 1, 1)] * 6, [(1, 1, 1)] * 4)
    assert {c[0] for c in calls2["train"]} == {5e-5}


def test_epoch_loop_early_stop_fires_once(monkeypatch):
    import clipdlm
    hp = clipdlm.default_hparams(EPOCH_NUM=3, EARLY_STOP_RATIO=1.05)
    # epoch 0: val 6.0 <= 1.05 * 6.0; epoch 1: val 6.6 > 6.3 -> early stop; epoch 2: still above, hook must NOT fire again (:548-553)
    hist, calls, text, _, _ = _run_epoch_loop(monkeypatch, hp, [(1, 2, 3)] * 9, [(1, 2, 3)] * 2 + [(1.2, 2.2, 3.2)] * 4)
    assert [h["early_stopped"] for h in hist] == [False, True, True]
    assert calls["early"] == [1]
    assert text.count("early stop! \n") == 1 and text.index("early stop!") < text.index("epoch 1 average")


def test_epoch_loop_dynamic_rounding_weight_and_debug(monkeypatch):
    import clipdlm
    hp = clipdlm.default_hparams(EPOCH_NUM=1, DYNAMIC_ROUNDING_WEIGHT=2.0, ROUNDING_WEIGHT=0.5)
    # :535-536: after each batch ROUNDING_WEIGHT <- (acc_x_t + acc_x_1) / acc_prob * DYNAMIC_ROUNDING_WEIGHT (running sums of the epoch)
    hist, calls, _, model, _ = _run_epoch_loop(monkeypatch, hp, [(1, 1, 4), (3, 1, 2), (1, 1, 1)], [], with_val=False)
    assert calls["weights"][0] == 0.5
    assert calls["weights"][1] == pytest.approx(2.0 * 2 / 4)
    assert calls["weights"][2] == pytest.approx(2.0 * 6 / 6)
    assert model.hp["ROUNDING_WEIGHT"] == pytest.approx(2.0 * 8 / 7)
    assert "val_x_t" not in hist[0]
    # DEBUG: one batch, one epoch (:543-544,556-557); the averages still divide by len(train_loader) like the reference (:554)
    hp = clipdlm.default_hparams(EPOCH_NUM=4, DEBUG=True)
    hist, calls, _, _, _ = _run_epoch_loop(monkeypatch, hp, [(3, 6, 9)], [(1, 1, 1)] * 2)
    assert len(hist) == 1 and len(calls["train"]) == 1
    assert float(hist[0]["x_t_loss"]) == pytest.approx(1.0) and float(hist[0]["prob_loss"]) == pytest.approx(3.0)


def test_factored_softmax_gradient_arithmetic():
    """CPU emulation (torch, bf16 roundings where the kernels round) of the factored softmax-CE gradient chain that
    tests/test_experimental_gpu.py checks on the GPU: stored e = bf16(exp(s - c)), fp32 row sums, row factor scale * exp(c - lse),
    one-hot term as a gathered W row. It pins the identity the kernels rely on and the tolerances the GPU test asserts, and shows the
    CE part of the gradient is MORE accurate than the default bf16 path's bf16((softmax - onehot) * scale)."""
    torch.manual_seed(0)
    bf = lambda x: x.to(torch.bfloat16).float()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    M, D, V, scale = 48, 768, 30522, 0.37
    W = bf(torch.randn(V, D) * 0.05).double()
    x = bf(torch.randn(M, D)).double()
    logits = x @ W.t()
    tgt = torch.randint(0, V, (M,))
    ref = ((torch.softmax(logits, -1) - torch.nn.functional.one_hot(tgt, V).double()) * scale) @ W
    ref_lse = torch.logsumexp(logits, -1)
    for shift in (0.0, 3.5):
        e32 = torch.exp2(torch.clamp((logits.float() - shift) * 1.4426950408889634, max=100.0))
        lse = shift + torch.log(e32.sum(-1))
        assert rel(lse, ref_lse) < 1e-6
        rs = scale * torch.exp(shift - lse)
        new = (bf(e32).double() @ W) * rs.double()[:, None] - scale * W[tgt]
        assert rel(new, ref) < 2e-4
    dl = bf(torch.exp(bf(logits.float()) - ref_lse.float()[:, None]) * scale - torch.nn.functional.one_hot(tgt, V).float() * scale)
    old = dl.double() @ W
    assert rel(old, ref) > 5 * rel(new, ref)   # default path: ~2e-3 (bf16 logits, bf16 p - 1); factored path: ~2e-5


def test_exp_shift_bound_formula():
    """DistilBertModel._logit_bound / _shift_from_bound on a stub (CPU tensors): c = clamp((sqrt(D) max|w| + |b|) max_v |W_v| (1 + 2 w_cfg) - 69, 0, 60),
    and the bound really bounds the logits of LayerNorm outputs (and of classifier-free-guidance mixes of two of them)."""
    from clipdlm.model import DistilBertModel
    import types
    torch.manual_seed(0)
    D, V = 768, 500
    for scale, expect_zero, cfg_w in ((0.02, True, 0.0), (0.2, False, 0.0), (5.0, False, 0.0), (0.02, True, 0.5), (0.08, False, 0.5)):
        W = torch.randn(V, D) * scale
        w, b = 1 + 0.1 * torch.randn(D), 0.05 * torch.randn(D)
        stub = types.SimpleNamespace(_lm_head_max_norm=None, lm_head_weight=W, hp={"DIM": D, "CLASSIFIER_FREE_WEIGHT": cfg_w},
                                     _views={"model.vocab_layer_norm.weight": w, "model.vocab_layer_norm.bias": b})
        got = DistilBertModel._logit_bound(stub)
        bound = float((D ** 0.5 * w.abs().max() + b.norm()) * W.norm(dim=1).max()) * (1 + 2 * cfg_w)
        assert float(got) == pytest.approx(bound, rel=1e-5)
        c = float(DistilBertModel._shift_from_bound(got))
        assert c == pytest.approx(min(max(bound - 69.0, 0.0), 60.0), rel=1e-5, abs=1e-6)
        assert (c == 0.0) == expect_zero
        x = torch.nn.functional.layer_norm(torch.randn(64, D) * 3, (D,), w, b, eps=1e-12)
        y = torch.nn.functional.layer_norm(torch.randn(64, D) * 3, (D,), w, b, eps=1e-12)
        mix = (1 + cfg_w) * x - cfg_w * y     # CLIP-DDPM.py:313-317
        assert float((mix @ W.t()).max()) <= bound
    stub = types.SimpleNamespace(_exp_shift=None, fused_softmax_grad=True)
    DistilBertModel.refresh_exp_shift(stub)   # option off: nothing to do


def test_tokenizer_adapter_matches_per_item_tokenisation(tmp_path):
    """N2 (CLIP-DDPM.py:179-197): DeviceCaptionDataset.from_captions tokenises the caption list once; every row must equal what the reference's
    __getitem__ computes for that item - both branches: an HF PreTrainedTokenizer (a DistilBertTokenizer over a small local vocabulary) and the
    DictTokenizer / vocab_dict branch (characters of the caption string, [0] ... [1], UNK padding)."""
    import clipdlm
    from transformers import DistilBertTokenizer
    words = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "a", "dog", "runs", "on", "the", "grass", "two", "men", "play", "##ing", "ball", ".", ","]
    vf = tmp_path / "vocab.txt"
    vf.write_text("\n".join(words))
    tok = DistilBertTokenizer(vocab_file=str(vf), do_lower_case=True)
    caps = ["A dog runs on the grass .", "Two men playing ball , the dog runs on the grass , a dog runs on the grass , two men", "unknownword dog", ""]
    n = len(caps)
    img, txt = torch.randn(n, 512), torch.randn(n, 512)
    ds = clipdlm.DeviceCaptionDataset.from_captions(caps, [f"i{k}.jpg" for k in range(n)], img, txt, tok, max_length=16, chunk=3)
    assert tuple(ds.input_ids.shape) == (n, 16) and ds.input_ids.dtype == torch.int64
    for i, c in enumerate(caps):
        t = tok(text=c, return_tensors="pt", padding="max_length", truncation=True, max_length=16)   # :183
        assert torch.equal(ds.input_ids[i], t["input_ids"].squeeze()) and torch.equal(ds.attention_mask[i], t["attention_mask"].squeeze()), i
    b = ds.batch(torch.tensor([1, 0]))
    assert b["text"] == [caps[1], caps[0]] and b["image"] == ["i1.jpg", "i0.jpg"] and int(b["input_ids"][1, 0]) == words.index("[CLS]")
    vocab = {"UNK": 2, **{ch: 3 + k for k, ch in enumerate("abcdefghijklmnopqrstuvwxyz ")}}
    ds2 = clipdlm.DeviceCaptionDataset.from_captions(caps, None, img, txt, vocab, max_length=16)
    for i, c in enumerate(caps):
        ids = [0] + [vocab.get(x, vocab["UNK"]) for x in c[:16 - 2]] + [1]                            # :185-189
        pad = max(0, 16 - len(ids))
        assert ds2.input_ids[i].tolist() == ids + [vocab["UNK"]] * pad and ds2.attention_mask[i].tolist() == [1] * len(ids) + [0] * pad


def test_pinned_staging_loader_yields_the_same_batches():
    import clipdlm
    ds = clipdlm.synthetic_dataset(40, seed=2)
    ds.image = ds.text = None
    sub = clipdlm.CaptionSubset(ds, torch.arange(40))
    a = list(sub.loader(8, shuffle=True, generator=torch.Generator().manual_seed(1)))
    b = [{k: v.clone() for k, v in x.items()} for x in sub.loader(8, shuffle=True, generator=torch.Generator().manual_seed(1), pin_staging=True)]
    assert len(a) == len(b) == 5
    for x, y in zip(a, b):
        assert all(torch.equal(x[k], y[k]) for k in x)
