"""CPU (-m "not gpu"): checkpoint interop with the reference's whole-module pickles (CLIP-DDPM.py:551,560,570).

The "reference side" here is a module built exactly as CLIP-DDPM.py:227-256 builds it (HF DistilBertForMaskedLM with its
in/out embeddings replaced by nn.Sequential(), frozen embedding / lm_head, CLIP linears, segment embedding), living in
`__main__` under the name DistilBertModel, as in the reference script."""
import importlib
import os
import sys

import pytest
import torch
from torch import nn

from _util import O, ROOT

ck = importlib.import_module("diffusion-image-captioning_b200.checkpoint")
transformers = pytest.importorskip("transformers")


def _hp(**kw):
    hp = O.default_hparams()
    hp.update(N_LAYERS=1, VOCAB_SIZE=211, MAX_POSITION=32)
    hp.update(kw)
    return hp


class _RefLike(nn.Module):
    """Constructor body of the reference class (TRAIN_EMBEDDING=False, concat)."""

    def __init__(self, hp):
        super().__init__()
        cfg = ck._hf_config(hp)
        origin = transformers.DistilBertForMaskedLM(cfg)
        self.model = transformers.DistilBertForMaskedLM(cfg)
        import copy
        self.embedding = copy.deepcopy(origin.get_input_embeddings().requires_grad_(False))
        self.lm_head = copy.deepcopy(origin.get_output_embeddings().requires_grad_(False))
        self.lm_head.bias.data = torch.zeros(self.lm_head.bias.data.shape)
        self.model.set_input_embeddings(nn.Sequential())
        self.model.set_output_embeddings(nn.Sequential())
        self.image_linear = nn.Linear(512, hp["DIM"])
        self.text_linear = nn.Linear(512, hp["DIM"])
        self.segment_embedding = nn.Embedding(2, hp["DIM"])

    def forward(self, x):
        return self.model(inputs_embeds=x)[0]


_RefLike.__module__ = "__main__"
_RefLike.__qualname__ = _RefLike.__name__ = "DistilBertModel"


class _as_main:
    """Installs a class as __main__.DistilBertModel for the duration (what running inside the reference script looks like)."""

    def __init__(self, cls):
        self.cls = cls

    def __enter__(self):
        self.main = sys.modules["__main__"]
        self.had = hasattr(self.main, "DistilBertModel")
        self.prev = getattr(self.main, "DistilBertModel", None)
        if self.cls is not None:
            setattr(self.main, "DistilBertModel", self.cls)
        elif self.had:
            delattr(self.main, "DistilBertModel")

    def __exit__(self, *a):
        if self.had:
            setattr(self.main, "DistilBertModel", self.prev)
        elif hasattr(self.main, "DistilBertModel"):
            delattr(self.main, "DistilBertModel")


def test_reference_whole_module_pickle_loads_without_the_reference_script(tmp_path):
    hp = _hp()
    torch.manual_seed(0)
    ref = _RefLike(hp)
    path = tmp_path / "model.pickle"
    with _as_main(_RefLike):
        torch.save(ref.cpu(), path)  # CLIP-DDPM.py:560
    with _as_main(None):  # a process that does not define the reference's class
        with pytest.raises(Exception, match="allow_pickle"):
            ck.load_reference_checkpoint(str(path))          # whole-module pickles run code on load: only on request
        sd = ck.load_reference_checkpoint(str(path), allow_pickle=True)
    want = {k: v for k, v in ref.state_dict().items() if not k.endswith("position_ids")}
    assert set(sd) == set(want)
    for k in want:
        assert torch.equal(sd[k], want[k]), k
    names = {n for n, _ in O.param_names(hp)}
    assert names <= set(sd) | {"embedding.weight"}  # every tensor the engine's flat buffer holds is present under the same name
    assert "lm_head.weight" in sd and "lm_head.bias" in sd and "embedding.weight" in sd


def test_export_is_loadable_by_the_reference_class(tmp_path):
    hp = _hp()
    torch.manual_seed(1)
    src = _RefLike(hp)
    sd = ck.load_reference_checkpoint(src)
    path = tmp_path / "export.pickle"
    with _as_main(None):
        ck.save_reference_pickle(sd, hp, str(path))  # written by a process without the reference's class
        assert not hasattr(sys.modules["__main__"], "DistilBertModel")
    with _as_main(_RefLike):  # the reference script: torch.load resolves __main__.DistilBertModel to ITS class (:570)
        back = torch.load(str(path), weights_only=False)
    assert type(back) is _RefLike
    bs = back.state_dict()
    for k, v in src.state_dict().items():
        assert torch.equal(bs[k], v), k
    x = torch.randn(2, 18, hp["DIM"])
    back.eval(); src.eval()
    assert torch.allclose(back(x), src(x))  # behaviour comes from the class, weights from the pickle


def test_export_rejects_incomplete_or_misshapen_state():
    hp = _hp()
    sd = ck.load_reference_checkpoint(_RefLike(hp))
    bad = dict(sd); bad.pop("image_linear.weight")
    with pytest.raises(KeyError):
        ck.to_reference_module(bad, hp)
    bad = dict(sd); bad["text_linear.bias"] = torch.zeros(3)
    with pytest.raises(ValueError):
        ck.to_reference_module(bad, hp)


def test_train_embedding_layout_round_trip():
    hp = _hp(TRAIN_EMBEDDING=True, IN_CHANNEL=16)
    torch.manual_seed(2)
    m = ck.to_reference_module(_te_state(hp), hp)
    sd = ck.load_reference_checkpoint(m)
    assert tuple(sd["embedding.weight"].shape) == (hp["VOCAB_SIZE"], 16)
    assert tuple(sd["lm_head.weight"].shape) == (hp["VOCAB_SIZE"], 16) and "lm_head.bias" not in sd
    assert tuple(sd["input_projection.weight"].shape) == (hp["DIM"], 16)
    assert tuple(sd["output_projection.weight"].shape) == (16, hp["DIM"])


def _te_state(hp):
    base = ck.load_reference_checkpoint(_RefLike(hp))
    V, d = hp["VOCAB_SIZE"], hp["DIM"]
    base.pop("lm_head.bias")
    base.update({"embedding.weight": torch.randn(V, 16), "lm_head.weight": torch.randn(V, 16),
                 "input_projection.weight": torch.randn(d, 16), "input_projection.bias": torch.randn(d),
                 "output_projection.weight": torch.randn(16, d), "output_projection.bias": torch.randn(16)})
    return base


class _FakeModel:
    """state_dict()/load_state_dict() surface of clipdlm.DistilBertModel (the real one needs a GPU)."""

    def __init__(self, sd, hp):
        self.sd, self.hp, self.precision = {k: v.clone() for k, v in sd.items()}, hp, "bf16"

    def state_dict(self):
        return self.sd

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self.sd if k not in sd]
        if strict and missing:
            raise KeyError(missing)
        for k in self.sd:
            if k in sd:
                self.sd[k] = sd[k].clone()


class _FakeTrainer:
    def __init__(self, n):
        self.m, self.v, self.t, self.param_groups = torch.randn(n), torch.rand(n), 7, [dict(lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)]

    def state_dict(self):
        return dict(m=self.m.clone(), v=self.v.clone(), t=self.t, param_groups=[dict(g) for g in self.param_groups])

    def load_state_dict(self, sd):
        self.m, self.v, self.t, self.param_groups = sd["m"].clone(), sd["v"].clone(), int(sd["t"]), [dict(g) for g in sd["param_groups"]]


def test_resume_checkpoint_round_trip(tmp_path):
    hp = _hp()
    sd = ck.load_reference_checkpoint(_RefLike(hp))
    a, ta = _FakeModel(sd, hp), _FakeTrainer(11)
    path = tmp_path / "resume.pt"
    ck.save_checkpoint(a, ta, str(path), epoch=3, extra=dict(note="x"))
    b, tb = _FakeModel({k: torch.zeros_like(v) for k, v in sd.items()}, hp), _FakeTrainer(11)
    epoch, extra = ck.load_checkpoint(b, str(path), tb)     # this package's own file: weights_only=True suffices, no unpickler involved
    assert epoch == 3 and extra == dict(note="x")
    assert all(torch.equal(b.sd[k], sd[k]) for k in sd)
    assert torch.equal(tb.m, ta.m) and torch.equal(tb.v, ta.v) and tb.t == 7 and tb.param_groups[0]["lr"] == 3e-5


@pytest.mark.skipif(not os.path.exists("/root/reference/CLIP-DDPM.py"), reason="the real reference is only present in the build container")
def test_real_reference_class_pickle(tmp_path):
    """Same as the first test but with the REAL reference class (exec'd from its own source by oracle/reference_harness.py)."""
    from oracle import reference_harness as RH
    hp = O.default_hparams(); hp.update(N_LAYERS=1, VOCAB_SIZE=211)
    ns = RH.build_namespace(hp)
    cls = ns["DistilBertModel"]
    cls.__module__ = "__main__"
    model = RH.build_model(ns, hp, seed=0)
    path = tmp_path / "ref.pickle"
    with _as_main(cls):
        torch.save(model.cpu(), str(path))
    with _as_main(None):
        sd = ck.load_reference_checkpoint(str(path), allow_pickle=True)
        out = tmp_path / "ours.pickle"
        hp2 = dict(hp, MAX_POSITION=512)
        ck.save_reference_pickle(sd, hp2, str(out))
    with _as_main(cls):
        back = torch.load(str(out), weights_only=False)
    assert type(back) is cls
    for k, v in model.state_dict().items():
        assert torch.equal(back.state_dict()[k], v), k
    B = 2
    x = torch.randn(B, 16, 768); img = torch.randn(B, 1, 512); txt = torch.randn(B, 1, 512)
    mask = torch.ones(B, 16, dtype=torch.int64); cm = torch.tensor([[1, 0]]).repeat(B, 1)
    model.eval(); back.eval()
    with torch.no_grad():
        a, b = model(x, img, txt, mask, cm), back(x, img, txt, mask, cm)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_adamw_state_refuses_a_foreign_slice():
    """under the fused data-parallel step a rank holds the Adam moments of its parameter slice only. Loading rank 0's slice
    on rank 1 must raise (equal slice lengths made it silent before); full-size states are cut to the rank's own slice."""
    from clipdlm.model import AdamW
    import types
    n = 64
    def trainer(lo, hi):
        tr = AdamW.__new__(AdamW)
        tr.model = types.SimpleNamespace(n_params=n, device="cpu")
        tr._slice, tr.m, tr.v, tr.t, tr.param_groups = (lo, hi), torch.zeros(hi - lo), torch.zeros(hi - lo), 0, [dict(lr=1e-4)]
        return tr
    r0, r1 = trainer(0, 32), trainer(32, 64)
    r0.m += 1.0; r0.v += 2.0; r0.t = 5
    sd0 = r0.state_dict()
    assert sd0["slice"] == (0, 32) and sd0["n_params"] == n
    with pytest.raises(ValueError, match="slice"):
        r1.load_state_dict(sd0)
    r0b = trainer(0, 32); r0b.load_state_dict(sd0)
    assert torch.equal(r0b.m, r0.m) and r0b.t == 5
    full = dict(m=torch.arange(n, dtype=torch.float32), v=torch.arange(n, dtype=torch.float32) * 2, t=9, param_groups=[dict(lr=3e-5)], slice=(0, n), n_params=n)
    r1.load_state_dict(full)
    assert torch.equal(r1.m, torch.arange(32, 64, dtype=torch.float32)) and torch.equal(r1.v, 2 * torch.arange(32, 64, dtype=torch.float32)) and r1.t == 9
    single = trainer(0, n); single.load_state_dict(full)
    assert torch.equal(single.m, full["m"])
    with pytest.raises(ValueError, match="parameters"):
        r1.load_state_dict(dict(full, n_params=n + 1))
    legacy = dict(m=torch.ones(n), v=torch.ones(n), t=1, param_groups=[dict(lr=1e-4)])   # round-1 files: no slice key, full size
    single.load_state_dict(legacy); r1.load_state_dict(legacy)
