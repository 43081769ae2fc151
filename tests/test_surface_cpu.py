"""CPU: the verbatim call surface. The step functions must be callable exactly as CLIP-DDPM.py calls them:
`diffuse_t(x, t)` (:347), `generate_diffuse_pair(x_0, t, t_next=None)` (:364), `loss(model, x_t, x_1, x_tgt, x_0, image_clip, text_clip, mask,
idx, loss_func)` (:382), `train_func(model, trainer, x, train=True)` (:458), `validate(model)` (:488), `DistilBertModel(embedding=None,
projection=None, config=None).forward(x, image_clip, text_clip, mask, concat_mask)` (:228,271)."""
import ast
import inspect
import os

import pytest
import torch

REF = "/root/reference/CLIP-DDPM.py"
# (name, leading positional parameters with the reference's names, {name: default})
EXPECTED = {
    "diffuse_t": (["x", "t"], {}),
    "generate_diffuse_pair": (["x_0", "t", "t_next"], {"t_next": None}),
    "loss": (["model", "x_t", "x_1", "x_tgt", "x_0", "image_clip", "text_clip", "mask", "idx", "loss_func"], {}),
    "train_func": (["model", "trainer", "x", "train"], {"train": True}),
    "validate": (["model"], {}),
}
EXPECTED_CLASS = {"__init__": (["self", "embedding", "projection", "config"], {"embedding": None, "projection": None, "config": None}),
                  "forward": (["self", "x", "image_clip", "text_clip", "mask", "concat_mask"], {})}


def _check(fn, names, defaults, required_exact=True):
    sig = inspect.signature(fn)
    params = list(sig.parameters.values())
    lead = [p.name for p in params[:len(names)]]
    assert lead == names, (fn.__name__, lead, names)
    for p in params[:len(names)]:
        assert p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD), (fn.__name__, p.name)
        if p.name in defaults:
            assert p.default == defaults[p.name] or p.default is defaults[p.name], (fn.__name__, p.name, p.default)
    # everything after the reference's parameters is optional: the reference's call forms must bind
    for p in params[len(names):]:
        assert p.default is not inspect.Parameter.empty or p.kind in (p.KEYWORD_ONLY, p.VAR_KEYWORD, p.VAR_POSITIONAL), (fn.__name__, p.name)
    if required_exact:   # and the parameters the reference requires are exactly the ones required here (loss_func may be omitted)
        req = [p.name for p in params if p.default is inspect.Parameter.empty and p.kind == p.POSITIONAL_OR_KEYWORD]
        assert req == [n for n in names if n not in defaults and n != "loss_func"], (fn.__name__, req)


def test_signatures_match_the_reference_call_forms():
    import clipdlm
    for name, (names, defaults) in EXPECTED.items():
        _check(getattr(clipdlm, name), names, defaults)
    for name, (names, defaults) in EXPECTED_CLASS.items():
        _check(getattr(clipdlm.DistilBertModel, name), names, defaults, required_exact=False)
    assert issubclass(clipdlm.DistilBertModel, torch.nn.Module)
    F = clipdlm.bind(clipdlm.default_hparams(BATCH_SIZE=4), val_loader=[1, 2])
    try:
        assert F.hp["BATCH_SIZE"] == 4 and F.diffuse_t is clipdlm.diffuse_t and F.validate is clipdlm.validate
        from clipdlm import hparams
        assert hparams.ACTIVE["val_loader"] == [1, 2] and hparams.GLOBALS is F.hp
        with pytest.raises(KeyError):
            clipdlm.set_globals(NOT_A_KEY=1)
    finally:
        clipdlm.set_globals(clipdlm.default_hparams())
        hparams.ACTIVE["val_loader"] = None


@pytest.mark.skipif(not os.path.exists(REF), reason="the real reference is only present in the build container")
def test_expected_signatures_are_the_references():
    """The table above against the reference's own source (AST, nothing executed)."""
    tree = ast.parse(open(REF).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in EXPECTED:
            found[node.name] = node
        if isinstance(node, ast.ClassDef) and node.name == "DistilBertModel":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name in EXPECTED_CLASS:
                    found["DistilBertModel." + sub.name] = sub
    def sig(node):
        a = node.args
        names = [x.arg for x in a.args]
        defaults = {n: ast.literal_eval(d) for n, d in zip(names[len(names) - len(a.defaults):], a.defaults)}
        return names, defaults
    for name, exp in EXPECTED.items():
        assert sig(found[name]) == exp, name
    for name, exp in EXPECTED_CLASS.items():
        assert sig(found["DistilBertModel." + name]) == exp, name
