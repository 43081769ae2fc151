"""Pins oracle/clipdlm_oracle.py against the REAL reference executed in this container (exec'd CLIP-DDPM.py slices driving
HF DistilBertForMaskedLM). Run here (CPU): `python oracle/validate_against_reference.py`. Exit code 0 = every check passed.

Checks (small model shapes so the whole script runs in ~1 min):
  1. alpha_cumprod (cosine and linear)              bit-exact
  2. diffuse_t with injected noise                   bit-exact
  3. DistilBertModel.forward, concat + add fusion    <= 1e-5 abs on x_out / logits (eval mode), identical argmax
  4. loss() three terms, all four LOSS_FUNCs          <= 1e-5 rel
  5. one full train_func step (dropout 0): losses <= 1e-5 rel, every gradient <= 2e-4 of scale; AdamW restatement vs torch <= 1e-6
  6. the 5-step denoise loop: final argmax ids        identical
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import clipdlm_oracle as O  # noqa: E402
from oracle import reference_harness as H  # noqa: E402


def small_hp(**kw):
    hp = O.default_hparams()
    hp.update(BATCH_SIZE=3, SAMPLE_SIZE=4, N_LAYERS=2, VOCAB_SIZE=997, DROPOUT=0.0, ATTENTION_DROPOUT=0.0)
    hp.update(kw)
    return hp


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main() -> int:
    if not H.available():
        print("reference or transformers unavailable: nothing validated")
        return 2
    torch.set_num_threads(8)
    fails = []

    def check(name, ok, info=""):
        print(("PASS " if ok else "FAIL ") + name + (" " + info if info else ""))
        if not ok:
            fails.append(name)

    # 1-2. schedule + q_sample
    for cos in (True, False):
        hp = small_hp(COSIN_SCHEDULE=cos)
        ns = H.build_namespace(hp)
        check(f"alpha_cumprod cosine={cos}", torch.equal(ns["alpha_cumprod"], O.alpha_cumprod(hp)))
    hp = small_hp()
    ns = H.build_namespace(hp)
    x = torch.randn(3, 16, 768)
    t = torch.tensor([0, 1, 500, 999]).reshape(4, 1, 1)
    torch.manual_seed(5)
    ref = ns["diffuse_t"](x, t)
    torch.manual_seed(5)
    noise = torch.normal(0, 1, x.shape)
    check("diffuse_t", torch.equal(ref, O.diffuse_t(x, t, O.alpha_cumprod(hp), noise)))

    # 3-6 per fusion / loss function
    for fusion in ("concat", "add"):
        for lf in (("series_sum_sample_mean", "series_sum", "mse_series_mean", "mse_series_sum") if fusion == "concat" else ("series_sum_sample_mean",)):
            hp = small_hp(CLIP_ADDING_METHOD=fusion, LOSS_FUNC=lf)
            ns = H.build_namespace(hp)
            model = H.build_model(ns, hp, seed=1)
            P = H.export_params(model)
            acp = O.alpha_cumprod(hp)
            batch = O.synthetic_batch(hp, seed=3, ragged=True)
            S, B = hp["SAMPLE_SIZE"], hp["BATCH_SIZE"]
            if lf == "series_sum_sample_mean":
                model.eval()
                R = 5
                xin = torch.randn(R, 16, 768)
                img, txt = torch.randn(R, 1, 512), torch.randn(R, 1, 512)
                mask = (torch.rand(R, 16) > 0.3).long(); mask[:, 0] = 1
                cm = torch.tensor([1, 0]).repeat(R, 1)
                with torch.no_grad():
                    lo_r, xo_r = model(xin, img, txt, mask, cm)
                    lo_o, xo_o = O.model_forward(P, xin, img, txt, mask, cm, hp, False)
                check(f"forward[{fusion}] x_out", (xo_r - xo_o).abs().max() < 1e-5, f"max_abs={(xo_r - xo_o).abs().max():.2e}")
                check(f"forward[{fusion}] logits+argmax", (lo_r - lo_o).abs().max() < 1e-5 and torch.equal(lo_r.argmax(-1), lo_o.argmax(-1)))
                # denoise loop
                torch.manual_seed(11)
                restored = torch.randn(B, xo_r.shape[1], 768)
                r = restored.clone()
                with torch.no_grad():
                    for _ in range(5):
                        out, r = model(r[:, :16, :], batch["image_clip"].unsqueeze(1), torch.zeros_like(batch["image_clip"]).unsqueeze(1),
                                       torch.ones(B, 16), torch.tensor([1, 0]).repeat(B, 1))
                    ids_ref = torch.softmax(out, -1).argmax(-1)
                ids_o, _ = O.sample(P, batch["image_clip"], hp, 5, restored.clone())
                check(f"denoise[{fusion}] argmax ids", torch.equal(ids_ref, ids_o))
                model.train()
            # one train step with pinned draws (dropout is 0 in small_hp)
            tt = torch.tensor([0, 17, 400, 999]).reshape(S, 1, 1)
            opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
            torch.manual_seed(21)
            ns_randint = torch.randint
            ns["torch"] = torch
            # pin t by monkeypatching randint inside the reference namespace for this call
            class _T:
                def __getattr__(self, k):
                    return getattr(torch, k)
                def randint(self, *a, **k):
                    return tt.clone()
            ns["torch"] = _T()
            l_ref, a_ref, b_ref, c_ref = ns["train_func"](model, opt, batch)
            ns["torch"] = torch
            torch.manual_seed(21)
            n_t = torch.normal(0, 1, (B, 16, 768)); n_1 = torch.normal(0, 1, (B, 16, 768))
            Po = {k: v.clone() for k, v in P.items()}
            trainable = O.make_trainable(Po, hp)
            oopt = O.AdamW(trainable, lr=1e-3)
            l_o, a_o, b_o, c_o = O.train_func(Po, oopt, batch, hp, acp, True, t=tt, noise_t=n_t, noise_1=n_1)
            ok = max(rel(a_o.detach(), a_ref.detach()), rel(b_o.detach(), b_ref.detach()), rel(c_o.detach(), c_ref.detach())) < 1e-5
            check(f"train_func[{fusion},{lf}] losses", ok, f"ref=({a_ref.item():.5f},{b_ref.item():.5f},{c_ref.item():.5f}) oracle=({a_o.item():.5f},{b_o.item():.5f},{c_o.item():.5f})")
            # gradients of every trainable tensor (k_lin.bias / text_linear have analytically zero gradients: compare against the
            # global gradient scale, not per tensor)
            gref = {n: p.grad.detach() for n, p in model.named_parameters() if p.grad is not None}
            gscale = max(float(g.double().norm()) for g in gref.values())
            names = [k for k in O.trainable_names(hp) if not (Po[k].grad is None and k not in gref)]  # unused tensors: None in both
            worst = max(float((Po[k].grad.double() - gref[k].double()).norm()) / max(float(gref[k].double().norm()), 1e-4 * gscale)
                        for k in names)
            check(f"train_func[{fusion},{lf}] gradients", worst < 2e-4, f"worst rel={worst:.2e}")
            zero_txt = fusion == "concat"
            check(f"train_func[{fusion},{lf}] text_linear grad is exactly zero (key 17 masked)",
                  (not zero_txt) or (float(gref["text_linear.weight"].abs().max()) == 0.0 and float(Po["text_linear.weight"].grad.abs().max()) == 0.0))

    # 7. classifier-free-guidance training (CLIP-DDPM.py:313-317,406-410): losses + gradients with the reference's own guidance draw
    for fusion in ("concat", "add"):
        hp = small_hp(CLIP_ADDING_METHOD=fusion, CLASSIFIER_FREE_WEIGHT=0.3, CLASSIFIER_FREE_PROB=0.4)
        ns = H.build_namespace(hp)
        model = H.build_model(ns, hp, seed=2)
        P = H.export_params(model)
        batch = O.synthetic_batch(hp, seed=4, ragged=True)
        S, B = hp["SAMPLE_SIZE"], hp["BATCH_SIZE"]
        x_0 = P["embedding.weight"][batch["input_ids"]]
        tt = torch.tensor([3, 200, 600, 950]).reshape(S, 1, 1)
        acp = O.alpha_cumprod(hp)
        g = torch.Generator().manual_seed(31)
        x_t = O.diffuse_t(x_0, tt, acp, torch.randn(x_0.shape, generator=g))
        x_1 = O.diffuse_t(x_0, torch.ones(1, dtype=torch.int64), acp, torch.randn(x_0.shape, generator=g))
        torch.manual_seed(77)
        lr = ns["loss"](model, x_t, x_1, None, x_0, batch["image_clip"], batch["text_clip"], batch["attention_mask"], batch["input_ids"],
                        ns["LOSS_FUNC"])
        sum(lr).backward()
        torch.manual_seed(77)
        cmask = (torch.rand((S * B, 1)) > hp["CLASSIFIER_FREE_PROB"]).type(torch.float32)
        cmask[0] = 0; cmask[1] = 1
        Po = {k: v.clone() for k, v in P.items()}
        O.make_trainable(Po, hp)
        lo = O.loss(Po, x_t, x_1, None, x_0, batch["image_clip"], batch["text_clip"], batch["attention_mask"], batch["input_ids"], hp,
                    train=True, classifier_mask=cmask)
        sum(lo).backward()
        ok = max(rel(a.detach(), b.detach()) for a, b in zip(lo, lr)) < 1e-5
        check(f"CFG loss[{fusion}]", ok, f"ref={[round(float(v), 5) for v in lr]} oracle={[round(float(v), 5) for v in lo]}")
        gref = {n: p.grad.detach() for n, p in model.named_parameters() if p.grad is not None}
        gscale = max(float(v.double().norm()) for v in gref.values())
        names = [k for k in O.trainable_names(hp) if not (Po[k].grad is None and k not in gref)]
        worst = max(float((Po[k].grad.double() - gref[k].double()).norm()) / max(float(gref[k].double().norm()), 1e-4 * gscale) for k in names)
        check(f"CFG gradients[{fusion}]", worst < 2e-4, f"worst rel={worst:.2e}")

    # 8. TRAIN_EMBEDDING=True (CLIP-DDPM.py:238-243,292-293,319-320): 16-channel learned embedding, trainable lm_head and in/out
    #    projections; the gradient reaches embedding.weight through x_t, x_1 AND the loss target x_0
    for lf, x0pred in (("series_sum_sample_mean", True), ("mse_series_mean", True), ("series_sum_sample_mean", False)):
        hp = small_hp(TRAIN_EMBEDDING=True, IN_CHANNEL=16, LOSS_FUNC=lf, X_0_PREDICTION=x0pred, X_T_STEP_INTERVAL=100)
        ns = H.build_namespace(hp)
        model = H.build_model(ns, hp, seed=5)
        P = H.export_params(model)
        acp = O.alpha_cumprod(hp)
        batch = O.synthetic_batch(hp, seed=6, ragged=True)
        S, B = hp["SAMPLE_SIZE"], hp["BATCH_SIZE"]
        model.eval()
        R = 4
        xin = torch.randn(R, 16, 16)
        img, txt = torch.randn(R, 1, 512), torch.randn(R, 1, 512)
        mask = (torch.rand(R, 16) > 0.3).long(); mask[:, 0] = 1
        cm = torch.tensor([1, 0]).repeat(R, 1)
        with torch.no_grad():
            lo_r, xo_r = model(xin, img, txt, mask, cm)
            lo_o, xo_o = O.model_forward(P, xin, img, txt, mask, cm, hp, False)
        check(f"TRAIN_EMBEDDING forward[{lf}]", (xo_r - xo_o).abs().max() < 1e-5 and (lo_r - lo_o).abs().max() < 1e-4 and tuple(xo_r.shape) == (R, 18, 16),
              f"max_abs x_out={(xo_r - xo_o).abs().max():.2e} logits={(lo_r - lo_o).abs().max():.2e}")
        model.train()
        x_0r = model.embedding(batch["input_ids"])
        tt = torch.tensor([3, 200, 600, 950]).reshape(S, 1, 1)
        g = torch.Generator().manual_seed(41)
        n_t, n_1, n_g = (torch.randn(B, 16, 16, generator=g) for _ in range(3))
        one = torch.ones(1, dtype=torch.int64)
        t_next = torch.max(tt - hp["X_T_STEP_INTERVAL"], torch.zeros_like(tt))
        x_t = O.diffuse_t(x_0r, tt, acp, n_t); x_1 = O.diffuse_t(x_0r, one, acp, n_1)
        x_tgt = None if x0pred else O.diffuse_t(x_0r, t_next, acp, n_g)
        lr = ns["loss"](model, x_t, x_1, x_tgt, x_0r, batch["image_clip"], batch["text_clip"], batch["attention_mask"], batch["input_ids"], ns["LOSS_FUNC"])
        sum(lr).backward()
        Po = {k: v.clone() for k, v in P.items()}
        O.make_trainable(Po, hp)
        x_0o = torch.nn.functional.embedding(batch["input_ids"], Po["embedding.weight"])
        x_to = O.diffuse_t(x_0o, tt, acp, n_t); x_1o = O.diffuse_t(x_0o, one, acp, n_1)
        x_tgto = None if x0pred else O.diffuse_t(x_0o, t_next, acp, n_g)
        lo = O.loss(Po, x_to, x_1o, x_tgto, x_0o, batch["image_clip"], batch["text_clip"], batch["attention_mask"], batch["input_ids"], hp, train=True)
        sum(lo).backward()
        ok = max(rel(a.detach(), b.detach()) for a, b in zip(lo, lr)) < 1e-5
        check(f"TRAIN_EMBEDDING loss[{lf},x0={x0pred}]", ok, f"ref={[round(float(v), 5) for v in lr]} oracle={[round(float(v), 5) for v in lo]}")
        gref = {n: p.grad.detach() for n, p in model.named_parameters() if p.grad is not None}
        gscale = max(float(v.double().norm()) for v in gref.values())
        names = [k for k in O.trainable_names(hp) if not (Po[k].grad is None and k not in gref)]
        worst = max(float((Po[k].grad.double() - gref[k].double()).norm()) / max(float(gref[k].double().norm()), 1e-4 * gscale) for k in names)
        check(f"TRAIN_EMBEDDING gradients[{lf},x0={x0pred}]", worst < 2e-4 and {"embedding.weight", "lm_head.weight", "input_projection.weight", "output_projection.bias"} <= set(names),
              f"worst rel={worst:.2e} over {len(names)} tensors")

    # AdamW restatement vs torch.optim.AdamW on identical synthetic gradients (3 steps, incl. an all-zero gradient tensor)
    torch.manual_seed(0)
    ps = [torch.randn(7, 5), torch.randn(11), torch.randn(3, 3)]
    ref_p = [p.clone().requires_grad_(True) for p in ps]
    ora_p = [p.clone().requires_grad_(True) for p in ps]
    topt = torch.optim.AdamW(ref_p, lr=1e-2)
    oopt = O.AdamW(ora_p, lr=1e-2)
    for step in range(3):
        gs = [torch.randn_like(p) * (10.0 ** (step - 1)) for p in ps]
        gs[2].zero_()
        for a, b, g in zip(ref_p, ora_p, gs):
            a.grad = g.clone(); b.grad = g.clone()
        topt.step(); oopt.step()
    worst = max(rel(b.detach(), a.detach()) for a, b in zip(ref_p, ora_p))
    check("AdamW restatement (3 steps)", worst < 1e-6, f"worst rel={worst:.2e}")
    print("FAILED: " + ", ".join(fails) if fails else "oracle pinned against the reference: all checks passed")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
