"""Worker of tests/test_dp_fused_gpu.py (launched with torchrun, one rank per GPU): the fused reduce-scatter + AdamW + all-gather step
(csrc/dp_fused.cu, symmetric memory / NVSwitch multicast) against the NCCL all-reduce + full AdamW path on identical inputs."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import clipdlm  # noqa: E402
from clipdlm import parallel  # noqa: E402
from _util import O, golden_hp  # noqa: E402


def main():
    rank, local_rank, world = parallel.init_process_group_from_env("nccl")
    dev = torch.device("cuda", local_rank)
    te = os.environ.get("DP_TEST_TE", "0") == "1"
    hp = golden_hp(BATCH_SIZE=4, SAMPLE_SIZE=3, DROPOUT=0.1, ATTENTION_DROPOUT=0.1, **(dict(TRAIN_EMBEDDING=True, IN_CHANNEL=16) if te else {}))
    P = O.init_params(hp, seed=1, closed_form=False)
    cfg = clipdlm.DistilBertConfig(n_layers=hp["N_LAYERS"], dropout=0.1, attention_dropout=0.1)
    emb = None if te else P["embedding.weight"]
    models, trainers = [], []
    for fused in (True, False):
        m = clipdlm.DistilBertModel(emb, emb, cfg, hp=hp, precision="bf16", chunk_rows=8)
        m.load_state_dict({k: v.clone() for k, v in P.items()})
        parallel.enable_data_parallel(m, fused=fused)
        models.append(m)
        trainers.append(clipdlm.AdamW(m.parameters(), lr=1e-3))
    assert models[0].dp_fused is not None and models[1].dp_fused is None
    for step in range(3):
        batch = {k: v.to(dev) for k, v in O.synthetic_batch(hp, seed=100 * rank + step, ragged=True).items()}
        g = torch.Generator().manual_seed(7 + step)
        t = torch.randint(0, 1000, (hp["SAMPLE_SIZE"], 1, 1), generator=g)
        ch = hp["IN_CHANNEL"]
        n_t, n_1 = torch.randn(4, 16, ch, generator=g), torch.randn(4, 16, ch, generator=g)
        out = [clipdlm.train_func(m, tr, batch, t=t, noise_t=n_t, noise_1=n_1, dropout_seed=1234 + step) for m, tr in zip(models, trainers)]
        for a, b in zip(*out):
            assert abs(a.item() - b.item()) <= 1e-5 * abs(b.item()) or os.environ.get("DP_TEST_VERBOSE"), (step, a.item(), b.item())
    torch.cuda.synchronize()
    a, b = models[0].flat, models[1].flat
    # Same arithmetic in both paths; what differs is summation order (cross-rank sum, and the fp32 atomics of the split-K weight
    # gradients inside EACH model). Adam normalises the gradient, so tensors whose gradient is analytically zero (k_lin.bias: softmax is
    # shift invariant) move by +-lr per step on rounding noise alone: those are bounded by the step size, everything else must agree.
    pa, pb = dict(models[0].named_parameters()), dict(models[1].named_parameters())
    err, worst = 0.0, ""
    if os.environ.get("DP_TEST_VERBOSE"):
        top = sorted(((float((pa[k].double() - pb[k].double()).norm() / pb[k].double().norm().clamp_min(1e-6)), k) for k in pa), reverse=True)[:8]
        print(f"DP_FUSED_TOP rank={rank}", [(f"{e:.2e}", k) for e, k in top], flush=True)
    for k in pa:
        d = float((pa[k] - pb[k]).abs().max())
        assert d <= 2 * 1e-3 * 3 + 1e-6, (k, d)   # no element can differ by more than Adam's step size x steps
        if "k_lin.bias" not in k:
            e = float((pa[k].double() - pb[k].double()).norm() / pb[k].double().norm().clamp_min(1e-6))   # norm-wise: single noise-flipped elements do not dominate
            if e > err:
                err, worst = e, k
    assert err < 2e-4, (worst, err)
    assert torch.equal(models[0].shadow_hi.float(), models[0].flat.bfloat16().float())   # shadow of EVERY slice refreshed (also the peers')
    ref = a.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, a), "ranks diverged"
    assert float(models[0].grad.abs().max()) == 0.0
    lo, hi = models[0].dp_fused["slice"]
    assert trainers[0].m.numel() == max(hi - lo, 4) and trainers[0].m.numel() < a.numel()
    print(f"DP_FUSED_OK rank={rank} world={world} multicast={models[0].dp_fused['multicast']} max_rel_diff={err:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
