"""-m gpu: every hand-written kernel against a plain PyTorch fp32 reference of the same op, through the C-ABI.
pair=True is the parity ("bf16x3") storage mode, pair=False plain bf16."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from _util import O, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import _gpu
    assert torch.cuda.is_available(), "GPU tests selected on a box without CUDA"
    torch.manual_seed(0)
    return _gpu


# ------------------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("shape", [(16, 768, 768), (144, 2304, 768), (1000, 768, 3072), (4096 + 48, 3072, 768)])
def test_gemm_fwd_epilogues(G, shape, pair):
    M, N, K = shape
    a = torch.randn(M, K, device=G.DEV); w = torch.randn(N, K, device=G.DEV) * 0.05
    bias = torch.randn(N, device=G.DEV); res = torch.randn(M, N, device=G.DEV)
    ah, al = G.split(a, pair); wh, wl = G.split(w, pair); rh, rl = G.split(res, pair)
    ref_in_a, ref_in_w, ref_res = G.join(ah, al), G.join(wh, wl), G.join(rh, rl)
    ref = ref_in_a.double() @ ref_in_w.double().t() + bias.double()
    oh, ol = G.empty_pair((M, N), pair)
    G.gemm(a_hi=ah, a_lo=al, b_hi=wh, b_lo=wl, lda=K, ldb=K, M=M, N=N, K=K, epilogue=0, out_hi=oh, out_lo=ol, ldo=N, bias=bias,
           res_hi=rh, res_lo=rl, ldr=N)
    assert rel(G.join(oh, ol), ref + ref_res.double()) < (3e-5 if pair else 4e-3)
    # dual store: pre-activation + gelu
    o2h, o2l = G.empty_pair((M, N), pair)
    G.gemm(a_hi=ah, a_lo=al, b_hi=wh, b_lo=wl, lda=K, ldb=K, M=M, N=N, K=K, epilogue=0, out_hi=oh, out_lo=ol, out2_hi=o2h, out2_lo=o2l,
           ldo=N, bias=bias)
    assert rel(G.join(oh, ol), ref) < (3e-5 if pair else 4e-3)
    assert rel(G.join(o2h, o2l), F.gelu(ref.float()).double()) < (2e-5 if pair else 6e-3)
    # gelu-only output (inference path: out NULL, out2 set) and fp32 output
    o3h, o3l = G.empty_pair((M, N), pair)
    of = torch.zeros(M, N, device=G.DEV)
    G.gemm(a_hi=ah, a_lo=al, b_hi=wh, b_lo=wl, lda=K, ldb=K, M=M, N=N, K=K, epilogue=0, out2_hi=o3h, out2_lo=o3l, out_f32=of, ldo=N, bias=bias)
    assert torch.equal(o3h, o2h)
    assert rel(of, ref) < (3e-5 if pair else 4e-3)


@pytest.mark.parametrize("pair", [False, True])
def test_gemm_dgrad_wgrad(G, pair):
    T, N, K = 1000, 3072, 768  # y[T,N] = x[T,K] W[N,K]^T
    dy = torch.randn(T, N, device=G.DEV); w = torch.randn(N, K, device=G.DEV) * 0.05; x = torch.randn(T, K, device=G.DEV)
    u = torch.randn(T, K, device=G.DEV); res = torch.randn(T, K, device=G.DEV)
    dh, dl = G.split(dy, pair); wh, wl = G.split(w, pair); xh, xl = G.split(x, pair); uh, ul = G.split(u, pair); rh, rl = G.split(res, pair)
    dyr, wr, xr, ur, rr = G.join(dh, dl).double(), G.join(wh, wl).double(), G.join(xh, xl).double(), G.join(uh, ul), G.join(rh, rl).double()
    oh, ol = G.empty_pair((T, K), pair)
    G.gemm(a_hi=dh, a_lo=dl, b_hi=wh, b_lo=wl, lda=N, ldb=K, M=T, N=K, K=N, a_major=0, b_major=1, epilogue=0, out_hi=oh, out_lo=ol, ldo=K,
           res_hi=rh, res_lo=rl, ldr=K)
    assert rel(G.join(oh, ol), dyr @ wr + rr) < (3e-5 if pair else 4e-3)
    G.gemm(a_hi=dh, a_lo=dl, b_hi=wh, b_lo=wl, lda=N, ldb=K, M=T, N=K, K=N, a_major=0, b_major=1, epilogue=0, out_hi=oh, out_lo=ol, ldo=K,
           u_hi=uh, u_lo=ul, ldu=K)
    ur_ = ur.clone().requires_grad_(True)
    F.gelu(ur_).sum().backward()
    assert rel(G.join(oh, ol), (dyr @ wr) * ur_.grad.double()) < (2e-5 if pair else 6e-3)
    acc = torch.full((N, K), 0.5, device=G.DEV)
    G.gemm(a_hi=dh, a_lo=dl, b_hi=xh, b_lo=xl, lda=N, ldb=K, M=N, N=K, K=T, a_major=1, b_major=1, epilogue=1, acc_f32=acc, ldo=K)
    assert rel(acc, dyr.t() @ xr + 0.5) < (3e-5 if pair else 4e-3)


@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("R", [1, 9, 300])
def test_gemm_lse_smgrad_gather_scatter(G, R, pair):
    Ltxt, Lf, D, V = 16, 18, 768, 30522
    M = R * Ltxt
    xo = torch.randn(R * Lf, D, device=G.DEV)
    E = torch.randn(V, D, device=G.DEV) * 0.02
    Epad = torch.zeros((V + 255) // 256 * 256, D, device=G.DEV); Epad[:V] = E
    xh, xl = G.split(xo, pair); eh, el = G.split(Epad, pair)
    B = max(1, R // 3) if R % 3 == 0 else R
    tgt = torch.randint(0, V, (B * Ltxt,), device=G.DEV, dtype=torch.int32)
    nt = (V + 255) // 256
    pm = torch.zeros(2 * nt, M, device=G.DEV); ps = torch.zeros(2 * nt, M, device=G.DEV); pa = torch.zeros(2 * nt, M, device=G.DEV, dtype=torch.int32)
    tl = torch.zeros(M, device=G.DEV)
    G.gemm(a_hi=xh, a_lo=xl, b_hi=eh, b_lo=el, lda=D, ldb=D, M=M, N=V, K=D, gather_len=Ltxt, gather_stride=Lf, epilogue=2,
           part_max=pm, part_sum=ps, part_arg=pa, tgt_logit=tl, targets=tgt, tgt_period=B * Ltxt)
    lse = torch.zeros(M, device=G.DEV); am = torch.zeros(M, device=G.DEV, dtype=torch.int32)
    acc = torch.zeros(1, device=G.DEV, dtype=torch.float64)
    G.L.check(G.lib().clipdlm_lse_combine(pm.data_ptr(), ps.data_ptr(), pa.data_ptr(), 2 * nt, M, tl.data_ptr(), lse.data_ptr(), am.data_ptr(),
                                          acc.data_ptr(), 1.0 / R, G.st()))
    xg = G.join(xh, xl).view(R, Lf, D)[:, :Ltxt].reshape(M, D).double()
    logits = xg @ G.join(eh, el)[:V].double().t()
    ref_lse = torch.logsumexp(logits, -1)
    tfull = tgt.long().repeat(M // (B * Ltxt))
    ref_loss = (ref_lse - logits.gather(1, tfull[:, None])[:, 0]).sum() / R
    assert rel(lse, ref_lse) < (1e-6 if pair else 1e-4)
    assert abs(acc.item() - ref_loss.item()) < (1e-5 if pair else 2e-3) * abs(ref_loss.item())
    top2 = logits.topk(2, -1).values
    safe = (top2[:, 0] - top2[:, 1]) > (1e-4 if pair else 2e-2)
    assert torch.equal(am.long()[safe], logits.argmax(-1)[safe]) and safe.float().mean() > 0.5
    # softmax-CE gradient + dgrad with scatter/residual into [R*Lf, D]
    ldl = nt * 256
    dlh, dll = G.empty_pair((M, ldl), pair)
    G.gemm(a_hi=xh, a_lo=xl, b_hi=eh, b_lo=el, lda=D, ldb=D, M=M, N=V, K=D, gather_len=Ltxt, gather_stride=Lf, epilogue=3, out_hi=dlh,
           out_lo=dll, ldo=ldl, lse=lse, targets=tgt, tgt_period=B * Ltxt, grad_scale=0.37)
    ref_dl = (torch.softmax(logits, -1) - F.one_hot(tfull, V).double()) * 0.37
    assert rel(G.join(dlh, dll)[:, :V], ref_dl) < (2e-5 if pair else 6e-3)
    assert float(G.join(dlh, dll)[:, V:].abs().max()) == 0.0
    g0 = torch.randn(R * Lf, D, device=G.DEV)
    gh, gl = G.split(g0, pair)
    before = G.join(gh, gl).clone()
    G.gemm(a_hi=dlh, a_lo=dll, b_hi=eh, b_lo=el, lda=ldl, ldb=D, M=M, N=D, K=V, a_major=0, b_major=1, epilogue=0, out_hi=gh, out_lo=gl, ldo=D,
           res_hi=gh, res_lo=gl, ldr=D, scatter_len=Ltxt, scatter_stride=Lf)
    ref_g = before.view(R, Lf, D).double().clone()
    ref_g[:, :Ltxt] += (G.join(dlh, dll)[:, :V].double() @ G.join(eh, el)[:V].double()).view(R, Ltxt, D)
    assert rel(G.join(gh, gl), ref_g.view(R * Lf, D)) < (1e-5 if pair else 5e-3)
    assert torch.equal(G.join(gh, gl).view(R, Lf, D)[:, Ltxt:], before.view(R, Lf, D)[:, Ltxt:])  # CLIP rows untouched


def test_gemm_dropout_epilogue(G):
    M, N, K = 512, 768, 256
    a = torch.randn(M, K, device=G.DEV).bfloat16(); w = (torch.randn(N, K, device=G.DEV) * 0.05).bfloat16()
    o1 = torch.zeros(M, N, device=G.DEV); o2 = torch.zeros(M, N, device=G.DEV); o0 = torch.zeros(M, N, device=G.DEV)
    G.gemm(a_hi=a, b_hi=w, lda=K, ldb=K, M=M, N=N, K=K, epilogue=0, out_f32=o0, ldo=N)
    G.gemm(a_hi=a, b_hi=w, lda=K, ldb=K, M=M, N=N, K=K, epilogue=0, out_f32=o1, ldo=N, drop_seed=1234, drop_site=5, drop_p=0.1)
    G.gemm(a_hi=a, b_hi=w, lda=K, ldb=K, M=M, N=N, K=K, epilogue=0, out_f32=o2, ldo=N, drop_seed=1234, drop_site=5, drop_p=0.1)
    assert torch.equal(o1, o2)
    keep = o1 != 0
    assert abs(keep.float().mean().item() - 0.9) < 0.01
    assert rel(o1[keep], o0[keep] / 0.9) < 1e-6


# ------------------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("D", [768, 1024])
def test_layernorm_fwd_bwd(G, D, pair):
    rows = 1003
    z = torch.randn(rows, D, device=G.DEV) * 2 + 0.3
    w = torch.randn(D, device=G.DEV) * 0.2 + 1; b = torch.randn(D, device=G.DEV) * 0.1
    dy = torch.randn(rows, D, device=G.DEV)
    u = torch.randn(rows, D, device=G.DEV)
    zh, zl = G.split(z, pair); dh, dl = G.split(dy, pair); uh, ul = G.split(u, pair)
    zr = G.join(zh, zl).clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    y_ref = F.layer_norm(zr, (D,), wr, br, 1e-12)
    yh, yl = G.empty_pair((rows, D), pair)
    yf = torch.zeros(rows, D, device=G.DEV)
    lib, L = G.lib(), G.L
    L.check(lib.clipdlm_layernorm_fwd(C.byref(G.bfp(zh, zl)), w.data_ptr(), b.data_ptr(), 1e-12, rows, D, C.byref(G.bfp(yh, yl)), yf.data_ptr(),
                                      0, 0, 0.0, G.st()))
    assert rel(yf, y_ref) < 1e-5
    assert rel(G.join(yh, yl), y_ref) < (3e-5 if pair else 4e-3)
    y_ref.backward(G.join(dh, dl))
    gh, gl = G.empty_pair((rows, D), pair)
    dw = torch.full((D,), 0.25, device=G.DEV); db = torch.zeros(D, device=G.DEV); dbias = torch.zeros(D, device=G.DEV)
    L.check(lib.clipdlm_layernorm_bwd(C.byref(G.bfp(zh, zl)), C.byref(G.bfp(dh, dl)), w.data_ptr(), 1e-12, rows, D, C.byref(G.bfp(gh, gl)),
                                      dw.data_ptr(), db.data_ptr(), 0, 0, 0.0, None, 0, 0.0, None, dbias.data_ptr(), G.st()))
    assert rel(G.join(gh, gl), zr.grad) < (2e-5 if pair else 5e-3)
    assert rel(dw - 0.25, wr.grad) < 1e-4 and rel(db, br.grad) < 1e-4
    assert rel(dbias, G.join(gh, gl).sum(0)) < (1e-4 if pair else 5e-3)
    # gelu' fusion (MLM transform head): dz *= gelu'(u)
    ur = G.join(uh, ul).clone().requires_grad_(True)
    F.gelu(ur).sum().backward()
    g2h, g2l = G.empty_pair((rows, D), pair)
    L.check(lib.clipdlm_layernorm_bwd(C.byref(G.bfp(zh, zl)), C.byref(G.bfp(dh, dl)), w.data_ptr(), 1e-12, rows, D, C.byref(G.bfp(g2h, g2l)),
                                      None, None, 0, 0, 0.0, None, 0, 0.0, C.byref(G.bfp(uh, ul)), None, G.st()))
    assert rel(G.join(g2h, g2l), zr.grad * ur.grad) < (3e-5 if pair else 8e-3)


@pytest.mark.parametrize("D", [768, 1024])
@pytest.mark.parametrize("combo", ["dbias", "dz_drop+dbias", "gelu+dbias", "drop_out", "plain"])
def test_layernorm_bwd_fast_matches_generic(G, D, combo):
    """The packed-math specialisations (plain bf16, D = 768 / 1024, the five feature combinations the engine uses) against the generic
    kernel fed the same values through pair storage (lo = 0): same dropout masks, same sums."""
    rows, p = 777, 0.1
    z = (torch.randn(rows, D, device=G.DEV) * 2 + 0.3).bfloat16(); dy = torch.randn(rows, D, device=G.DEV).bfloat16()
    u = torch.randn(rows, D, device=G.DEV).bfloat16()
    w = torch.randn(D, device=G.DEV) * 0.2 + 1
    zero = torch.zeros(rows, D, device=G.DEV, dtype=torch.bfloat16)
    lib, L = G.lib(), G.L
    outs = []
    for pair in (False, True):
        lo = (lambda: zero.clone()) if pair else (lambda: None)
        dz = (torch.zeros_like(z), lo()); dzd = (torch.zeros_like(z), lo())
        dw = torch.zeros(D, device=G.DEV); db = torch.zeros(D, device=G.DEV); dbias = torch.zeros(D, device=G.DEV)
        use_dd, use_g, use_do = combo == "dz_drop+dbias", combo == "gelu+dbias", combo == "drop_out"
        use_b = combo in ("dbias", "dz_drop+dbias", "gelu+dbias")
        L.check(lib.clipdlm_layernorm_bwd(C.byref(G.bfp(z, lo())), C.byref(G.bfp(dy, lo())), w.data_ptr(), 1e-12, rows, D, C.byref(G.bfp(*dz)),
                                          dw.data_ptr(), db.data_ptr(), 4321, 6, p if use_do else 0.0,
                                          C.byref(G.bfp(*dzd)) if use_dd else None, 9, p if use_dd else 0.0,
                                          C.byref(G.bfp(u, lo())) if use_g else None, dbias.data_ptr() if use_b else None, G.st()))
        torch.cuda.synchronize()
        outs.append((G.join(*dz), G.join(*dzd), dw, db, dbias))
    fast, gen = outs
    assert rel(fast[0], gen[0]) < 4e-3 and rel(fast[2], gen[2]) < 1e-4 and rel(fast[3], gen[3]) < 1e-4
    if combo == "dz_drop+dbias":
        assert rel(fast[1], gen[1]) < 4e-3
        assert torch.equal(fast[1] == 0, gen[1] == 0) or ((fast[1] == 0) != (gen[1] == 0)).float().mean() < 1e-4   # same mask
    if combo != "drop_out" and combo != "plain":
        assert rel(fast[4], gen[4]) < 1e-3


def test_layernorm_dropout_masks_consistent(G):
    rows, D, p = 512, 768, 0.1
    z = torch.randn(rows, D, device=G.DEV)
    w = torch.ones(D, device=G.DEV); b = torch.full((D,), 3.0, device=G.DEV)  # b = 3 keeps outputs away from 0
    zh, zl = G.split(z, True)
    yf = torch.zeros(rows, D, device=G.DEV)
    lib, L = G.lib(), G.L
    L.check(lib.clipdlm_layernorm_fwd(C.byref(G.bfp(zh, zl)), w.data_ptr(), b.data_ptr(), 1e-12, rows, D, None, yf.data_ptr(), 77, 3, p, G.st()))
    mask = (yf != 0).float()
    assert abs(mask.mean().item() - (1 - p)) < 0.01
    ref = F.layer_norm(G.join(zh, zl), (D,), w, b, 1e-12) * mask / (1 - p)
    assert rel(yf, ref) < 1e-5
    # the backward regenerates the same mask for dy (drop_out) and applies a second site's mask to dz (dz_drop)
    dy = torch.randn(rows, D, device=G.DEV)
    dh, dl = G.split(dy, True)
    gh, gl = G.empty_pair((rows, D), True); g2h, g2l = G.empty_pair((rows, D), True)
    dbias = torch.zeros(D, device=G.DEV)
    L.check(lib.clipdlm_layernorm_bwd(C.byref(G.bfp(zh, zl)), C.byref(G.bfp(dh, dl)), w.data_ptr(), 1e-12, rows, D, C.byref(G.bfp(gh, gl)),
                                      None, None, 77, 3, p, C.byref(G.bfp(g2h, g2l)), 3, p, None, dbias.data_ptr(), G.st()))
    zr = G.join(zh, zl).clone().requires_grad_(True)
    (F.layer_norm(zr, (D,), w, b, 1e-12) * mask / (1 - p)).backward(G.join(dh, dl))
    assert rel(G.join(gh, gl), zr.grad) < 3e-5
    assert rel(G.join(g2h, g2l), zr.grad * mask / (1 - p)) < 3e-5  # same (seed, site) => same mask
    assert rel(dbias, G.join(g2h, g2l).sum(0)) < 1e-4


# ------------------------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, km_bool, R, L, D, H, mask_mult=None):
    q, k, v = qkv.view(R, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    s = s.masked_fill(~km_bool[:, None, None, :], float("-inf"))
    a = torch.softmax(s, -1)
    if mask_mult is not None:
        a = a * mask_mult
    return (a @ v).permute(0, 2, 1, 3).reshape(R * L, D)


@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("L,R,D", [(18, 37, 768), (16, 5, 768), (66, 7, 1024), (1, 3, 768), (33, 5, 768), (40, 9, 768), (64, 4, 768),
                                   (97, 3, 1024), (128, 2, 768)])
def test_attention_fwd_bwd(G, L, R, D, pair):
    H = D // 64
    qkv = torch.randn(R * L, 3 * D, device=G.DEV)
    dctx = torch.randn(R * L, D, device=G.DEV)
    km = torch.rand(R, L, device=G.DEV) > 0.3
    km[:, min(L - 1, 2)] = True
    kw = (L + 31) // 32
    words = torch.zeros(R, kw, device=G.DEV, dtype=torch.int64)
    for j in range(L):
        words[:, j // 32] |= km[:, j].long() << (j % 32)
    words32 = (words & 0xFFFFFFFF).to(torch.int64)
    words32 = torch.where(words32 >= 2 ** 31, words32 - 2 ** 32, words32).to(torch.int32).contiguous()
    qh, ql = G.split(qkv, pair); dh, dl = G.split(dctx, pair)
    ch, cl = G.empty_pair((R * L, D), pair)
    lib, Lb = G.lib(), G.L
    Lb.check(lib.clipdlm_attn_fwd(C.byref(G.bfp(qh, ql)), words32.data_ptr(), R, L, D, H, C.byref(G.bfp(ch, cl)), 0, 0, 0.0, G.st()))
    qr = G.join(qh, ql).double().requires_grad_(True)
    ref = _attn_ref(qr, km, R, L, D, H)
    assert rel(G.join(ch, cl), ref) < (1e-5 if pair else 5e-3)
    ref.backward(G.join(dh, dl).double())
    gh, gl = G.empty_pair((R * L, 3 * D), pair)
    Lb.check(lib.clipdlm_attn_bwd(C.byref(G.bfp(qh, ql)), words32.data_ptr(), C.byref(G.bfp(dh, dl)), R, L, D, H, C.byref(G.bfp(gh, gl)),
                                  0, 0, 0.0, G.st()))
    assert rel(G.join(gh, gl), qr.grad) < (2e-5 if pair else 6e-3)


def test_attention_dropout_fwd_bwd_consistent(G):
    R, L, D, H, p = 11, 18, 768, 12, 0.1
    qkv = torch.randn(R * L, 3 * D, device=G.DEV)
    v = torch.zeros(R, L, H, 64, device=G.DEV)
    for j in range(L):
        v[:, j, :, j] = 1.0  # v_j = e_j: the context then equals the (dropped) probabilities
    probe = qkv.clone(); probe.view(R, L, 3, H, 64)[:, :, 2] = v
    words = torch.full((R, 1), (1 << L) - 1, device=G.DEV, dtype=torch.int32)
    km = torch.ones(R, L, device=G.DEV, dtype=torch.bool)
    ph, pl = G.split(probe, True)
    ch, cl = G.empty_pair((R * L, D), True)
    lib, Lb = G.lib(), G.L
    Lb.check(lib.clipdlm_attn_fwd(C.byref(G.bfp(ph, pl)), words.data_ptr(), R, L, D, H, C.byref(G.bfp(ch, cl)), 99, 7, p, G.st()))
    a = G.join(ch, cl).view(R, L, H, 64)[..., :L].permute(0, 2, 1, 3)  # [R, H, i, j]
    mask = (a != 0).double()
    assert abs(mask.mean().item() - (1 - p)) < 0.02
    # same seed / site on the real v: forward and backward must use exactly that mask
    qh, ql = G.split(qkv, True)
    Lb.check(lib.clipdlm_attn_fwd(C.byref(G.bfp(qh, ql)), words.data_ptr(), R, L, D, H, C.byref(G.bfp(ch, cl)), 99, 7, p, G.st()))
    qr = G.join(qh, ql).double().requires_grad_(True)
    ref = _attn_ref(qr, km, R, L, D, H, mask / (1 - p))
    assert rel(G.join(ch, cl), ref) < 1e-5
    dctx = torch.randn(R * L, D, device=G.DEV)
    dh, dl = G.split(dctx, True)
    ref.backward(G.join(dh, dl).double())
    gh, gl = G.empty_pair((R * L, 3 * D), True)
    Lb.check(lib.clipdlm_attn_bwd(C.byref(G.bfp(qh, ql)), words.data_ptr(), C.byref(G.bfp(dh, dl)), R, L, D, H, C.byref(G.bfp(gh, gl)), 99, 7, p,
                                  G.st()))
    assert rel(G.join(gh, gl), qr.grad) < 3e-5


def test_attention_mma_path_matches_simt_with_dropout(G):
    """plain bf16, L <= 32 runs on mma.sync tensor-core tiles; the SIMT kernel must see the same dropout mask."""
    R, L, D, H, p = 13, 18, 768, 12, 0.1
    qkv = (torch.randn(R * L, 3 * D, device=G.DEV)).bfloat16()
    dctx = torch.randn(R * L, D, device=G.DEV).bfloat16()
    km = torch.rand(R, L, device=G.DEV) > 0.2
    km[:, 16] = True
    words = torch.zeros(R, 1, device=G.DEV, dtype=torch.int64)
    for j in range(L):
        words[:, 0] |= km[:, j].long() << j
    words = words.to(torch.int32).contiguous()
    lib, Lb = G.lib(), G.L
    outs = {}
    for simt in (0, 1, 2, 3):  # default (packed tcgen05 tiles), fp32 SIMT, mma.sync TMA ring, tcgen05 with 32-row slots
        lib.clipdlm_attn_force_simt(simt)
        ctx = torch.zeros(R * L, D, device=G.DEV, dtype=torch.bfloat16)
        dq = torch.zeros(R * L, 3 * D, device=G.DEV, dtype=torch.bfloat16)
        Lb.check(lib.clipdlm_attn_fwd(C.byref(G.bfp(qkv, None)), words.data_ptr(), R, L, D, H, C.byref(G.bfp(ctx, None)), 4242, 9, p, G.st()))
        Lb.check(lib.clipdlm_attn_bwd(C.byref(G.bfp(qkv, None)), words.data_ptr(), C.byref(G.bfp(dctx, None)), R, L, D, H, C.byref(G.bfp(dq, None)),
                                      4242, 9, p, G.st()))
        torch.cuda.synchronize()
        outs[simt] = (ctx.float(), dq.float())
    lib.clipdlm_attn_force_simt(0)
    for path in (0, 2, 3):
        assert rel(outs[path][0], outs[1][0]) < 1e-2 and rel(outs[path][1], outs[1][1]) < 1.5e-2, path
    # a differing mask would change ~10 % of the probabilities by 100 %: far outside these bounds


@pytest.mark.parametrize("L,R", [(18, 2500), (18, 6), (16, 1031), (18, 7 * 296 + 3)])
def test_attention_packed_tiles_and_folded_bias_gradients(G, L, R):
    """attention_packed.cu (the default for L = 16 / 18, plain bf16): 7 back-to-back sequences per 128-row tcgen05 tile. Many tiles per CTA
    (R = 2500: 358 tiles over 24 CTAs per head), a ragged last tile, dropout on; against the fp32 SIMT kernels (same dropout masks), and the
    bias gradients folded into the backward's epilogue (clipdlm_attn_bwd_bias) against the column sums of its own dqkv: d(q bias) from the
    per-thread accumulators, d(v bias) from the row-sum column of P' on the tensor core, d(k bias) left untouched (analytically zero)."""
    D, H, p = 768, 12, 0.1
    qkv = torch.randn(R * L, 3 * D, device=G.DEV).bfloat16()
    dctx = torch.randn(R * L, D, device=G.DEV).bfloat16()
    km = torch.rand(R, L, device=G.DEV) > 0.2
    km[:, 1] = True
    words = torch.zeros(R, 1, device=G.DEV, dtype=torch.int64)
    for j in range(L):
        words[:, 0] |= km[:, j].long() << j
    words = words.to(torch.int32).contiguous()
    lib, Lb = G.lib(), G.L
    outs = {}
    for path in (0, 1, 4):   # packed + software-pipelined backward (default), fp32 SIMT, packed without the pipeline
        lib.clipdlm_attn_force_simt(path)
        ctx = torch.zeros(R * L, D, device=G.DEV, dtype=torch.bfloat16)
        dq = torch.zeros(R * L, 3 * D, device=G.DEV, dtype=torch.bfloat16)
        dbias = torch.full((3 * D,), 0.25, device=G.DEV)
        folded = torch.zeros(1, dtype=torch.int32)
        Lb.check(lib.clipdlm_attn_fwd(C.byref(G.bfp(qkv, None)), words.data_ptr(), R, L, D, H, C.byref(G.bfp(ctx, None)), 4242, 9, p, G.st()))
        Lb.check(lib.clipdlm_attn_bwd_bias(C.byref(G.bfp(qkv, None)), words.data_ptr(), C.byref(G.bfp(dctx, None)), R, L, D, H, C.byref(G.bfp(dq, None)),
                                           4242, 9, p, dbias.data_ptr(), folded.data_ptr(), G.st()))
        torch.cuda.synchronize()
        outs[path] = (ctx.float(), dq.float(), dbias.clone(), int(folded))
    lib.clipdlm_attn_force_simt(0)
    assert outs[0][3] == 1 and outs[1][3] == 0 and outs[4][3] == 1   # the SIMT path leaves the bias gradients to the caller's colsum
    assert rel(outs[4][1], outs[0][1]) < 1e-6 and rel(outs[4][2], outs[0][2]) < 1e-4   # same arithmetic with and without the pipeline
    assert torch.equal(outs[1][2], torch.full((3 * D,), 0.25, device=G.DEV))
    assert rel(outs[0][0], outs[1][0]) < 1e-2 and rel(outs[0][1], outs[1][1]) < 1.5e-2
    got = outs[0][2] - 0.25                                # += semantics
    want = outs[1][1].double().sum(0)                      # column sums of the fp32-SIMT dqkv
    own = outs[0][1].double().sum(0)
    assert rel(got[:D], want[:D]) < 1e-2 and rel(got[2 * D:], want[2 * D:]) < 1e-2
    assert rel(got[:D], own[:D]) < 5e-3 and rel(got[2 * D:], own[2 * D:]) < 5e-3
    assert float(got[D:2 * D].abs().max()) == 0.0 and float(want[D:2 * D].abs().max()) < 1e-2 * float(want[:D].abs().max())
    # without the dropout, and repeated launches on the persistent P / dS tiles: against fp64 torch
    qr = qkv.double().requires_grad_(True)
    ref = _attn_ref(qr, km, R, L, D, H)
    ref.backward(dctx.double())
    ctx = torch.zeros(R * L, D, device=G.DEV, dtype=torch.bfloat16); dq = torch.zeros(R * L, 3 * D, device=G.DEV, dtype=torch.bfloat16)
    Lb.check(lib.clipdlm_attn_fwd(C.byref(G.bfp(qkv, None)), words.data_ptr(), R, L, D, H, C.byref(G.bfp(ctx, None)), 0, 0, 0.0, G.st()))
    Lb.check(lib.clipdlm_attn_bwd(C.byref(G.bfp(qkv, None)), words.data_ptr(), C.byref(G.bfp(dctx, None)), R, L, D, H, C.byref(G.bfp(dq, None)), 0, 0, 0.0, G.st()))
    assert rel(ctx, ref) < 5e-3 and rel(dq, qr.grad) < 6e-3


@pytest.mark.parametrize("L,R", [(40, 7), (66, 5), (128, 3)])
def test_attention_long_rows_umma_matches_simt_with_dropout(G, L, R):
    """32 < L <= 128, plain bf16: tcgen05 tiles with two / one sequence(s) per tile (the bert-large, seq_len 64 shape is L = 66). The
    fp32 SIMT kernels must see the same dropout mask, forward and backward."""
    D, H, p = 1024, 16, 0.1
    qkv = torch.randn(R * L, 3 * D, device=G.DEV).bfloat16()
    dctx = torch.randn(R * L, D, device=G.DEV).bfloat16()
    km = torch.rand(R, L, device=G.DEV) > 0.2
    km[:, 1] = True
    kw = (L + 31) // 32
    words = torch.zeros(R, kw, device=G.DEV, dtype=torch.int64)
    for j in range(L):
        words[:, j // 32] |= km[:, j].long() << (j % 32)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).contiguous()
    lib, Lb = G.lib(), G.L
    outs = {}
    for simt in (0, 1):
        lib.clipdlm_attn_force_simt(simt)
        ctx = torch.zeros(R * L, D, device=G.DEV, dtype=torch.bfloat16)
        dq = torch.zeros(R * L, 3 * D, device=G.DEV, dtype=torch.bfloat16)
        Lb.check(lib.clipdlm_attn_fwd(C.byref(G.bfp(qkv, None)), words.data_ptr(), R, L, D, H, C.byref(G.bfp(ctx, None)), 777, 5, p, G.st()))
        Lb.check(lib.clipdlm_attn_bwd(C.byref(G.bfp(qkv, None)), words.data_ptr(), C.byref(G.bfp(dctx, None)), R, L, D, H, C.byref(G.bfp(dq, None)),
                                      777, 5, p, G.st()))
        torch.cuda.synchronize()
        outs[simt] = (ctx.float(), dq.float())
    lib.clipdlm_attn_force_simt(0)
    assert rel(outs[0][0], outs[1][0]) < 1e-2 and rel(outs[0][1], outs[1][1]) < 1.5e-2
    # repeated launches on the persistent tiles (stale P / dS blocks of earlier groups must not leak): a second, different input
    qkv2 = torch.randn(R * L, 3 * D, device=G.DEV).bfloat16()
    ctx_a = torch.zeros(R * L, D, device=G.DEV, dtype=torch.bfloat16); ctx_b = torch.zeros_like(ctx_a)
    Lb.check(lib.clipdlm_attn_fwd(C.byref(G.bfp(qkv2, None)), words.data_ptr(), R, L, D, H, C.byref(G.bfp(ctx_a, None)), 0, 0, 0.0, G.st()))
    lib.clipdlm_attn_force_simt(1)
    Lb.check(lib.clipdlm_attn_fwd(C.byref(G.bfp(qkv2, None)), words.data_ptr(), R, L, D, H, C.byref(G.bfp(ctx_b, None)), 0, 0, 0.0, G.st()))
    lib.clipdlm_attn_force_simt(0)
    torch.cuda.synchronize()
    assert rel(ctx_a.float(), ctx_b.float()) < 1e-2


# ------------------------------------------------------------------------------------------------------------ embed / loss / misc
@pytest.mark.parametrize("fusion,guided", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("mode", [0, 1])
def test_embed_fwd_bwd(G, fusion, guided, mode):
    B, S, Ltxt, D, V = 5, 3, 16, 768, 1000
    Lf = Ltxt + 2 if fusion == 0 else Ltxt
    R = S * B
    E = torch.randn(V, D, device=G.DEV) * 0.02
    ids = torch.randint(0, V, (B, Ltxt), device=G.DEV, dtype=torch.int32)
    noise = torch.randn(B, Ltxt, D, device=G.DEV)
    ca = torch.rand(S, device=G.DEV); cb = torch.rand(S, device=G.DEV)
    img = torch.randn(B, D, device=G.DEV); txt = torch.randn(B, D, device=G.DEV)
    seg = torch.randn(2, D, device=G.DEV); pos = torch.randn(512, D, device=G.DEV) * 0.1
    w = torch.randn(D, device=G.DEV) * 0.1 + 1; b = torch.randn(D, device=G.DEV) * 0.1
    x_expl = (ca[:, None, None, None] * E[ids.long()][None] + cb[:, None, None, None] * noise[None]).reshape(R, Ltxt, D)
    if fusion == 0:
        z_ref = torch.cat([x_expl, img.repeat(S, 1)[:, None], txt.repeat(S, 1)[:, None]], 1) + seg[[0] * Ltxt + [1, 1]] + pos[:Lf]
    else:
        z_ref = x_expl + img.repeat(S, 1)[:, None] + (txt.repeat(S, 1)[:, None] if guided else 0) + pos[:Lf]
    e = G.L.Embed()
    e.R, e.B, e.Ltxt, e.L, e.D, e.fusion, e.mode, e.guided = R, B, Ltxt, Lf, D, fusion, mode, guided
    xin = x_expl.contiguous()
    if mode == 0:
        e.x_in = xin.data_ptr()
    else:
        e.emb_table, e.ids, e.noise, e.coef_a, e.coef_b = E.data_ptr(), ids.data_ptr(), noise.data_ptr(), ca.data_ptr(), cb.data_ptr()
    e.img_proj, e.txt_proj, e.seg, e.pos = img.data_ptr(), txt.data_ptr(), seg.data_ptr(), pos.data_ptr()
    e.ln_w, e.ln_b, e.ln_eps = w.data_ptr(), b.data_ptr(), 1e-12
    zh, zl = G.empty_pair((R * Lf, D), True); hh, hl = G.empty_pair((R * Lf, D), True)
    e.z, e.h = G.bfp(zh, zl), G.bfp(hh, hl)
    G.L.check(G.lib().clipdlm_embed_fwd(C.byref(e), G.st()))
    assert rel(G.join(zh, zl), z_ref.reshape(R * Lf, D)) < 1e-5
    assert rel(G.join(hh, hl), F.layer_norm(z_ref, (D,), w, b, 1e-12).reshape(R * Lf, D)) < 1e-5
    # backward of the fusion
    dz = torch.randn(R * Lf, D, device=G.DEV)
    dh, dl = G.split(dz, True)
    dzr = G.join(dh, dl).view(S, B, Lf, D).double()
    d_pos = torch.zeros(512, D, device=G.DEV); d_seg = torch.zeros(2, D, device=G.DEV)
    d_img = torch.zeros(B, D, device=G.DEV); d_txt = torch.zeros(B, D, device=G.DEV)
    G.L.check(G.lib().clipdlm_embed_bwd(C.byref(G.bfp(dh, dl)), R, B, Ltxt, Lf, D, fusion, guided, d_pos.data_ptr(), d_seg.data_ptr(),
                                        d_img.data_ptr(), d_txt.data_ptr(), G.st()))
    assert rel(d_pos[:Lf], dzr.sum((0, 1))) < 1e-5 and float(d_pos[Lf:].abs().max()) == 0
    if fusion == 0:
        assert rel(d_seg[0], dzr[:, :, :Ltxt].sum((0, 1, 2))) < 1e-5 and rel(d_seg[1], dzr[:, :, Ltxt:].sum((0, 1, 2))) < 1e-5
        assert rel(d_img, dzr[:, :, Ltxt].sum(0)) < 1e-5 and rel(d_txt, dzr[:, :, Ltxt + 1].sum(0)) < 1e-5
    else:
        assert rel(d_img, dzr.sum((0, 2))) < 1e-5
        if guided:
            assert rel(d_txt, dzr.sum((0, 2))) < 1e-5
        else:
            assert float(d_txt.abs().max()) == 0


@pytest.mark.parametrize("kind,name", [(0, "series_sum_sample_mean"), (1, "series_sum"), (2, "mse_series_mean"), (3, "mse_series_sum")])
def test_embed_loss(G, kind, name):
    B, S, Ltxt, Lf, D, V = 4, 3, 16, 18, 768, 500
    R = S * B
    hp = O.default_hparams(); hp["BATCH_SIZE"] = B
    E = torch.randn(V, D, device=G.DEV) * 0.5
    ids = torch.randint(0, V, (B, Ltxt), device=G.DEV, dtype=torch.int32)
    xo = torch.randn(R * Lf, D, device=G.DEV)
    xh, xl = G.split(xo, True)
    xr = G.join(xh, xl).view(R, Lf, D).clone().requires_grad_(True)
    ref = O.LOSS_FUNCS[name](xr[:, :Ltxt], E[ids.long()].repeat(S, 1, 1), hp)
    ref.backward()
    acc = torch.zeros(1, device=G.DEV, dtype=torch.float64)
    gh, gl = G.empty_pair((R * Lf, D), True)
    gh.fill_(7.0)
    G.L.check(G.lib().clipdlm_embed_loss(C.byref(G.bfp(xh, xl)), E.data_ptr(), ids.data_ptr(), None, 0, R, B, Ltxt, Lf, D, kind, R, B, 1.0,
                                         acc.data_ptr(), C.byref(G.bfp(gh, gl)), G.st()))
    assert abs(acc.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel(G.join(gh, gl).view(R, Lf, D), xr.grad) < 1e-5
    # explicit target tensor (x_{t-1}-prediction objective)
    tgt = torch.randn(R, Ltxt, D, device=G.DEV)
    xr2 = G.join(xh, xl).view(R, Lf, D).clone().requires_grad_(True)
    ref2 = O.LOSS_FUNCS[name](xr2[:, :Ltxt], tgt, hp)
    acc.zero_()
    G.L.check(G.lib().clipdlm_embed_loss(C.byref(G.bfp(xh, xl)), None, None, tgt.data_ptr(), R, R, B, Ltxt, Lf, D, kind, R, B, 1.0,
                                         acc.data_ptr(), None, G.st()))
    assert abs(acc.item() - ref2.item()) < 1e-5 * abs(ref2.item())


def test_colsum_small_linear_qsample_convert(G):
    lib, L = G.lib(), G.L
    x = torch.randn(5000, 3072, device=G.DEV)
    xh, xl = G.split(x, True)
    out = torch.full((3072,), 1.0, device=G.DEV)
    L.check(lib.clipdlm_colsum(C.byref(G.bfp(xh, xl)), 5000, 3072, out.data_ptr(), G.st()))
    assert rel(out - 1.0, G.join(xh, xl).double().sum(0)) < 1e-5
    B, K, N = 13, 512, 768
    xi = torch.randn(B, K, device=G.DEV); w = torch.randn(N, K, device=G.DEV) * 0.05; b = torch.randn(N, device=G.DEV)
    y = torch.zeros(B, N, device=G.DEV)
    L.check(lib.clipdlm_small_linear_fwd(xi.data_ptr(), w.data_ptr(), b.data_ptr(), B, K, N, y.data_ptr(), G.st()))
    assert rel(y, F.linear(xi, w, b)) < 1e-5
    dy = torch.randn(B, N, device=G.DEV); dw = torch.zeros(N, K, device=G.DEV); db = torch.zeros(N, device=G.DEV)
    L.check(lib.clipdlm_small_linear_bwd(xi.data_ptr(), dy.data_ptr(), B, K, N, dw.data_ptr(), db.data_ptr(), G.st()))
    assert rel(dw, dy.t() @ xi) < 1e-5 and rel(db, dy.sum(0)) < 1e-5
    x0 = torch.randn(4, 16, 768, device=G.DEV); nz = torch.randn(4, 16, 768, device=G.DEV)
    hp = O.default_hparams()
    t = torch.tensor([0, 1, 500, 999, 7], device=G.DEV)
    got = clipdlm_mod().diffuse_t(x0, t, hp, nz)
    ref = O.diffuse_t(x0.cpu(), t.cpu().reshape(5, 1, 1), O.alpha_cumprod(hp), nz.cpu())
    assert rel(got, ref) < 1e-6 and torch.equal(got[:4].cpu(), x0.cpu())
    f = torch.randn(1000, 768, device=G.DEV)
    hi = torch.zeros(1000, 768, device=G.DEV, dtype=torch.bfloat16); lo = torch.zeros_like(hi); back = torch.zeros_like(f)
    L.check(lib.clipdlm_to_bf16(f.data_ptr(), hi.data_ptr(), lo.data_ptr(), f.numel(), G.st()))
    L.check(lib.clipdlm_to_f32(hi.data_ptr(), lo.data_ptr(), back.data_ptr(), f.numel(), G.st()))
    assert torch.equal(hi, f.bfloat16()) and rel(back, f) < 1e-5
    g = torch.zeros(10 * 16, 768, device=G.DEV)
    src_h, src_l = G.split(torch.randn(10 * 18, 768, device=G.DEV), True)
    L.check(lib.clipdlm_gather_rows_f32(C.byref(G.bfp(src_h, src_l)), 160, 16, 18, 768, g.data_ptr(), G.st()))
    assert torch.equal(g.view(10, 16, 768), G.join(src_h, src_l).view(10, 18, 768)[:, :16])


def clipdlm_mod():
    import clipdlm
    return clipdlm


def test_adamw_matches_torch(G):
    n = 4096 * 3 + 8
    p0 = torch.randn(n, device=G.DEV)
    p = p0.clone(); m = torch.zeros(n, device=G.DEV); v = torch.zeros(n, device=G.DEV)
    hi = torch.zeros(n, device=G.DEV, dtype=torch.bfloat16); lo = torch.zeros_like(hi)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=1e-3)
    for step in range(1, 4):
        g = torch.randn(n, device=G.DEV) * (10.0 ** (step - 2))
        g[:100] = 0
        ref_p.grad = g.clone() / 2  # grad_scale = 0.5 below
        opt.step()
        gg = g.clone()
        G.L.check(G.lib().clipdlm_adamw(p.data_ptr(), gg.data_ptr(), m.data_ptr(), v.data_ptr(), hi.data_ptr(), lo.data_ptr(), n, 1e-3, 0.9,
                                        0.999, 1e-8, 0.01, step, 0.5, 1, G.st()))
        assert float(gg.abs().max()) == 0.0  # zero_grad fused
        assert rel(p, ref_p.detach()) < 2e-6
        assert rel(hi.float() + lo.float(), p) < 1e-5 and torch.equal(hi, p.bfloat16())


# ------------------------------------------------------------------------------------------------------------ TRAIN_EMBEDDING glue
@pytest.mark.parametrize("pair", [False, True])
def test_gemm_narrow_operands(G, pair):
    """The 64-channel-padded lm_head GEMMs of the TRAIN_EMBEDDING path: K = 64 forward, N = 64 dgrad (MN-major B, ragged K = V),
    N = 64 wgrad with a ragged M = V."""
    M, V, C64 = 200, 997, 64
    Vp = (V + 255) // 256 * 256
    a = torch.zeros(M, C64, device=G.DEV); a[:, :16] = torch.randn(M, 16, device=G.DEV)
    w = torch.zeros(Vp, C64, device=G.DEV); w[:V, :16] = torch.randn(V, 16, device=G.DEV) * 0.2
    ah, al = G.split(a, pair); wh, wl = G.split(w, pair)
    ar, wr = G.join(ah, al).double(), G.join(wh, wl).double()
    n32 = (V + 31) // 32 * 32
    out = torch.zeros(M, n32, device=G.DEV)
    G.gemm(a_hi=ah, a_lo=al, b_hi=wh, b_lo=wl, lda=C64, ldb=C64, M=M, N=n32, K=C64, epilogue=0, out_f32=out, ldo=n32)
    assert rel(out[:, :V], ar @ wr[:V].t()) < (3e-5 if pair else 4e-3)
    assert float(out[:, V:].abs().max()) == 0.0
    ldl = Vp
    dl = torch.zeros(M, ldl, device=G.DEV); dl[:, :V] = torch.randn(M, V, device=G.DEV) * 0.1
    dh, dll = G.split(dl, pair)
    dr = G.join(dh, dll).double()
    dce = torch.full((M, C64), 7.0, device=G.DEV)
    G.gemm(a_hi=dh, a_lo=dll, b_hi=wh, b_lo=wl, lda=ldl, ldb=C64, M=M, N=C64, K=V, a_major=0, b_major=1, epilogue=0, out_f32=dce, ldo=C64)
    assert rel(dce, dr[:, :V] @ wr[:V]) < (3e-5 if pair else 4e-3)
    assert float(dce[:, 16:].abs().max()) == 0.0
    acc = torch.zeros(Vp, C64, device=G.DEV); acc[:V] = 0.25
    G.gemm(a_hi=dh, a_lo=dll, b_hi=ah, b_lo=al, lda=ldl, ldb=C64, M=V, N=C64, K=M, a_major=1, b_major=1, epilogue=1, acc_f32=acc, ldo=C64)
    assert rel(acc[:V], dr[:, :V].t() @ ar + 0.25) < (3e-5 if pair else 4e-3)
    assert float(acc[V:].abs().max()) == 0.0 and float((acc[:V, 16:] - 0.25).abs().max()) == 0.0


@pytest.mark.parametrize("kind,name", [(0, "series_sum_sample_mean"), (1, "series_sum"), (2, "mse_series_mean"), (3, "mse_series_sum")])
def test_feature_loss_f32(G, kind, name):
    R, B, Ltxt, Lf, ch = 13, 4, 16, 18, 16
    hp = O.default_hparams(); hp.update(BATCH_SIZE=B)
    y = torch.randn(R, Lf, ch, device=G.DEV).requires_grad_(True)
    for trows in (B, R):
        tgt = torch.randn(trows, Ltxt, ch, device=G.DEV).requires_grad_(True)
        rep = tgt.repeat((R + trows - 1) // trows, 1, 1)[:R] if trows != R else tgt
        ref = O.LOSS_FUNCS[name](y[:, :Ltxt], rep, hp)
        if kind == 0:
            ref = ref  # mean over [R, ch]
        gy, gt = torch.autograd.grad(ref * 0.5, [y, tgt])
        dce = torch.randn(R * Ltxt, 64, device=G.DEV)
        dy = torch.full((R, Lf, ch), 9.0, device=G.DEV); dt = torch.zeros(trows, Ltxt, ch, device=G.DEV)
        acc = torch.zeros(1, device=G.DEV, dtype=torch.float64)
        G.L.check(G.lib().clipdlm_feature_loss_f32(y.data_ptr(), tgt.data_ptr(), trows, R, Ltxt, Lf, ch, kind, R, B, 0.5, acc.data_ptr(),
                                                   dce.data_ptr(), 64, dy.data_ptr(), dt.data_ptr(), G.st()))
        torch.cuda.synchronize()
        assert abs(acc.item() - ref.item()) < 1e-5 * abs(ref.item())   # value is unweighted, gradient carries the weight
        want = gy.clone(); want[:, :Ltxt] += dce.view(R, Ltxt, 64)[:, :, :ch]
        assert rel(dy, want) < 1e-5 and float(dy[:, Ltxt:].abs().max()) == 0.0
        assert rel(dt, gt) < 1e-5


def test_pack_rows_and_embedding_bwd(G):
    R, Ltxt, Lf, ch = 7, 16, 18, 16
    y = torch.randn(R, Lf, ch, device=G.DEV)
    for pair in (False, True):
        hi, lo = G.empty_pair((R * Ltxt, 64), pair)
        hi.fill_(3.0)
        G.L.check(G.lib().clipdlm_pack_rows_bf16(y.data_ptr(), R * Ltxt, Ltxt, Lf, ch, 64, hi.data_ptr(), G.L.ptr(lo), G.st()))
        got = G.join(hi, lo)
        assert rel(got[:, :ch], y[:, :Ltxt].reshape(-1, ch)) < (1e-5 if pair else 4e-3) and float(got[:, ch:].abs().max()) == 0.0
    S, B, V = 5, 3, 50
    ids = torch.randint(0, V, (B * Ltxt,), device=G.DEV, dtype=torch.int32)
    ids[:4] = 7  # collisions
    dx = torch.randn(S, B * Ltxt, ch, device=G.DEV); scale = torch.rand(S, device=G.DEV)
    dE = torch.full((V, ch), 0.5, device=G.DEV)
    G.L.check(G.lib().clipdlm_embedding_bwd(dx.data_ptr(), scale.data_ptr(), ids.data_ptr(), S, B * Ltxt, ch, dE.data_ptr(), G.st()))
    ref = torch.full((V, ch), 0.5, device=G.DEV, dtype=torch.float64)
    ref.index_add_(0, ids.long(), (dx.double() * scale.double()[:, None, None]).sum(0))
    assert rel(dE, ref) < 1e-6
    dE2 = torch.zeros(V, ch, device=G.DEV)
    G.L.check(G.lib().clipdlm_embedding_bwd(dx.data_ptr(), None, ids.data_ptr(), 1, B * Ltxt, ch, dE2.data_ptr(), G.st()))
    ref2 = torch.zeros(V, ch, device=G.DEV, dtype=torch.float64).index_add_(0, ids.long(), dx[0].double())
    assert rel(dE2, ref2) < 1e-6


@pytest.mark.parametrize("B,K,N", [(10007 * 4, 16, 768), (9001 * 4, 768, 16)])
def test_small_linear_long_reduction(G, B, K, N):
    """The TRAIN_EMBEDDING projections: forward over every token of a chunk, weight / bias gradients reduced over the tokens (split across
    blocks, fp32 atomics)."""
    x = torch.randn(B, K, device=G.DEV); w = torch.randn(N, K, device=G.DEV) * 0.1; b = torch.randn(N, device=G.DEV)
    y = torch.empty(B, N, device=G.DEV)
    G.L.check(G.lib().clipdlm_small_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), B, K, N, y.data_ptr(), G.st()))
    assert rel(y, x.double() @ w.double().t() + b.double()) < 1e-5
    dy = torch.randn(B, N, device=G.DEV)
    dw = torch.full((N, K), 0.5, device=G.DEV); db = torch.full((N,), -1.0, device=G.DEV)
    G.L.check(G.lib().clipdlm_small_linear_bwd(x.data_ptr(), dy.data_ptr(), B, K, N, dw.data_ptr(), db.data_ptr(), G.st()))
    assert rel(dw, dy.double().t() @ x.double() + 0.5) < 1e-5
    assert rel(db, dy.double().sum(0) - 1.0) < 1e-5


@pytest.mark.parametrize("pair", [False, True])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_adamw_dp_peer_path_emulated(G, world, pair):
    """clipdlm_adamw_dp without multicast: `world` sets of flat buffers on ONE device stand in for the ranks; every "rank" runs the
    fused kernel for its slice. All copies must end up holding AdamW(sum of the gradients) exactly as clipdlm_adamw computes it."""
    n = 8 * 12345 + 8   # not a multiple of 8 * world
    torch.manual_seed(3)
    p0 = torch.randn(n, device=G.DEV)
    gs = [torch.randn(n, device=G.DEV) * 0.1 for _ in range(world)]
    ps = [p0.clone() for _ in range(world)]
    his = [torch.zeros(n, device=G.DEV, dtype=torch.bfloat16) for _ in range(world)]
    los = [torch.zeros(n, device=G.DEV, dtype=torch.bfloat16) if pair else None for _ in range(world)]
    m_ref, v_ref = torch.rand(n, device=G.DEV) * 0.01, torch.rand(n, device=G.DEV) * 0.01
    ms, vs = m_ref.clone(), v_ref.clone()
    lib, Lb = G.lib(), G.L
    # reference: plain kernel on the summed gradient
    p_ref = p0.clone(); g_sum = torch.stack(gs).sum(0)
    if world == 3:
        g_sum = (gs[0] + gs[1]) + gs[2]
    hi_ref = torch.zeros(n, device=G.DEV, dtype=torch.bfloat16); lo_ref = torch.zeros_like(hi_ref) if pair else None
    Lb.check(lib.clipdlm_adamw(p_ref.data_ptr(), g_sum.clone().data_ptr(), m_ref.data_ptr(), v_ref.data_ptr(), hi_ref.data_ptr(), Lb.ptr(lo_ref), n,
                               1e-3, 0.9, 0.999, 1e-8, 0.01, 5, 1.0 / world, 0, G.st()))
    covered = torch.zeros(n, device=G.DEV, dtype=torch.int32)
    for rank in range(world):
        d = Lb.DpBuffers()
        d.rank, d.world = rank, world
        for r in range(world):
            d.p[r], d.g[r], d.shadow_hi[r] = ps[r].data_ptr(), gs[r].data_ptr(), his[r].data_ptr()
            d.shadow_lo[r] = los[r].data_ptr() if pair else None
        b, e = C.c_int64(), C.c_int64()
        Lb.check(lib.clipdlm_dp_slice(n, rank, world, C.byref(b), C.byref(e)))
        assert b.value % 8 == 0 and (e.value % 8 == 0 or e.value == n)
        covered[b.value:e.value] += 1
        Lb.check(lib.clipdlm_adamw_dp(C.byref(d), ms.data_ptr(), vs.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 0.01, 5, 1.0 / world, G.st()))
    torch.cuda.synchronize()
    assert bool((covered == 1).all())
    tol = 0.0 if world <= 2 else 1e-6   # the order of a 3-way sum is the kernel's (rank order starting at the owner)
    for r in range(world):
        assert float((ps[r] - p_ref).abs().max()) <= tol * float(p_ref.abs().max()) + 0.0 if world <= 2 else rel(ps[r], p_ref) < 1e-6
        if world <= 2:
            assert torch.equal(his[r], hi_ref)
            if pair:
                assert torch.equal(los[r], lo_ref)
    if world <= 2:
        assert torch.equal(ms, m_ref) and torch.equal(vs, v_ref)
