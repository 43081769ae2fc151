"""GPU bring-up harness for the tcgen05 GEMM (run on the B200 box via gpurun; not a pytest file).

Each case runs in its own subprocess so a trapped kernel (sticky CUDA error) cannot poison the others.
usage: python tests/gpu_bringup_gemm.py            # all cases
       python tests/gpu_bringup_gemm.py --case kk  # one case in-process
"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _load():
    import importlib.util
    spec = importlib.util.spec_from_file_location("clipdlm_lib", os.path.join(ROOT, "diffusion-image-captioning_b200", "_lib.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def rel_err(a, b):
    import torch
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def split(x):
    import torch
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def run_case(case):
    import torch
    L = _load()
    # bind only the GEMM symbols so this harness works while the rest of the library is being brought up
    lib = C.CDLL(L.LIB_PATH)
    for name in ("clipdlm_last_error", "clipdlm_gemm", "clipdlm_gemm_debug_mn_desc", "clipdlm_lse_combine"):
        fn = getattr(lib, name); fn.restype, fn.argtypes = L._SIGS[name]
    L._lib = lib
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    st = torch.cuda.current_stream().cuda_stream

    def gemm(**kw):
        g = L.Gemm()
        for k, v in kw.items():
            if hasattr(v, "data_ptr"):
                v = v.data_ptr()
            setattr(g, k, v)
        L.check(lib.clipdlm_gemm(C.byref(g), st))
        torch.cuda.synchronize()

    def mk(m, k, scale=1.0):
        return (torch.randn(m, k, device=dev) * scale)

    if case == "kk":
        for (M, N, K) in [(128, 256, 64), (256, 256, 128), (1000, 768, 768), (4096, 2304, 768), (300, 3072, 3072)]:
            a = mk(M, K).bfloat16(); b = mk(N, K, 0.05).bfloat16()
            out = torch.zeros(M, N, device=dev)
            gemm(a_hi=a, b_hi=b, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_f32=out, ldo=N)
            ref = a.float() @ b.float().t()
            print(f"[kk] {M}x{N}x{K} f32-out rel_err={rel_err(out, ref):.3e} max_abs={(out-ref).abs().max().item():.3e}")
            ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            gemm(a_hi=a, b_hi=b, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_hi=ob, ldo=N)
            print(f"[kk] {M}x{N}x{K} bf16-out rel_err={rel_err(ob.float(), ref):.3e}")
    elif case == "epi":
        M, N, K = 520, 768, 256
        a = mk(M, K).bfloat16(); b = mk(N, K, 0.05).bfloat16()
        bias = torch.randn(N, device=dev)
        res = mk(M, N).bfloat16()
        u = mk(M, N).bfloat16()
        ref = a.float() @ b.float().t()
        out = torch.zeros(M, N, device=dev)
        gemm(a_hi=a, b_hi=b, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_f32=out, ldo=N, bias=bias)
        print(f"[epi] bias rel_err={rel_err(out, ref + bias):.3e}")
        gemm(a_hi=a, b_hi=b, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_f32=out, ldo=N, bias=bias, res_hi=res, ldr=N)
        print(f"[epi] bias+res rel_err={rel_err(out, ref + bias + res.float()):.3e}")
        gemm(a_hi=a, b_hi=b, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_f32=out, ldo=N, u_hi=u, ldu=N)
        uf = u.float().requires_grad_(True)
        torch.nn.functional.gelu(uf).sum().backward()
        print(f"[epi] dgelu rel_err={rel_err(out, ref * uf.grad):.3e}")
        o1 = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); o2 = torch.zeros_like(o1)
        gemm(a_hi=a, b_hi=b, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_hi=o1, out2_hi=o2, ldo=N, bias=bias)
        print(f"[epi] gelu-dual pre rel_err={rel_err(o1.float(), ref + bias):.3e} post rel_err={rel_err(o2.float(), torch.nn.functional.gelu(ref + bias)):.3e}")
        # dropout: kept fraction, scaling, determinism
        gemm(a_hi=a, b_hi=b, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_f32=out, ldo=N, bias=bias, drop_seed=1234, drop_site=7, drop_p=0.1)
        kept = (out != 0).float().mean().item()
        mask = out != 0
        print(f"[epi] dropout kept={kept:.4f} (expect 0.9) scaled rel_err={rel_err(out[mask], ((ref + bias) / 0.9)[mask]):.3e}")
        out2 = torch.zeros_like(out)
        gemm(a_hi=a, b_hi=b, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_f32=out2, ldo=N, bias=bias, drop_seed=1234, drop_site=7, drop_p=0.1)
        print(f"[epi] dropout deterministic={torch.equal(out, out2)}")
        # scatter rows: logical row m -> (m/16)*18 + m%16
        M2 = 512
        a2 = mk(M2, K).bfloat16()
        outs = torch.zeros(M2 // 16 * 18, N, device=dev, dtype=torch.bfloat16)
        ress = mk(M2 // 16 * 18, N).bfloat16()
        gemm(a_hi=a2, b_hi=b, lda=K, ldb=K, M=M2, N=N, K=K, epilogue=L.EPI_STORE, out_hi=outs, ldo=N, res_hi=ress, ldr=N, scatter_len=16, scatter_stride=18)
        ref2 = a2.float() @ b.float().t()
        got = outs.view(-1, 18, N)[:, :16].reshape(M2, N).float()
        exp = ref2 + ress.view(-1, 18, N)[:, :16].reshape(M2, N).float()
        print(f"[epi] scatter rel_err={rel_err(got, exp):.3e} untouched_rows_zero={(outs.view(-1, 18, N)[:, 16:] == 0).all().item()}")
        # gather A rows (3-D tensor map)
        R = 40
        xa = mk(R * 18, K).bfloat16()
        outg = torch.zeros(R * 16, N, device=dev)
        gemm(a_hi=xa, b_hi=b, lda=K, ldb=K, M=R * 16, N=N, K=K, gather_len=16, gather_stride=18, epilogue=L.EPI_STORE, out_f32=outg, ldo=N)
        refg = xa.view(R, 18, K)[:, :16].reshape(R * 16, K).float() @ b.float().t()
        print(f"[epi] gather rel_err={rel_err(outg, refg):.3e}")
    elif case == "split":
        M, N, K = 512, 768, 768
        a = mk(M, K); b = mk(N, K, 0.05)
        ah, al = split(a); bh, bl = split(b)
        out = torch.zeros(M, N, device=dev)
        gemm(a_hi=ah, a_lo=al, b_hi=bh, b_lo=bl, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_f32=out, ldo=N)
        ref = (a.double() @ b.double().t())
        print(f"[split] bf16x3 rel_err={rel_err(out, ref):.3e} (plain bf16 would be ~3e-3)")
        oh = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); ol = torch.zeros_like(oh)
        gemm(a_hi=ah, a_lo=al, b_hi=bh, b_lo=bl, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_hi=oh, out_lo=ol, ldo=N)
        print(f"[split] pair-out rel_err={rel_err(oh.float() + ol.float(), ref):.3e}")
    elif case in ("dgrad", "wgrad"):
        combos = [(0, 0)]
        if "--sweep" in sys.argv:
            combos = [(0, 0), (1024, 8192), (8192, 128), (128, 8192), (16, 1024), (1024, 16), (8192, 2048), (2048, 8192)]
        for (lbo, sbo) in combos:
            lib.clipdlm_gemm_debug_mn_desc(lbo, sbo)
            if case == "dgrad":
                # dX[M, Kout] = dY[M, Nred] @ W[Nred, Kout] : A K-major (dY), B MN-major (W stored [Nred][Kout])
                for (M, Nred, Kout) in [(256, 64, 256), (512, 768, 768), (1000, 3072, 768)]:
                    dy = mk(M, Nred).bfloat16(); w = mk(Nred, Kout, 0.05).bfloat16()
                    out = torch.zeros(M, Kout, device=dev)
                    gemm(a_hi=dy, b_hi=w, lda=Nred, ldb=Kout, M=M, N=Kout, K=Nred, a_major=0, b_major=1, epilogue=L.EPI_STORE, out_f32=out, ldo=Kout)
                    ref = dy.float() @ w.float()
                    print(f"[dgrad lbo={lbo} sbo={sbo}] {M}x{Kout}x{Nred} rel_err={rel_err(out, ref):.3e}")
            else:
                # dW[Nout, Kin] = dY[T, Nout]^T @ X[T, Kin] : both MN-major, reduction over T
                for (T, Nout, Kin, ks) in [(64, 128, 256, 1), (1024, 768, 768, 0), (5000, 3072, 768, 0), (4096, 768, 3072, 3)]:
                    dy = mk(T, Nout).bfloat16(); x = mk(T, Kin, 0.1).bfloat16()
                    acc = torch.ones(Nout, Kin, device=dev)
                    gemm(a_hi=dy, b_hi=x, lda=Nout, ldb=Kin, M=Nout, N=Kin, K=T, a_major=1, b_major=1, epilogue=L.EPI_WGRAD, acc_f32=acc, ldo=Kin, k_splits=ks)
                    ref = dy.float().t() @ x.float() + 1.0
                    print(f"[wgrad lbo={lbo} sbo={sbo}] T={T} {Nout}x{Kin} ks={ks} rel_err={rel_err(acc, ref):.3e}")
    elif case == "lse":
        R, Ltxt, Lall, K, V = 24, 16, 18, 768, 30522
        x = mk(R * Lall, K).bfloat16()
        E = mk(V, K, 0.05).bfloat16()
        M = R * Ltxt
        B = 8
        ids = torch.randint(0, V, (B * Ltxt,), device=dev, dtype=torch.int32)
        nt = (V + 255) // 256
        pm = torch.zeros(nt, M, device=dev); ps = torch.zeros(nt, M, device=dev); pa = torch.zeros(nt, M, device=dev, dtype=torch.int32)
        tl = torch.zeros(M, device=dev)
        gemm(a_hi=x, b_hi=E, lda=K, ldb=K, M=M, N=V, K=K, gather_len=Ltxt, gather_stride=Lall, epilogue=L.EPI_LSE,
             part_max=pm, part_sum=ps, part_arg=pa, tgt_logit=tl, targets=ids, tgt_period=B * Ltxt)
        lse = torch.zeros(M, device=dev); am = torch.zeros(M, device=dev, dtype=torch.int32)
        acc = torch.zeros(1, device=dev, dtype=torch.float64)
        L.check(lib.clipdlm_lse_combine(pm.data_ptr(), ps.data_ptr(), pa.data_ptr(), nt, M, tl.data_ptr(), lse.data_ptr(), am.data_ptr(), acc.data_ptr(), 1.0 / R, st))
        torch.cuda.synchronize()
        logits = x.view(R, Lall, K)[:, :Ltxt].reshape(M, K).float() @ E.float().t()
        ref_lse = torch.logsumexp(logits, -1)
        tgt = ids.long().repeat(R // B)
        ref_loss = (ref_lse - logits.gather(1, tgt[:, None]).squeeze(1)).sum() / R
        print(f"[lse] lse rel_err={rel_err(lse, ref_lse):.3e} argmax_match={(am.long() == logits.argmax(-1)).float().mean().item():.4f} loss={acc.item():.6f} ref={ref_loss.item():.6f}")
        # softmax-grad epilogue
        ldo = nt * 256
        dl = torch.full((M, ldo), 7.0, device=dev, dtype=torch.bfloat16)
        gemm(a_hi=x, b_hi=E, lda=K, ldb=K, M=M, N=V, K=K, gather_len=Ltxt, gather_stride=Lall, epilogue=L.EPI_SMGRAD,
             out_hi=dl, ldo=ldo, lse=lse, targets=ids, tgt_period=B * Ltxt, grad_scale=0.5)
        refg = torch.softmax(logits, -1)
        refg[torch.arange(M, device=dev), tgt] -= 1
        refg *= 0.5
        print(f"[smgrad] rel_err={rel_err(dl[:, :V].float(), refg):.3e} pad_zero={(dl[:, V:] == 0).all().item()}")
    elif case == "perf":
        for (M, N, K, tag) in [(92160, 2304, 768, "qkv"), (92160, 3072, 768, "ffn1"), (92160, 768, 3072, "ffn2"), (81920, 30522, 768, "lm_head-lse")]:
            a = mk(M, K).bfloat16(); b = mk(N, K, 0.05).bfloat16()
            if tag == "lm_head-lse":
                nt = (N + 255) // 256
                pm = torch.zeros(nt, M, device=dev); ps = torch.zeros(nt, M, device=dev); pa = torch.zeros(nt, M, device=dev, dtype=torch.int32)
                tl = torch.zeros(M, device=dev); ids = torch.zeros(M, device=dev, dtype=torch.int32)
                kw = dict(a_hi=a, b_hi=b, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_LSE, part_max=pm, part_sum=ps, part_arg=pa, tgt_logit=tl, targets=ids, tgt_period=M)
            else:
                out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
                bias = torch.zeros(N, device=dev)
                kw = dict(a_hi=a, b_hi=b, lda=K, ldb=K, M=M, N=N, K=K, epilogue=L.EPI_STORE, out_hi=out, ldo=N, bias=bias)
            for _ in range(3):
                gemm(**kw)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            g = L.Gemm()
            for k, v in kw.items():
                setattr(g, k, v.data_ptr() if hasattr(v, "data_ptr") else v)
            e0.record()
            for _ in range(10):
                lib.clipdlm_gemm(C.byref(g), st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"[perf] {tag} {M}x{N}x{K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
            t0 = time.time()
            for _ in range(10):
                ref = a @ b.t()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                ref = a @ b.t()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"[perf] cuBLAS {tag}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
            del ref
        # wgrad + dgrad perf
        T, Nout, Kin = 92160, 3072, 768
        dy = mk(T, Nout).bfloat16(); x = mk(T, Kin).bfloat16(); w = mk(Nout, Kin, 0.05).bfloat16()
        acc = torch.zeros(Nout, Kin, device=dev)
        outd = torch.zeros(T, Kin, device=dev, dtype=torch.bfloat16)
        for name, kw in [("wgrad", dict(a_hi=dy, b_hi=x, lda=Nout, ldb=Kin, M=Nout, N=Kin, K=T, a_major=1, b_major=1, epilogue=L.EPI_WGRAD, acc_f32=acc, ldo=Kin)),
                         ("dgrad", dict(a_hi=dy, b_hi=w, lda=Nout, ldb=Kin, M=T, N=Kin, K=Nout, a_major=0, b_major=1, epilogue=L.EPI_STORE, out_hi=outd, ldo=Kin))]:
            for _ in range(3):
                gemm(**kw)
            g = L.Gemm()
            for k, v in kw.items():
                setattr(g, k, v.data_ptr() if hasattr(v, "data_ptr") else v)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                lib.clipdlm_gemm(C.byref(g), st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"[perf] {name} T={T} {Nout}x{Kin}: {ms:.3f} ms  {2.0 * T * Nout * Kin / ms / 1e9:.1f} TFLOP/s")
    else:
        raise SystemExit(f"unknown case {case}")


def main():
    if "--case" in sys.argv:
        run_case(sys.argv[sys.argv.index("--case") + 1])
        return
    extra = [a for a in sys.argv[1:] if a.startswith("--")]
    for case in ["kk", "epi", "split", "dgrad", "wgrad", "lse", "perf"]:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", case, *extra], capture_output=True, text=True, timeout=240)
            out = r.stdout + ("\n[stderr] " + r.stderr[-3000:] if r.returncode != 0 else "")
            print(f"===== case {case}: rc={r.returncode} ({time.time() - t0:.1f}s)\n{out}", flush=True)
        except subprocess.TimeoutExpired as ex:
            print(f"===== case {case}: TIMEOUT\n{(ex.stdout or b'').decode() if isinstance(ex.stdout, bytes) else ex.stdout}", flush=True)


if __name__ == "__main__":
    main()
