#!/usr/bin/env python
"""Micro-benchmark of the GEMM shapes one train chunk launches (T = 4096 rows x 18 positions = 73 728 tokens), through the C-ABI,
next to torch.matmul (cuBLAS) on the bare shape. Triage tool, not a bench line:  python tools/gemm_perf.py [--flags 0,1,2,4] [--rows 4096]

flags: clipdlm_gemm_debug_flags bits (1 = no epilogue stores, 2 = no aux loads, 4 = TMEM drain only, 4096 = no band tile order in the lm_head passes)."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import clipdlm  # noqa: E402,F401
from clipdlm import _lib as L  # noqa: E402

DEV = "cuda:0"


def bf(*shape):
    return (torch.randn(*shape, device=DEV) * 0.05).to(torch.bfloat16)


def time_fn(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def make(keep, **kw):
    g = L.Gemm()
    for k, v in kw.items():
        if hasattr(v, "data_ptr"):
            keep.append(v)
            v = v.data_ptr()
        setattr(g, k, v)
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--flags", default="0")
    ap.add_argument("--only", default="")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--no-cublas", action="store_true")
    args = ap.parse_args()
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    T = args.rows * 18
    M16 = args.rows * 16
    D, F, V = 768, 3072, 30522
    VP = (V + 255) // 256 * 256
    keep = []
    x768, x3072, x2304 = bf(T, D), bf(T, F), bf(T, 3 * D)
    o768, o3072, o2304, o3072b = bf(T, D), bf(T, F), bf(T, 3 * D), bf(T, F)
    wqkv, wo, w1, w2 = bf(3 * D, D), bf(D, D), bf(F, D), bf(D, F)
    E = bf(VP, D)
    bias = torch.randn(F, device=DEV)
    accs = {n: torch.zeros(s, device=DEV) for n, s in (("qkv", (3 * D, D)), ("o", (D, D)), ("f1", (F, D)), ("f2", (D, F)))}
    nt = VP // 256
    pm = torch.zeros(2 * nt, M16, device=DEV); ps = torch.zeros_like(pm); pa = torch.zeros(2 * nt, M16, device=DEV, dtype=torch.int32)
    tl = torch.zeros(M16, device=DEV); lse = torch.full((M16,), 5.0, device=DEV)
    tgt = torch.randint(0, V, (512 * 16,), device=DEV, dtype=torch.int32)
    dlog = torch.empty(M16, VP, device=DEV, dtype=torch.bfloat16)

    cases = [
        # name, flops, gemm desc, cuBLAS equivalent
        ("fwd qkv  bias", 2 * T * 3 * D * D, dict(a_hi=x768, b_hi=wqkv, lda=D, ldb=D, M=T, N=3 * D, K=D, out_hi=o2304, ldo=3 * D, bias=bias), (x768, wqkv.t())),
        ("fwd o    bias+res", 2 * T * D * D, dict(a_hi=x768, b_hi=wo, lda=D, ldb=D, M=T, N=D, K=D, out_hi=o768, ldo=D, bias=bias, res_hi=x768, ldr=D), (x768, wo.t())),
        ("fwd ffn1 bias+gelu dual", 2 * T * F * D, dict(a_hi=x768, b_hi=w1, lda=D, ldb=D, M=T, N=F, K=D, out_hi=o3072, out2_hi=o3072b, ldo=F, bias=bias), (x768, w1.t())),
        ("fwd ffn1 bias only", 2 * T * F * D, dict(a_hi=x768, b_hi=w1, lda=D, ldb=D, M=T, N=F, K=D, out_hi=o3072, ldo=F, bias=bias), None),
        ("fwd ffn2 bias+drop+res", 2 * T * F * D, dict(a_hi=x3072, b_hi=w2, lda=F, ldb=F, M=T, N=D, K=F, out_hi=o768, ldo=D, bias=bias, res_hi=x768, ldr=D, drop_seed=7, drop_site=3, drop_p=0.1), (x3072, w2.t())),
        ("fwd vt   bias+gelu dual", 2 * T * D * D, dict(a_hi=x768, b_hi=wo, lda=D, ldb=D, M=T, N=D, K=D, out_hi=o768, out2_hi=bf(T, D), ldo=D, bias=bias), None),
        ("dgrad ffn2 *gelu'(u)", 2 * T * F * D, dict(a_hi=x768, b_hi=w2, lda=D, ldb=F, M=T, N=F, K=D, b_major=1, out_hi=o3072, ldo=F, u_hi=x3072, ldu=F), (x768, w2)),
        ("EPI6 fwd ffn1 gelu'+gelu", 2 * T * F * D, dict(a_hi=x768, b_hi=w1, lda=D, ldb=D, M=T, N=F, K=D, epilogue=6, out_hi=o3072, out2_hi=o3072b, ldo=F, bias=bias), None),
        ("EPI7 dgrad ffn2 *stored gelu'", 2 * T * F * D, dict(a_hi=x768, b_hi=w2, lda=D, ldb=F, M=T, N=F, K=D, b_major=1, epilogue=7, out_hi=o3072, ldo=F, u_hi=x3072, ldu=F), None),
        ("EPI7 ... + bias colsum", 2 * T * F * D, dict(a_hi=x768, b_hi=w2, lda=D, ldb=F, M=T, N=F, K=D, b_major=1, epilogue=7, out_hi=o3072, ldo=F, u_hi=x3072, ldu=F, acc_f32=torch.zeros(F, device=DEV)), None),
        ("dgrad ffn1 +res", 2 * T * F * D, dict(a_hi=x3072, b_hi=w1, lda=F, ldb=D, M=T, N=D, K=F, b_major=1, out_hi=o768, ldo=D, res_hi=x768, ldr=D), (x3072, w1)),
        ("dgrad qkv +res", 2 * T * 3 * D * D, dict(a_hi=x2304, b_hi=wqkv, lda=3 * D, ldb=D, M=T, N=D, K=3 * D, b_major=1, out_hi=o768, ldo=D, res_hi=x768, ldr=D), (x2304, wqkv)),
        ("dgrad o", 2 * T * D * D, dict(a_hi=x768, b_hi=wo, lda=D, ldb=D, M=T, N=D, K=D, b_major=1, out_hi=o768, ldo=D), (x768, wo)),
        ("wgrad ffn2", 2 * T * F * D, dict(a_hi=x768, b_hi=x3072, lda=D, ldb=F, M=D, N=F, K=T, a_major=1, b_major=1, epilogue=1, acc_f32=accs["f2"], ldo=F), (x768.t(), x3072)),
        ("wgrad ffn1", 2 * T * F * D, dict(a_hi=x3072, b_hi=x768, lda=F, ldb=D, M=F, N=D, K=T, a_major=1, b_major=1, epilogue=1, acc_f32=accs["f1"], ldo=D), (x3072.t(), x768)),
        ("wgrad qkv", 2 * T * 3 * D * D, dict(a_hi=x2304, b_hi=x768, lda=3 * D, ldb=D, M=3 * D, N=D, K=T, a_major=1, b_major=1, epilogue=1, acc_f32=accs["qkv"], ldo=D), (x2304.t(), x768)),
        ("wgrad o", 2 * T * D * D, dict(a_hi=x768, b_hi=x768, lda=D, ldb=D, M=D, N=D, K=T, a_major=1, b_major=1, epilogue=1, acc_f32=accs["o"], ldo=D), (x768.t(), x768)),
        ("lm_head LSE", 2 * M16 * V * D, dict(a_hi=x768, b_hi=E, lda=D, ldb=D, M=M16, N=V, K=D, gather_len=16, gather_stride=18, epilogue=2, part_max=pm, part_sum=ps, part_arg=pa, tgt_logit=tl, targets=tgt, tgt_period=512 * 16), None),
        ("lm_head LSE_EXP + bf16 exp out", 2 * M16 * V * D, dict(a_hi=x768, b_hi=E, lda=D, ldb=D, M=M16, N=V, K=D, gather_len=16, gather_stride=18, epilogue=4, part_max=pm, part_sum=ps, tgt_logit=tl, targets=tgt, tgt_period=512 * 16, out_hi=dlog, ldo=VP), None),
        ("lm_head SMGRAD", 2 * M16 * V * D, dict(a_hi=x768, b_hi=E, lda=D, ldb=D, M=M16, N=V, K=D, gather_len=16, gather_stride=18, epilogue=3, out_hi=dlog, ldo=VP, lse=lse, targets=tgt, tgt_period=512 * 16, grad_scale=1e-3), None),
        ("lm_head dgrad scatter+res", 2 * M16 * V * D, dict(a_hi=dlog, b_hi=E, lda=VP, ldb=D, M=M16, N=D, K=V, b_major=1, out_hi=o768, ldo=D, res_hi=o768, ldr=D, scatter_len=16, scatter_stride=18), None),
    ]
    flag_list = [int(f) for f in args.flags.split(",")]
    print(f"T = {T}; columns: debug flags {flag_list} (min of {args.rounds} interleaved rounds), ms and TFLOP/s")
    for name, flops, kw, cb in cases:
        if args.only and args.only not in name:
            continue
        g = make(keep, **kw)

        def run(g=g):
            L.check(lib.clipdlm_gemm(C.byref(g), st))
        best = {f: 1e9 for f in flag_list}
        best_c = 1e9
        for _ in range(args.rounds):
            for f in flag_list:
                lib.clipdlm_gemm_debug_flags(f)
                best[f] = min(best[f], time_fn(run, iters=10, warm=2))
            if cb is not None and not args.no_cublas:
                a, b = cb
                best_c = min(best_c, time_fn(lambda: torch.matmul(a, b), iters=10, warm=2))
        line = f"{name:28s}" + "".join(f" | {best[f]:6.3f} {flops / best[f] / 1e9:6.0f}" for f in flag_list)
        if best_c < 1e9:
            line += f" || cuBLAS {best_c:6.3f} {flops / best_c / 1e9:6.0f}"
        print(line, flush=True)
    lib.clipdlm_gemm_debug_flags(0)


if __name__ == "__main__":
    main()
