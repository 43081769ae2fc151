"""Thin test-side wrappers over the raw C-ABI kernels (ctypes), torch tensors in / out."""
import ctypes as C

import torch

import clipdlm  # noqa: F401  (import alias of the package)
from clipdlm import _lib as L

DEV = "cuda:0"


def lib():
    return L.load()


def st():
    return torch.cuda.current_stream().cuda_stream


def split(x, pair):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16) if pair else None
    return hi, lo


def join(hi, lo):
    return hi.float() + (lo.float() if lo is not None else 0)


def empty_pair(shape, pair):
    hi = torch.zeros(shape, device=DEV, dtype=torch.bfloat16)
    lo = torch.zeros(shape, device=DEV, dtype=torch.bfloat16) if pair else None
    return hi, lo


def bfp(hi, lo):
    return L.Bf(L.ptr(hi), L.ptr(lo))


def gemm(**kw):
    g = L.Gemm()
    keep = []
    for k, v in kw.items():
        if hasattr(v, "data_ptr"):
            keep.append(v)
            v = v.data_ptr()
        setattr(g, k, v)
    L.check(lib().clipdlm_gemm(C.byref(g), st()))
    torch.cuda.synchronize()


def tol(pair, tight=2e-5, loose=1e-2):
    return tight if pair else loose
