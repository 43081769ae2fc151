#!/usr/bin/env python
"""Micro-benchmark of the attention kernels at the train-chunk shape (R = 8192 sequences x L = 18, 12 heads): paths 0 (tcgen05, back-to-back
packed sequences; backward also with the folded bias gradients), 3 (tcgen05, 32-row slots), 2 (mma.sync TMA ring), with and without dropout. Triage tool.
CLIPDLM_ATTN_BWD_CONSUMERS=9..11 overrides the consumer-warp count of the TMA-ring backward (default 10).
ATTN_PATHS=0,4 restricts the paths; SEQ / DIM / ROWS env vars select other shapes (SEQ > 32: the one- / two-sequence-per-tile tcgen05 kernels, e.g. SEQ=66 DIM=1024)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import clipdlm  # noqa: E402,F401
from clipdlm import _lib as L  # noqa: E402

DEV = "cuda:0"
R, Ls, D = int(os.environ.get("ROWS", 8192)), int(os.environ.get("SEQ", 18)), int(os.environ.get("DIM", 768))
H = D // 64
lib = L.load()
st = torch.cuda.current_stream().cuda_stream
qkv = torch.randn(R * Ls, 3 * D, device=DEV).bfloat16()
dctx = torch.randn(R * Ls, D, device=DEV).bfloat16()
ctx = torch.empty(R * Ls, D, device=DEV, dtype=torch.bfloat16)
dqkv = torch.empty(R * Ls, 3 * D, device=DEV, dtype=torch.bfloat16)
KW = (Ls + 31) // 32
km = torch.full((R, KW), -1, device=DEV, dtype=torch.int32)  # every key visible (bits >= L are ignored)
if Ls <= 32:
    km = torch.full((R,), (1 << 17) - 1, device=DEV, dtype=torch.int32)
bq, bc, bd, bg = L.Bf(qkv.data_ptr(), None), L.Bf(ctx.data_ptr(), None), L.Bf(dctx.data_ptr(), None), L.Bf(dqkv.data_ptr(), None)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


fwd_bytes, bwd_bytes = R * Ls * D * 2 * 4, R * Ls * D * 2 * 7
dbias = torch.zeros(3 * D, device=DEV)
folded = (C.c_int32 * 1)()
PATHS = tuple(int(x) for x in os.environ["ATTN_PATHS"].split(",")) if os.environ.get("ATTN_PATHS") else None
for path in (PATHS or ((0, 4, 3, 2) if Ls <= 32 else (0,))):
    lib.clipdlm_attn_force_simt(path)
    for p in (0.0, 0.1):
        f = timeit(lambda: L.check(lib.clipdlm_attn_fwd(C.byref(bq), km.data_ptr(), R, Ls, D, H, C.byref(bc), 1, 1, p, st)))
        b = timeit(lambda: L.check(lib.clipdlm_attn_bwd(C.byref(bq), km.data_ptr(), C.byref(bd), R, Ls, D, H, C.byref(bg), 1, 1, p, st)))
        extra = ""
        if path in (0, 4) and Ls <= 32:
            bb = timeit(lambda: L.check(lib.clipdlm_attn_bwd_bias(C.byref(bq), km.data_ptr(), C.byref(bd), R, Ls, D, H, C.byref(bg), 1, 1, p, dbias.data_ptr(),
                                                                  C.cast(folded, C.c_void_p), st)))
            extra = f"   bwd + folded bias gradients {bb:7.1f} us (folded = {folded[0]})"
        print(f"path {path} p={p}: fwd {f:7.1f} us ({fwd_bytes / f / 1e3:6.0f} GB/s)   bwd {b:7.1f} us ({bwd_bytes / b / 1e3:6.0f} GB/s){extra}", flush=True)
lib.clipdlm_attn_force_simt(0)
