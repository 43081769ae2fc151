"""Data parallelism for the train step: one process per GPU (torchrun), identical weights, captions sharded across ranks,
ONE collective per step — a sum all-reduce of the flat fp32 gradient buffer over NCCL (NVLink 5 / NVSwitch; NVLS in-switch
reduction when available) — followed by the same fused AdamW on every rank (which applies the 1/world scale).

The reference has no distributed code (single `cuda:0`, CLIP-DDPM.py:21); this is the data-parallel shape BASELINE.json's
north star asks for. `t` is broadcast from rank 0 each step because the reference shares one `t` draw across the whole batch
(:461). Works with the gloo backend on CPU tensors for the host-logic tests (tests/test_parallel_gloo.py).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def init_process_group_from_env(backend: Optional[str] = None) -> tuple:
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT). Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(n_items: int, rank: int, world: int) -> tuple:
    """Contiguous [lo, hi) slice of n_items owned by `rank` (captions of a global batch, images of a sampling job)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_flat(flat: torch.Tensor, group=None, src: int = 0):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean all-reduce of a flat buffer (used by tests; the train step all-reduces a sum and folds 1/world into AdamW)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
        flat.div_(dist.get_world_size(group))
    return flat


def _enable_fused_step(model, group):
    """Move the model's flat buffers into symmetric memory (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into
    every rank of the node, with an NVSwitch multicast object when the fabric supports it) and describe them to
    clipdlm_adamw_dp. PyTorch is plumbing here: allocation, handle exchange and the device-side barrier; the data path is ours."""
    import ctypes as C

    import torch.distributed._symmetric_memory as symm_mem

    from . import _lib as L
    grp = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(grp), dist.get_rank(grp)
    if world > L.MAX_PEERS:
        raise RuntimeError(f"fused data-parallel step supports up to {L.MAX_PEERS} GPUs of one node")
    n = model.n_params
    names = [("flat", torch.float32), ("grad", torch.float32), ("shadow_hi", torch.bfloat16)]
    if model.shadow_lo is not None:
        names.append(("shadow_lo", torch.bfloat16))
    tensors, handles = {}, {}
    for name, dtype in names:
        t = symm_mem.empty(n, dtype=dtype, device=model.device)
        handles[name] = symm_mem.rendezvous(t, grp)
        tensors[name] = t
    model._rebind_buffers(tensors["flat"], tensors["grad"], tensors["shadow_hi"], tensors.get("shadow_lo"))
    bufs = L.DpBuffers()
    bufs.rank, bufs.world = rank, world
    for r in range(world):
        bufs.p[r] = int(handles["flat"].buffer_ptrs[r])
        bufs.g[r] = int(handles["grad"].buffer_ptrs[r])
        bufs.shadow_hi[r] = int(handles["shadow_hi"].buffer_ptrs[r])
        bufs.shadow_lo[r] = int(handles["shadow_lo"].buffer_ptrs[r]) if "shadow_lo" in handles else None
    mc = all(h.has_multicast_support and int(h.multicast_ptr) != 0 for h in handles.values()) and os.environ.get("CLIPDLM_DP_MULTICAST", "1") != "0"
    if mc:
        bufs.p_mc, bufs.g_mc = int(handles["flat"].multicast_ptr), int(handles["grad"].multicast_ptr)
        bufs.shadow_hi_mc = int(handles["shadow_hi"].multicast_ptr)
        bufs.shadow_lo_mc = int(handles["shadow_lo"].multicast_ptr) if "shadow_lo" in handles else None
    b, e = C.c_int64(), C.c_int64()
    L.check(L.load().clipdlm_dp_slice(n, rank, world, C.byref(b), C.byref(e)))
    gh = handles["grad"]
    model.dp_fused = dict(bufs=bufs, handles=handles, slice=(int(b.value), int(e.value)), multicast=bool(mc),
                          barrier=lambda channel: gh.barrier(channel=channel))


def enable_data_parallel(model, group=None, fused=None):
    """Make every rank start from rank 0's weights (trainable flat buffer + frozen embedding / lm_head) and switch the
    model's train_func to DP mode (shared t + gradient exchange).

    fused=True: the exchange is part of the optimizer kernel (reduce-scatter + AdamW on the rank's slice + all-gather of the new weights
    over NVLink peer memory / NVSwitch multicast, csrc/dp_fused.cu) - create the AdamW AFTER this call. fused=False: NCCL sum all-reduce
    of the flat gradient buffer, then the full-size AdamW on every rank. fused=None: environment CLIPDLM_DP_FUSED (default "0")."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return model
    if fused is None:
        fused = os.environ.get("CLIPDLM_DP_FUSED", "0") == "1"
    if fused:
        _enable_fused_step(model, group)
    broadcast_flat(model.flat, group)
    if not model.hp["TRAIN_EMBEDDING"]:  # (there, embedding / lm_head are trainable slots of the flat buffer)
        broadcast_flat(model.embedding_weight, group)
        broadcast_flat(model.lm_head_weight, group)
    model.sync_shadow()
    model.dp_group = group if group is not None else dist.group.WORLD
    model.dp_world = dist.get_world_size(group)
    return model
