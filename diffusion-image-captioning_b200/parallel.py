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


def enable_data_parallel(model, group=None):
    """Make every rank start from rank 0's weights (trainable flat buffer + frozen embedding / lm_head) and switch the
    model's train_func to DP mode (gradient all-reduce + shared t)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return model
    broadcast_flat(model.flat, group)
    if not model.hp["TRAIN_EMBEDDING"]:  # (there, embedding / lm_head are trainable slots of the flat buffer)
        broadcast_flat(model.embedding_weight, group)
        broadcast_flat(model.lm_head_weight, group)
    model.sync_shadow()
    model.dp_group = group if group is not None else dist.group.WORLD
    model.dp_world = dist.get_world_size(group)
    return model
