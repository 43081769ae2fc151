"""CPU, world_size 2 over gloo: the data-parallel host logic (weight broadcast, flat-gradient all-reduce, shared t, caption
sharding) that the NCCL path uses on the GPU box."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import O, ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import clipdlm
    from clipdlm import parallel
    r, lr, w = parallel.init_process_group_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(rank)
    flat = torch.randn(1000)
    parallel.broadcast_flat(flat)                       # every rank now holds rank 0's weights
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    # DP equivalence on a toy quadratic: mean of shard gradients == gradient of the global batch
    torch.manual_seed(123)
    data = torch.randn(8, 1000)
    lo, hi = parallel.shard_range(8, rank, world)
    g_local = (flat[None] - data[lo:hi]).mean(0)        # d/dw of 0.5*|w - x|^2 averaged over the shard
    g = parallel.allreduce_mean_(g_local.clone())
    g_global = (flat[None] - data).mean(0)
    t = torch.randint(0, 1000, (100, 1, 1))
    dist.broadcast(t, src=0)
    ts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(ts, t)
    out[rank] = (same, float((g - g_global).abs().max()), all(torch.equal(x, ts[0]) for x in ts))
    dist.destroy_process_group()


def test_dp_host_logic_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for r in range(2):
        same, err, t_same = out[r]
        assert same and err < 1e-6 and t_same
