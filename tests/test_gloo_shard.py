"""world_size-2 gloo test of the multi-GPU host logic: contiguous batch shards, keys replicated, no
data-path collective; only the timing reduction (max over ranks) and the final parity check communicate.
The per-rank compute is stood in for by the oracle here (no GPU on this box)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from tfhe_jl_b200.sharding import shard_range
    P = O.small_params(O.PARAMS_80, 16)
    keys = O.keygen(P, 99)            # every rank derives the same (replicated) keys
    ctx = O.Context(keys)
    rng = O.Rng(5)
    bits = np.random.default_rng(0).integers(0, 2, (5, 2)).astype(bool)
    x, y = O.encrypt(rng, keys, bits[:, 0]), O.encrypt(rng, keys, bits[:, 1])
    lo, hi = shard_range(5, rank, world)
    out = ctx.gate(O.NAND, x[lo:hi], y[lo:hi], nthreads=1)
    t = torch.tensor([1.0 + rank])    # stand-in for this rank's device time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, out))
    if rank == 0:
        full = ctx.gate(O.NAND, x, y, nthreads=1)
        ok = all(np.array_equal(full[a:b], o) for a, b, o in gathered)
        cover = sorted((a, b) for a, b, _ in gathered)
        q.put((ok, cover, float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    ok, cover, tmax = q.get(timeout=180)
    [p.join(60) for p in procs]
    assert ok and cover == [(0, 3), (3, 5)] and tmax == 2.0


def test_shard_range_partitions():
    from tfhe_jl_b200.sharding import shard_range
    for count in (0, 1, 7, 1 << 20):
        for world in (1, 2, 4, 8):
            r = [shard_range(count, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == count
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
