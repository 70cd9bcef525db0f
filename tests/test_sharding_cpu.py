"""N > 1 path on CPU: world_size-2 gloo processes exercise the batch sharding and the max-over-ranks reduction that
bench.py uses (the data path itself needs no collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from focalnet_b200.sharding import job_throughput, max_over_ranks, shard_range


def test_shard_range_covers_everything():
    for n in (1, 7, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ss2d_oracle as orc
    g = torch.Generator().manual_seed(0)  # every rank builds the same global batch, then scans only its slice
    Bn, Dm, N, L, G = 5, 8, 4, 40, 2
    u, dl = torch.randn(Bn, Dm, L, generator=g), 0.5 * torch.rand(Bn, Dm, L, generator=g)
    A = -0.5 * torch.rand(Dm, N, generator=g)
    Bm, Cm = torch.randn(Bn, G, N, L, generator=g), torch.randn(Bn, G, N, L, generator=g)
    lo, hi = shard_range(Bn, rank, world)
    mine = orc.scan_fwd(u[lo:hi], dl[lo:hi], A, Bm[lo:hi], Cm[lo:hi], None, None, None, True)["out"]
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, mine))
    t = max_over_ranks(1.0 + rank)                       # slowest rank defines the job time
    thr = job_throughput(hi - lo, 1.0 + rank)
    if rank == 0:
        full = orc.scan_fwd(u, dl, A, Bm, Cm, None, None, None, True)["out"]
        glued = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])], axis=0)
        out.put((bool(np.array_equal(glued, full)), t, thr))
    dist.destroy_process_group()


def test_two_rank_batch_sharding_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    same, t, thr = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert same                      # sharded scan == unsharded scan, bit for bit (independent sequences)
    assert t == 2.0                  # max over ranks
    assert abs(thr - 5 / 2.0) < 1e-12  # all units / slowest rank
