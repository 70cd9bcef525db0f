"""Data-parallel plumbing (config 5) on CPU: world_size-2 gloo processes run FlatBucket under a small torch model and check
that (1) parameters are views into the flat buffer and gradients are gathered into the flat gradient buffer, (2) the segment hooks issue every all-reduce during the
backward, (3) the reduced bucket equals the sum of the two ranks' gradients, in the flat layout the optimizer kernel
reads.  The optimizer kernel itself (ss2d_optim_clip_adam) is CUDA-only and covered by tests/test_dp_gpu.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _model():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(7, 13), torch.nn.GELU(), torch.nn.Linear(13, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from focalnet_b200.dp import FlatBucket, expected_allreduce_bytes
    m, ref = _model(), _model()
    b = FlatBucket(m.parameters(), segments=3)
    nseg = len(b.seg_bounds)
    ok = b.world == world and 2 <= nseg <= 3 and b.seg_bounds[0][0] == 0 and b.seg_bounds[-1][1] == b.numel
    ok &= all(x[1] == y[0] for x, y in zip(b.seg_bounds, b.seg_bounds[1:])) and b.numel % 4 == 0
    ok &= all(p.data_ptr() >= b.flat_param.data_ptr() and p.grad is None for p in m.parameters())
    ok &= expected_allreduce_bytes(b) == 4 * sum(p.numel() for p in ref.parameters())
    for step in range(2):  # two steps: the hooks re-arm
        g = torch.Generator().manual_seed(100 * step + rank)
        x = torch.randn(4, 7, generator=g)
        b.begin_step()
        m(x).square().mean().backward()
        issued = len(b._works)
        b.finish_reduce()
        ok &= issued == nseg  # every segment's all-reduce was launched from inside the backward
        # the unbucketed computation: both ranks' gradients summed
        tot = [torch.zeros_like(p) for p in ref.parameters()]
        for r in range(world):
            gr = torch.Generator().manual_seed(100 * step + r)
            ref.zero_grad()
            ref(torch.randn(4, 7, generator=gr)).square().mean().backward()
            for t, p in zip(tot, ref.parameters()):
                t += p.grad
        ok &= all(torch.allclose(b.grad_view(p), t, rtol=1e-6, atol=1e-7) for p, t in zip(m.parameters(), tot))  # flat layout
        b.zero_grad()
        ok &= float(b.flat_grad.abs().max()) == 0.0
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_flat_bucket_two_rank_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    ok = q.get(timeout=180)
    [p.join(timeout=60) for p in procs]
    assert ok


def test_flat_bucket_single_process_views_follow_updates():
    from focalnet_b200.dp import FlatBucket, ring_allreduce_wire_bytes
    m = _model()
    before = [p.detach().clone() for p in m.parameters()]
    b = FlatBucket(m.parameters(), segments=2)
    assert all(torch.equal(p, q) for p, q in zip(m.parameters(), before))   # values survive the re-homing
    b.flat_param.add_(1.0)                                                   # an optimizer writing the flat buffer ...
    assert all(torch.equal(p, q + 1.0) for p, q in zip(m.parameters(), before))  # ... is seen through the parameters
    b.begin_step()
    m(torch.randn(2, 7)).sum().backward()
    assert float(b.flat_grad.abs().sum()) > 0                               # autograd accumulated into the flat buffer
    b.finish_reduce()                                                        # world 1: nothing to wait for
    assert all(torch.equal(b.grad_view(p), p.grad) for p in m.parameters())  # gathered by the segment's multi-tensor copy
    b.begin_step()                                                           # next step: autograd's tensors are dropped
    assert all(p.grad is None for p in m.parameters())
    assert ring_allreduce_wire_bytes(1000, 1) == 0 and ring_allreduce_wire_bytes(1000, 8) == 1750.0
