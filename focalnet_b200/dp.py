"""Data-parallel training plumbing for the ITS model (configs 3 / 5 of BASELINE.json).

One process per GPU (torchrun), the batch axis is the only partition: every rank runs the unchanged model on its own 32
crops; the ONE exchange step of the path is the gradient all-reduce (2,541,673 fp32 = 10.17 MB), followed by the global-
norm clip at 0.001 and Adam of ITS/train.py:89-91.

* ``FlatBucket`` re-homes every parameter as a view into a flat fp32 buffer and owns a flat gradient buffer of the same
  layout (parameters in reverse registration order ~ the order the backward produces their gradients), split into a few
  contiguous segments.  A post-accumulate-grad hook per parameter counts a segment down; when the last gradient of a
  segment has landed, the segment's gradients are gathered into the flat buffer by ONE multi-tensor copy and its
  all-reduce (SUM) is issued asynchronously — NCCL runs it on its own stream while the backward keeps going, so only the
  last segment's ~tens of microseconds can be exposed.  (Pre-setting ``p.grad`` to views of the flat buffer would make
  autograd accumulate with one tiny ``add_`` kernel per parameter: 270 launches, 4 ms of a 195 ms step.)
* ``FusedClipAdam`` is the optimizer side: the C ABI's ``ss2d_optim_clip_adam`` (two launches over the flat bucket:
  deterministic sum of squares, then clip + Adam + zero-grad; 1 / world size folded in).  CUDA only — no fallback.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import _lib


class FlatBucket:
    def __init__(self, params: Iterable[torch.nn.Parameter], segments: int = 3, group=None):
        plist = [p for p in params if p.requires_grad]
        if not plist:
            raise ValueError("FlatBucket: no trainable parameters")
        dev = plist[0].device
        if any(p.dtype != torch.float32 or p.device != dev for p in plist):
            raise ValueError("FlatBucket: fp32 parameters on one device only")
        self.params: List[torch.nn.Parameter] = list(reversed(plist))  # ~ gradient arrival order in the backward
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        # every parameter starts on a 16-byte boundary (4 floats) so kernels can use 128-bit accesses on any slice
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self._offs = offs
        self.payload = sum(p.numel() for p in self.params)
        self.flat_param = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self._gviews = []
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                self.flat_param[o:o + p.numel()].view_as(p).copy_(p)
                p.data = self.flat_param[o:o + p.numel()].view_as(p)
                p.grad = None                                         # autograd hands over its own tensor: no add_ kernel
                self._gviews.append(self.flat_grad[o:o + p.numel()].view_as(p))
        # contiguous segments of roughly equal size, cut at parameter boundaries
        segments = max(1, min(segments, len(self.params)))
        target, self.seg_bounds, self._seg_of, cur = n / segments, [], [], 0
        start = 0
        for i, (p, o) in enumerate(zip(self.params, offs)):
            self._seg_of.append(cur)
            end = offs[i + 1] if i + 1 < len(offs) else n
            if (end - start >= target and cur < segments - 1) or i + 1 == len(offs):
                self.seg_bounds.append((start, end))
                start, cur = end, cur + 1
        self._seg_total = [self._seg_of.count(s) for s in range(len(self.seg_bounds))]
        self._pending = list(self._seg_total)
        self._works: list = []
        self._seg_members = [[i for i, sg in enumerate(self._seg_of) if sg == s] for s in range(len(self.seg_bounds))]
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(self._seg_of[i])) for i, p in enumerate(self.params)]

    # ---- backward-side: launch a segment's all-reduce as soon as its last gradient has been accumulated -------------
    def _flush_segment(self, seg):
        """Gather the segment's gradients into the flat buffer (one multi-tensor copy; a parameter without a gradient keeps
        the zeros the optimizer kernel left) and launch its all-reduce."""
        idx = [i for i in self._seg_members[seg] if self.params[i].grad is not None]
        if idx:
            with torch.no_grad():
                torch._foreach_copy_([self._gviews[i] for i in idx], [self.params[i].grad for i in idx])
        if self.world > 1:
            a, b = self.seg_bounds[seg]
            self._works.append(dist.all_reduce(self.flat_grad[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self._pending[seg] = 0

    def _make_hook(self, seg):
        def hook(_param):
            self._pending[seg] -= 1
            if self._pending[seg] == 0:
                self._flush_segment(seg)
        return hook

    def begin_step(self):
        """Call before the forward of every step: drops last step's gradient tensors (their values were consumed from the
        flat buffer, which the optimizer kernel zeroed) and re-arms the segment counters."""
        for p in self.params:
            p.grad = None
        self._pending = list(self._seg_total)
        self._works = []

    def finish_reduce(self):
        """After the backward: flush the segments whose last hook did not fire (a parameter that received no gradient), then
        make the current stream wait for the all-reduces (stream-level wait, the host does not block)."""
        for seg, left in enumerate(self._pending):
            if left > 0:
                self._flush_segment(seg)
        for w in self._works:
            w.wait()
        self._works = []

    def grad_view(self, p):
        """The slice of the flat gradient buffer that belongs to parameter `p` (what the optimizer kernel reads)."""
        return self._gviews[next(i for i, q in enumerate(self.params) if q is p)]

    def allreduce_whole(self):
        """One blocking-on-stream all-reduce of the whole bucket (what `allreduce_ms` times)."""
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)

    def zero_grad(self):
        self.flat_grad.zero_()


class FusedClipAdam:
    """clip_grad_norm_(params, max_norm) + Adam(lr, betas, eps).step() + zero_grad() of ITS/train.py:16,61,89-91 as one
    pass over a FlatBucket on this library's kernel.  `last_norm` is a device scalar (no host sync per step)."""

    def __init__(self, bucket: FlatBucket, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 0.001):
        if not bucket.flat_param.is_cuda:
            raise RuntimeError("FusedClipAdam runs on the CUDA library only (focalnet_b200 has no CPU path)")
        self.b, self.lr, self.betas, self.eps, self.max_norm = bucket, lr, betas, eps, max_norm
        self.exp_avg = torch.zeros_like(bucket.flat_param)
        self.exp_avg_sq = torch.zeros_like(bucket.flat_param)
        self.partials = torch.empty(int(_lib.lib().ss2d_optim_partials()), device=bucket.flat_param.device, dtype=torch.float32)
        self.last_norm = torch.zeros(1, device=bucket.flat_param.device, dtype=torch.float32)
        self.steps = 0

    # ---- checkpoint interop with the reference's training script (ITS/train.py:24-27,110-113 saves / restores
    # torch.optim.Adam.state_dict()): same dictionary layout, parameters indexed in model.parameters() order -------------
    def state_dict(self):
        b = self.b
        order = list(reversed(range(len(b.params))))  # b.params is reversed registration order
        state = {}
        if self.steps > 0:
            for idx, i in enumerate(order):
                o, n, p = b._offs[i], b.params[i].numel(), b.params[i]
                state[idx] = {"step": torch.tensor(float(self.steps)), "exp_avg": self.exp_avg[o:o + n].view_as(p).clone(),
                              "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "params": list(range(len(order)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        b = self.b
        order = list(reversed(range(len(b.params))))
        g = sd["param_groups"][0]
        if len(g["params"]) != len(order):
            raise ValueError("optimizer state was saved for a different parameter list")
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = 0
        for idx, i in enumerate(order):
            st = sd["state"].get(idx)
            if st is None:
                continue
            o, n, p = b._offs[i], b.params[i].numel(), b.params[i]
            self.exp_avg[o:o + n].view_as(p).copy_(st["exp_avg"])
            self.exp_avg_sq[o:o + n].view_as(p).copy_(st["exp_avg_sq"])
            steps = max(steps, int(float(st["step"])))
        self.steps = steps

    def step(self, lr: Optional[float] = None):
        self.steps += 1
        b = self.b
        stream = torch.cuda.current_stream(b.flat_param.device).cuda_stream
        with torch.cuda.device(b.flat_param.device):
            rc = _lib.lib().ss2d_optim_clip_adam(
                b.flat_param.data_ptr(), b.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), b.numel,
                self.partials.data_ptr(), self.last_norm.data_ptr(), float(self.lr if lr is None else lr), float(self.betas[0]),
                float(self.betas[1]), float(self.eps), self.steps, float(self.max_norm), 1.0 / b.world, stream)
        _lib.check(rc, "ss2d_optim_clip_adam")


def dp_train_step(model, bucket: FlatBucket, opt: FusedClipAdam, loss_fn, x, label):
    """One data-parallel iteration of ITS/train.py:57-91: forward, loss, backward (segment all-reduces overlap its tail),
    wait, fused clip + Adam + zero-grad.  Returns the local loss (device scalar)."""
    bucket.begin_step()
    loss = loss_fn(model(x), label)
    loss.backward()
    bucket.finish_reduce()
    opt.step()
    return loss.detach()


def expected_allreduce_bytes(bucket: FlatBucket) -> int:
    return 4 * bucket.payload


def ring_allreduce_wire_bytes(nbytes: int, world: int) -> float:
    """bytes each rank sends (= receives) for a ring / NVLS-tree all-reduce of nbytes: 2 (w-1)/w nbytes."""
    return 2.0 * (world - 1) / world * nbytes if world > 1 else 0.0


__all__ = ["FlatBucket", "FusedClipAdam", "dp_train_step", "expected_allreduce_bytes", "ring_allreduce_wire_bytes"]
