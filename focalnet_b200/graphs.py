"""CUDA-graph replay of a fixed-shape inference forward (full-resolution evaluation runs one image per call, ITS/eval.py:19:
`test_dataloader(..., batch_size=1)`; at that size the step is bound by ~1300 kernel launches issued from Python, not by the
GPU).  Everything this library launches goes to torch's current stream through the C ABI and allocates through torch's
caching allocator, and there is no host synchronisation on the path, so the whole forward of the patched model captures
into one graph; a replay costs one launch."""
from __future__ import annotations

import torch


class GraphedForward:
    """``g = GraphedForward(fn, example)``; ``y = g(x)`` copies x into the captured input, replays, returns the captured
    output tensor(s) (valid until the next call).  `fn` must be shape-static and free of host syncs (e.g.
    ``lambda x: eval_forward(model, x)`` under ``torch.no_grad()`` with ``patch_ss2d(model)`` applied)."""

    def __init__(self, fn, example: torch.Tensor, warmup: int = 3):
        if not example.is_cuda:
            raise RuntimeError("GraphedForward needs CUDA tensors (focalnet_b200 has no CPU path)")
        self._x = example.clone()
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # first launches: shared-memory opt-ins, cuDNN / cuBLAS plan selection, allocator growth
                fn(self._x)
        torch.cuda.current_stream(example.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self._graph):
            self._y = fn(self._x)

    def __call__(self, x: torch.Tensor):
        self._x.copy_(x, non_blocking=True)
        self._graph.replay()
        return self._y
