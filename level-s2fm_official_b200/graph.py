"""CUDA-graph capture of a whole render iteration (sampler -> fused field forward -> compositing -> loss -> backward).

One iteration of the hot path is ~25 of our launches plus ~40 tiny autograd / fill kernels; the device work of the small ones is a
few microseconds each, so the iteration is partly LAUNCH-bound.  Every entry point of libls2fm is asynchronous on the caller's
stream, allocates nothing and never synchronises the host (include/ls2fm.h), which is exactly what stream capture needs: the whole
iteration -- autograd backward included -- is recorded once and replayed as ONE graph launch.

The reference has nothing of the kind (its iteration synchronises the host several times, e.g. models/Renderer.py:213-313,
models/SDF.py:159-200); paths that still need a host read-back (``SDF.sphere_tracing`` reads the iteration count) cannot be captured.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    """``step = GraphedStep(fn, example_inputs)``; ``outs = step(*inputs)``.

    ``fn(*tensors) -> tensor | tuple of tensors`` must be a pure function of its tensor arguments and of persistent state
    (parameters, a ``parallel.GradBucket`` that it zeroes and accumulates into); it runs ``warmup`` times eagerly on a side
    stream, is captured once, and every later call copies the inputs into the captured input buffers and replays the graph.
    The returned tensors are the graph's static output buffers (overwritten by the next call)."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        self.static_in = [x.detach().clone() for x in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
