"""Serving the head's inference methods as CUDA graphs.

``ObjectDetection.forward`` (ref src/sihl/heads/object_detection.py:99-122) is ~30 kernel launches — laterals, the
location tower, top-K, the class / box towers on K rows, the decode — and at small batches the host spends longer
enqueueing them than the GPU spends running them.  :class:`GraphedInference` captures one call for fixed input shapes
into a ``torch.cuda.CUDAGraph`` and replays it: one host launch per call, same kernels, same results.  It works with
either tower backend (``mlp_backend = "torch"`` or ``"tcgen05"``); parameters are read at replay time from the buffers
the graph captured, so call :meth:`recapture` after anything re-packs weights (an optimizer step, ``load_state_dict``).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor


class GraphedInference:
    """``graphed = GraphedInference(model.forward, example_inputs)``; ``outputs = graphed(inputs)``.

    ``fn`` takes the pyramid (a list of tensors) and returns a tuple of tensors.  Inputs of later calls must have the
    shapes and dtypes of ``example_inputs``; they are copied into the graph's static buffers (entries with one element or
    fewer — unused levels — are skipped).  The returned tensors are the graph's static outputs: valid until the next call.
    """

    def __init__(self, fn: Callable[[List[Tensor]], Tuple[Tensor, ...]], example_inputs: Sequence[Tensor],
                 warmup: int = 2) -> None:
        if not example_inputs or not example_inputs[0].is_cuda:
            raise ValueError("GraphedInference needs CUDA inputs")
        self.fn = fn
        self.device = example_inputs[0].device
        self.static_inputs = [t.detach().clone() for t in example_inputs]
        self.warmup = max(1, int(warmup))
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_outputs: Optional[Tuple[Tensor, ...]] = None
        self.recapture()

    def recapture(self) -> None:
        """(Re)capture ``fn`` on the static inputs: warm-up calls on a side stream (lazy initialisation, weight packing
        and table caches happen there, outside the capture), then the capture itself."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):
                self.fn(self.static_inputs)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            out = self.fn(self.static_inputs)
        self.graph = graph
        self.static_outputs = tuple(out) if isinstance(out, (tuple, list)) else (out,)

    @torch.no_grad()
    def __call__(self, inputs: Sequence[Tensor]) -> Tuple[Tensor, ...]:
        if len(inputs) != len(self.static_inputs):
            raise ValueError(f"expected {len(self.static_inputs)} pyramid levels, got {len(inputs)}")
        for dst, src in zip(self.static_inputs, inputs):
            if dst.numel() <= 1:
                continue
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise ValueError(f"input {tuple(src.shape)} {src.dtype} does not match the captured {tuple(dst.shape)} {dst.dtype}")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_outputs
