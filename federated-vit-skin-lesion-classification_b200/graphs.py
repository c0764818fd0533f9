"""CUDA-graph capture of the forward pass for launch-bound batch sizes.

A ViT-B/16 forward is ~100 kernel launches; below batch ~64 the GPU finishes each kernel faster than
Python can enqueue the next one (profiles/r1_eval_sweep_vitb.jsonl: 4 ms per forward regardless of
batch). Every libfedvit entry point is capturable (asynchronous, no allocation, no host sync; TMA
descriptors are built on the host and passed by value), so the whole eval forward —
``validate``'s per-batch work, reference train.py:199-205 — replays as one graph launch.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


class GraphedForward:
    """``logits = GraphedForward(model, example)(images)`` for a fixed batch shape, eval mode."""

    def __init__(self, model: nn.Module, example: torch.Tensor, metadata: Optional[torch.Tensor] = None,
                 amp_dtype: Optional[torch.dtype] = torch.bfloat16, warmup: int = 3) -> None:
        if model.training:
            raise RuntimeError("GraphedForward captures the eval forward: call model.eval() first")
        if not example.is_cuda:
            raise RuntimeError("GraphedForward needs CUDA tensors")
        self.model = model
        self.amp_dtype = amp_dtype
        self.static_in = example.clone()
        self.static_meta = metadata.clone() if metadata is not None else None
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side):  # warm-up off the capture: lazy init, allocator pools, attributes
            for _ in range(max(1, warmup)):
                self._run()
        torch.cuda.current_stream(example.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._run()

    def _run(self) -> torch.Tensor:
        with torch.no_grad(), torch.amp.autocast("cuda", enabled=self.amp_dtype is not None,
                                                 dtype=self.amp_dtype or torch.bfloat16):
            return self.model(self.static_in, metadata=self.static_meta)["logits"]

    def __call__(self, images: torch.Tensor, metadata: Optional[torch.Tensor] = None) -> torch.Tensor:
        if images.shape != self.static_in.shape:
            raise ValueError(f"captured for batch shape {tuple(self.static_in.shape)}, got {tuple(images.shape)}")
        self.static_in.copy_(images, non_blocking=True)
        if self.static_meta is not None and metadata is not None:
            self.static_meta.copy_(metadata, non_blocking=True)
        self.graph.replay()
        return self.static_out
