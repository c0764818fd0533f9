"""CUDA-graph capture of the eval forward and of a whole training step.

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


class GraphedTrainStep:
    """One optimisation step — zero_grad, forward (autocast), loss, backward, global-norm clip, fused AdamW
    (+ EMA) — captured once and replayed: ``loss = step(images, labels)`` for a fixed batch shape.

    What train.py:139-162 of the reference does per iteration is ~250 dependent kernel launches here; at
    small batch sizes (BASELINE config 1: ViT-Tiny, batch 16) the host cannot enqueue them as fast as the
    GPU retires them. Everything on the path is capturable: the kernels take their TMA descriptors by
    value, activations come from the graph's private pool, the optimiser reads its step-dependent scalars
    from device memory (``FusedAdamW.enable_graph_mode``), stochastic depth / dropout draw from the
    graph-registered Philox state. The returned loss is a device scalar that the next replay overwrites.

    Building it runs ``warmup`` real steps on the example batch (lazy initialisation must happen outside
    the capture). With ``preserve_state=True`` (default) parameters, moments, EMA shadow, step count,
    module buffers and the CUDA RNG state are snapshotted before and restored after, so construction
    leaves the training trajectory untouched and the first call performs the first step.
    """

    def __init__(self, model: nn.Module, criterion, optimizer, images: torch.Tensor, labels: torch.Tensor,
                 metadata: Optional[torch.Tensor] = None, grad_clip: Optional[float] = 1.0,
                 amp_dtype: Optional[torch.dtype] = torch.bfloat16, warmup: int = 2,
                 preserve_state: bool = True) -> None:
        from .utils import clip_grad_norm

        if not model.training:
            raise RuntimeError("GraphedTrainStep captures a training step: call model.train() first")
        if not images.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors")
        if not hasattr(optimizer, "enable_graph_mode"):
            raise RuntimeError("GraphedTrainStep needs fedvit_b200.optim.FusedAdamW")
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.grad_clip, self.amp_dtype = grad_clip, amp_dtype
        self._clip = clip_grad_norm
        self.static_x, self.static_y = images.clone(), labels.clone()
        self.static_meta = metadata.clone() if metadata is not None else None
        optimizer.enable_graph_mode()
        dev = images.device
        snap = self._snapshot() if preserve_state else None
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # real steps: lazy init, allocator pools, kernel attributes
            for _ in range(max(1, warmup)):
                self._run()  # eager step() in graph mode ticks by itself
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        # torch.cuda.graph only records: the step it describes first RUNS at the first replay, so the step
        # count is not advanced here (host-side bookkeeping inside _run happens once, at capture)
        with torch.cuda.graph(self.graph):
            self.static_loss = self._run()
        if optimizer.ema is not None:  # step() counted a fused EMA update that the capture did not execute
            optimizer.ema._fused_updates -= 1
        if snap is not None:
            self._restore(snap)

    def _snapshot(self):
        opt, arena = self.optimizer, self.optimizer.arena
        ema = opt.ema
        return {
            "params": arena.params.clone(), "m": opt.exp_avg.clone(), "v": opt.exp_avg_sq.clone(),
            "ema": ema.flat.clone() if ema is not None else None,
            "ema_counts": (ema._fused_updates, ema._seen_fused) if ema is not None else None,
            "step": opt.step_count, "buffers": [b.clone() for b in self.model.buffers()],
            "rng": torch.cuda.get_rng_state(arena.device),
        }

    def _restore(self, snap) -> None:
        opt, arena = self.optimizer, self.optimizer.arena
        with torch.no_grad():
            arena.params.copy_(snap["params"])
            opt.exp_avg.copy_(snap["m"])
            opt.exp_avg_sq.copy_(snap["v"])
            if snap["ema"] is not None:
                opt.ema.flat.copy_(snap["ema"])
                opt.ema._fused_updates, opt.ema._seen_fused = snap["ema_counts"]
            for b, saved in zip(self.model.buffers(), snap["buffers"]):
                b.copy_(saved)
        opt.step_count = snap["step"]
        opt._set_device_step(snap["step"])
        opt._tick_pending = False
        torch.cuda.set_rng_state(snap["rng"], arena.device)
        if arena.lp is not None:
            arena.refresh_lp(force=True)  # the captured forward reads the bf16 shadow as it finds it

    def _run(self) -> torch.Tensor:
        self.optimizer.zero_grad(set_to_none=True)
        with torch.amp.autocast("cuda", enabled=self.amp_dtype is not None, dtype=self.amp_dtype or torch.bfloat16):
            loss = self.criterion(self.model(self.static_x, metadata=self.static_meta)["logits"], self.static_y)
        loss.backward()
        if self.grad_clip is not None:
            self._clip(self.model.parameters(), self.grad_clip, optimizer=self.optimizer)
        self.optimizer.step()
        return loss.detach()

    def __call__(self, images: torch.Tensor, labels: torch.Tensor,
                 metadata: Optional[torch.Tensor] = None) -> torch.Tensor:
        if images.shape != self.static_x.shape or labels.shape != self.static_y.shape:
            raise ValueError(f"captured for batch shape {tuple(self.static_x.shape)}, got {tuple(images.shape)}")
        self.static_x.copy_(images, non_blocking=True)
        self.static_y.copy_(labels, non_blocking=True)
        if self.static_meta is not None and metadata is not None:
            self.static_meta.copy_(metadata, non_blocking=True)
        arena = self.optimizer.arena
        if arena.lp is not None:
            # parameters written through PyTorch since the last step (a FedAvg install, a checkpoint load):
            # the captured forward reads the bf16 shadow as it finds it, so bring it up to date here
            arena.refresh_lp()
        if not arena.grads_clean:  # the captured zero_grad was a no-op (the sweep before it leaves zeros)
            arena.zero_grads()
        self.optimizer.graph_tick()
        self.graph.replay()
        arena.grads_clean = bool(getattr(self.optimizer, "fuse_zero_grad", False))
        self.optimizer._tick_pending = False  # consumed by the replayed sweep
        if self.optimizer.ema is not None:  # the replayed sweep updated the shadow (utils.EMA.update() then skips)
            self.optimizer.ema._fused_updates += 1
        return self.static_loss
