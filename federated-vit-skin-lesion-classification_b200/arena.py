"""Flat parameter arena: every parameter of a model as a view into ONE contiguous fp32 buffer,
with a matching gradient buffer and a bf16 shadow of the weights.

Why: the three HBM-bound sweeps of the hot path — clip+AdamW(+EMA) (reference train.py:156-162),
the FedAvg fold and its NCCL allreduce — each become a single pass over one
buffer instead of ~150 per-tensor launches, and the tensor-core GEMMs read their bf16 weights from
the shadow that the optimiser sweep refreshes for free.

Layout in HBM: parameters in ``named_parameters()`` order (cls_token, pos_embed, patch_embed,
blocks.0 … blocks.L-1, norm, metadata_branch, classifier — so every LLRD group of reference
model.py:228-270 is one contiguous range), each start aligned to 64 elements (256 B); padding is
zero and stays zero. ViT-B/16 + head: 86.2 M elements = 345 MB fp32 (+345 MB grads, +172 MB bf16).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops

ALIGN = 64  # elements


class FlatArena:
    def __init__(self, module: nn.Module, with_lp: bool = True) -> None:
        named = [(n, p) for n, p in module.named_parameters()]
        if not named:
            raise ValueError("module has no parameters")
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("FlatArena lives in HBM: move the model to a CUDA device first")
        self.device = dev
        self.names: List[str] = []
        self.offsets: Dict[str, Tuple[int, int]] = {}
        self._params: List[nn.Parameter] = []
        off = 0
        for n, p in named:
            if p.dtype != torch.float32:
                raise TypeError(f"parameter {n} is {p.dtype}; the arena holds fp32 master weights")
            self.names.append(n)
            self.offsets[n] = (off, p.numel())
            self._params.append(p)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self._index = {n: i for i, n in enumerate(self.names)}
        self.params = torch.zeros(off, device=dev, dtype=torch.float32)
        self.grads = torch.zeros(off, device=dev, dtype=torch.float32)
        self.lp = torch.zeros(off, device=dev, dtype=torch.bfloat16) if with_lp else None
        self._lp_versions: Optional[List[int]] = None
        # True while the gradient buffer is known to hold zeros (fresh, zeroed, or just swept by an
        # optimiser step that wrote zeros back): zero_grads() is then free. Every backward clears it.
        self.grads_clean = True
        # per-parameter views of the gradient buffer and the bf16 shadow, made once: every step asks for
        # each of them several times (zero_grad, clip, the optimiser's foreign-gradient check, the GEMMs'
        # weight operands), and slicing + reshaping ~150 tensors costs the host about a millisecond a time
        self._grad_views: Dict[str, torch.Tensor] = {}
        self._lp_views: Dict[str, torch.Tensor] = {}
        with torch.no_grad():
            for (n, p) in named:
                o, k = self.offsets[n]
                self.params[o:o + k].copy_(p.data.reshape(-1))
                p.data = self.params[o:o + k].view(p.shape)
                p._fv_arena = (self, n)
        module._fv_arena = self
        self.attach_grads()

    # ------------------------------------------------------------------------------------------
    def view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        o, k = self.offsets[name]
        p = self._params[self._index[name]]
        return buf[o:o + k].view(p.shape)

    def _cached(self, cache: Dict[str, torch.Tensor], buf: torch.Tensor, name: str) -> torch.Tensor:
        v = cache.get(name)
        if v is None:
            v = cache[name] = self.view(buf, name)
        return v

    def grad_view(self, p: nn.Parameter) -> torch.Tensor:
        return self._cached(self._grad_views, self.grads, p._fv_arena[1])

    def attach_grads(self) -> None:
        """Point every ``p.grad`` at its slice of the gradient buffer (requires_grad params only)."""
        for n, p in zip(self.names, self._params):
            if p.requires_grad:
                p.grad = self._cached(self._grad_views, self.grads, n)

    def zero_grads(self) -> None:
        if not self.grads_clean:
            self.grads.zero_()
            self.grads_clean = True
        self.attach_grads()

    def owns(self, p: nn.Parameter) -> bool:
        a = getattr(p, "_fv_arena", None)
        if a is None or a[0] is not self:
            return False
        o, _ = self.offsets[a[1]]
        return p.data_ptr() == self.params.data_ptr() + 4 * o

    def intact(self) -> bool:
        """False once something (``module.to()``, a load that re-allocates) re-pointed a parameter."""
        return all(self.owns(p) for p in self._params)

    # ------------------------------------------------------------------------------------------
    # bf16 shadow of the weights
    # ------------------------------------------------------------------------------------------
    def _versions(self) -> List[int]:
        return [p._version for p in self._params]

    def mark_lp_fresh(self) -> None:
        self._lp_versions = self._versions()

    def refresh_lp(self, force: bool = False) -> None:
        """Re-cast the shadow if any parameter changed through PyTorch since the last refresh.
        (The fused optimiser sweep writes the shadow itself and then calls ``mark_lp_fresh``.)"""
        if self.lp is None:
            raise RuntimeError("arena was built without a bf16 shadow")
        if force or self._lp_versions != self._versions():
            ops.cast_bf16(self.params, self.lp)
            self.mark_lp_fresh()

    def lp_view(self, p: nn.Parameter) -> torch.Tensor:
        return self._cached(self._lp_views, self.lp, p._fv_arena[1])

    # ------------------------------------------------------------------------------------------
    def segments(self, param_groups: Iterable[dict], by_group: bool = False) -> Tuple[List[int], List[float], List[float]]:
        """(end offsets, lr, weight_decay) per arena range; lr = -1 marks parameters that belong to
        no optimiser group (reference quirk: cls_token / pos_embed, model.py:228-270). Adjacent
        ranges with identical hyper-parameters are merged — or, with ``by_group``, adjacent ranges of the
        same param group whatever their values: the table then keeps its length when a scheduler moves
        the learning rates (a captured CUDA graph holds the table's device addresses)."""
        hp: Dict[int, Tuple[float, float]] = {}
        gid: Dict[int, int] = {}
        for gi, g in enumerate(param_groups):
            for p in g["params"]:
                hp[id(p)] = (float(g["lr"]), float(g["weight_decay"]))
                gid[id(p)] = gi
        ends: List[int] = []
        lrs: List[float] = []
        wds: List[float] = []
        last_key = None
        for i, (n, p) in enumerate(zip(self.names, self._params)):
            nxt = self.offsets[self.names[i + 1]][0] if i + 1 < len(self.names) else self.numel
            lr, wd = hp.get(id(p), (-1.0, 0.0))
            g = gid.get(id(p), -1)
            if not p.requires_grad:
                lr, wd, g = -1.0, 0.0, -1
            key = g if by_group else (lr, wd)
            if ends and key == last_key:
                ends[-1] = nxt
            else:
                ends.append(nxt)
                lrs.append(lr)
                wds.append(wd)
            last_key = key
        return ends, lrs, wds


def arena_of(module: nn.Module) -> Optional[FlatArena]:
    a = getattr(module, "_fv_arena", None)
    if a is not None and not a.intact():
        return None
    return a


def ensure_arena(module: nn.Module) -> FlatArena:
    a = arena_of(module)
    if a is None:
        a = FlatArena(module)
    return a
