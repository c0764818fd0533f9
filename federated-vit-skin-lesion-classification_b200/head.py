"""Classifier head on the libfedvit kernels — scope row f2.

``Linear(F, 512) -> GELU -> Dropout(p) -> Linear(512, C)`` as the reference builds it
(model.py:139-144) and calls it (model.py:206). Forward and backward are one autograd node over
our own GEMMs, so a training step launches no library GEMM at all:

  autocast (bf16)  u = feats W1^T + b1 on the tcgen05 GEMM with the GELU epilogue (activation and
                   GELU' out in fp32 — 512 columns, tiny); logits = h W2^T + b2 on the FFMA kernel
                   (C = 7 or 8 columns is below the tensor-core tile's granularity). Backward:
                   dW2 / db2 / dh on the FFMA kernel (dh fused with GELU'), dW1 + db1 on the split-K
                   tcgen05 weight-gradient kernel, dfeats on the tcgen05 dgrad kernel.
  fp32             the same five products on the FFMA kernel (the 1e-4 parity mode).

Parameter gradients are accumulated straight into ``p.grad`` (the FlatArena buffer when there is
one), like the backbone's. Dropout uses torch's RNG for the mask (plumbing) and folds it into the
saved GELU'. The 13-d metadata MLP (BatchNorm1d) has its own fused kernels: ``metadata.py``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from ._lib import FedVitError
from .vit import _grad_buffer, _split_k_for, _to_bf16, weight_operand

_K, _MN = ops.MAJOR_K, ops.MAJOR_MN
_E = ops.EPI


class _HeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats: Tensor, lin1: nn.Linear, lin2: nn.Linear, p_drop: float, lp: bool, *params):
        B = feats.shape[0]
        H1, C = lin1.weight.shape[0], lin2.weight.shape[0]
        dev = feats.device
        x32 = feats.float().contiguous()
        x_op = _to_bf16(x32) if lp else x32
        h = torch.empty((B, H1), device=dev, dtype=torch.float32)
        dact = torch.empty((B, H1), device=dev, dtype=torch.float32)
        ops.gemm_gelu(x_op, weight_operand(lin1.weight, lp), lin1.bias.detach(), h, dact)
        if p_drop > 0.0:
            keep = 1.0 - p_drop
            mask = torch.empty_like(h).bernoulli_(keep).div_(keep)
            h = h * mask
            dact = dact * mask  # d(dropout(gelu(u)))/du = mask * gelu'(u)
        logits = torch.empty((B, C), device=dev, dtype=torch.float32)
        ops.gemm(h, lin2.weight.detach(), lin2.bias.detach(), logits, None, _K, _K, _E["none"], 1, 0)
        ctx.lin1, ctx.lin2, ctx.lp = lin1, lin2, lp
        ctx.need_dx = feats.requires_grad
        ctx.save_for_backward(x_op, h, dact)
        ctx.nparams = len(params)
        return logits

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dlogits: Tensor):
        x_op, h, dact = ctx.saved_tensors
        lin1, lin2, lp = ctx.lin1, ctx.lin2, ctx.lp
        B, H1 = h.shape
        F = x_op.shape[1]
        dev = h.device
        dl = dlogits.float().contiguous()
        if lin2.weight.requires_grad:
            ops.gemm(dl, h, None, _grad_buffer(lin2.weight), None, _MN, _MN, _E["accum"], 1, 0)
        if lin2.bias is not None and lin2.bias.requires_grad:
            ops.colsum(dl, _grad_buffer(lin2.bias), True)
        # du = (dlogits W2) * gelu'(u) [* dropout mask]
        du = torch.empty((B, H1), device=dev, dtype=torch.float32)
        ops.gemm(dl, lin2.weight.detach(), None, du, dact, _K, _MN, _E["dgelu"], 1, 0)
        want_b = lin1.bias is not None and lin1.bias.requires_grad
        du_op = _to_bf16(du) if lp else du
        if lp and lin1.weight.requires_grad:
            ops.wgrad(du_op, x_op, _grad_buffer(lin1.weight), _grad_buffer(lin1.bias) if want_b else None,
                      _split_k_for(H1, F, B))
        else:
            if lin1.weight.requires_grad:
                ops.gemm(du_op, x_op, None, _grad_buffer(lin1.weight), None, _MN, _MN, _E["accum"], 1, 0)
            if want_b:
                ops.colsum(du, _grad_buffer(lin1.bias), True)
        dfeats = None
        if ctx.need_dx:
            dfeats = torch.empty((B, F), device=dev, dtype=torch.float32)
            ops.gemm(du_op, weight_operand(lin1.weight, lp), None, dfeats, None, _K, _MN, _E["none"], 1, 0)
        return (dfeats, None, None, None, None) + (None,) * ctx.nparams


def classifier_head(feats: Tensor, classifier: nn.Sequential, training: bool) -> Tensor:
    """logits = classifier(feats) for the reference's ``Sequential(Linear, GELU, Dropout, Linear)``."""
    if not feats.is_cuda:
        raise FedVitError("fedvit_b200 head runs on CUDA (sm_100a) only — no CPU/MPS fallback on this path")
    lin1, act, drop, lin2 = classifier[0], classifier[1], classifier[2], classifier[3]
    if not (isinstance(lin1, nn.Linear) and isinstance(act, nn.GELU) and isinstance(drop, nn.Dropout)
            and isinstance(lin2, nn.Linear)) or getattr(act, "approximate", "none") != "none":
        raise FedVitError("classifier_head: expected Sequential(Linear, GELU(erf), Dropout, Linear)")
    lp = torch.is_autocast_enabled("cuda")
    if lp and (lin1.weight.shape[1] % 8 or lin1.weight.shape[0] % 8):
        lp = False  # tensor-core operands need 16-byte aligned rows; odd widths take the FFMA kernel
    params = [p for p in (lin1.weight, lin1.bias, lin2.weight, lin2.bias) if p is not None]
    p_drop = float(drop.p) if training else 0.0
    return _HeadFunction.apply(feats, lin1, lin2, p_drop, lp, *params)
