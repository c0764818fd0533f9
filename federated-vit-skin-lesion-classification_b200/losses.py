"""Loss functions of the path — reference ``losses.py`` surface on the fused sm_100a loss kernel.

``AsymmetricFocalLoss(gamma_neg, gamma_pos, clip, eps).forward(logits (B,C), targets (B,)) ->
scalar`` and ``build_loss(config)`` keep the reference's signatures and config keys
(losses.py:28-41,74-82). One kernel launch produces the loss AND d loss/d logits (the reference's
forward builds ~12 elementwise autograd nodes, losses.py:47-67); ``backward`` only rescales.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, kind, gamma_neg, gamma_pos, clip, eps):
        if kind == "asl":
            loss, dlogits = ops.asl_loss(logits.detach(), targets, gamma_neg, gamma_pos, clip, eps)
        else:
            loss, dlogits = ops.ce_loss(logits.detach(), targets)
        ctx.save_for_backward(dlogits)
        ctx.in_dtype = logits.dtype
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (dlogits,) = ctx.saved_tensors
        return (dlogits * grad_out).to(ctx.in_dtype), None, None, None, None, None, None


class AsymmetricFocalLoss(nn.Module):
    """Asymmetric focal loss for single-label multi-class targets (no class weights).

    gamma_neg focuses the wrong-class terms, gamma_pos the true-class term, ``clip`` shifts the
    negative probabilities down before the log, ``eps`` guards the logs — reference losses.py:17-67.
    Computed in fp32 whatever the logits' dtype (what autocast does to softmax/log/pow)."""

    def __init__(self, gamma_neg: float = 4.0, gamma_pos: float = 1.0, clip: float = 0.05,
                 eps: float = 1e-8) -> None:
        super().__init__()
        self.gamma_neg = gamma_neg
        self.gamma_pos = gamma_pos
        self.clip = clip
        self.eps = eps

    def forward(self, logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        if logits.dim() != 2 or targets.dim() != 1 or targets.shape[0] != logits.shape[0]:
            raise ValueError("expected logits (B, C) and targets (B,)")
        return _FusedLoss.apply(logits, targets, "asl", float(self.gamma_neg), float(self.gamma_pos),
                                float(self.clip), float(self.eps))


def cross_entropy(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """Mean softmax cross-entropy — what the reference's eval loop calls (utils.py:262)."""
    return _FusedLoss.apply(logits, targets, "ce", 0.0, 0.0, 0.0, 0.0)


def build_loss(config: dict) -> nn.Module:
    """Always the asymmetric focal loss, from ``loss.asymmetric.{gamma_neg,gamma_pos,clip}``."""
    a = config.get("loss", {}).get("asymmetric", {})
    return AsymmetricFocalLoss(gamma_neg=float(a.get("gamma_neg", 4)), gamma_pos=float(a.get("gamma_pos", 1)),
                               clip=float(a.get("clip", 0.05)))
