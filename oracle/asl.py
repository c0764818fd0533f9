"""Asymmetric focal loss — restatement of reference losses.py:41-67 (AsymmetricFocalLoss.forward).

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def asymmetric_focal_loss(logits: torch.Tensor, targets: torch.Tensor, gamma_neg: float = 4.0,
                          gamma_pos: float = 1.0, clip: float = 0.05, eps: float = 1e-8) -> torch.Tensor:
    """losses.py:41-67: softmax -> one-hot -> clamps -> focal-weighted log terms, sum over classes,
    mean over the batch. Defaults are losses.py:28-34; config.yaml:137-140 gives 4 / 1 / 0.05."""
    c = logits.size(1)
    p = torch.softmax(logits, dim=1)                       # losses.py:47
    y = F.one_hot(targets, c).float()                      # losses.py:48
    p_pos = p.clamp(min=eps)                               # losses.py:51
    p_neg = p.clamp(max=1.0 - eps)                         # losses.py:52
    if clip > 0:
        p_neg = (p_neg - clip).clamp(min=eps)              # losses.py:55-56
    pos = y * torch.log(p_pos)                             # losses.py:59
    neg = (1.0 - y) * torch.log(1.0 - p_neg)               # losses.py:60
    w_pos = (1.0 - p).clamp(min=0.0) ** gamma_pos          # losses.py:63
    w_neg = p.clamp(min=0.0) ** gamma_neg                  # losses.py:64
    return (-(w_pos * pos + w_neg * neg)).sum(dim=1).mean()  # losses.py:66-67


def loss_from_config(config: dict):
    """losses.py:74-82 build_loss: reads loss.asymmetric.{gamma_neg,gamma_pos,clip}."""
    a = config.get("loss", {}).get("asymmetric", {})
    gn, gp, cl = float(a.get("gamma_neg", 4)), float(a.get("gamma_pos", 1)), float(a.get("clip", 0.05))
    return lambda logits, targets: asymmetric_focal_loss(logits, targets, gn, gp, cl)
