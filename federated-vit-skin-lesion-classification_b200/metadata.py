"""MetadataBranch on the libfedvit kernels — the remainder of scope row f2.

``Linear(13, 256) -> BatchNorm1d -> GELU -> Dropout(p) -> Linear(256, 128) -> BatchNorm1d -> GELU`` as
the reference builds it (model.py:27-60) and calls it (model.py:195-197). Each Linear + BatchNorm1d + GELU
(+ dropout) stage is ONE kernel forward (``fv_linear_bn_gelu_fwd``: batch statistics, running-stat update,
activation) and one backward (``fv_linear_bn_gelu_bwd``: dgamma / dbeta / dbias / dW and the gradient of
the Linear output); the gradient flowing from stage 2 to stage 1 is one FFMA GEMM. With this the
reference's DEFAULT forward (``metadata.enabled: true``, config.yaml:34-40) launches no cuBLAS and no
ATen batch-norm kernel.

The ``nn.Linear`` / ``nn.BatchNorm1d`` children stay as parameter / buffer holders (same state_dict keys,
so checkpoints and the FedAvg treatment of the running statistics are unchanged); ``num_batches_tracked``
is advanced like ``nn.BatchNorm1d`` does. Parameter gradients are accumulated straight into ``p.grad``
(the FlatArena buffer), like the backbone's. fp32 in both arithmetic modes (autocast keeps batch_norm in
fp32, and a 256 x 13 product is far below tensor-core granularity).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from ._lib import FedVitError
from .vit import _grad_buffer

_K, _MN = ops.MAJOR_K, ops.MAJOR_MN
_E = ops.EPI


def _bn_momentum(bn: nn.BatchNorm1d) -> float:
    if bn.momentum is None:  # cumulative moving average: factor 1 / num_batches_tracked (after the increment)
        return 1.0 / float(int(bn.num_batches_tracked) + 1)
    return float(bn.momentum)


class _MetaFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, lin1: nn.Linear, bn1: nn.BatchNorm1d, p_drop: float, lin2: nn.Linear,
                bn2: nn.BatchNorm1d, training: bool, save: bool, *params):
        B = x.shape[0]
        dev = x.device
        x = x.float().contiguous()
        mask = None
        if training and p_drop > 0.0:
            keep = 1.0 - p_drop
            mask = torch.empty((B, lin1.weight.shape[0]), device=dev, dtype=torch.float32).bernoulli_(keep)
            if keep > 0.0:
                mask.div_(keep)
        stats1 = training or not bn1.track_running_stats
        stats2 = training or not bn2.track_running_stats
        a1 = torch.empty((B, lin1.weight.shape[0]), device=dev, dtype=torch.float32)
        m1, m2 = (_bn_momentum(bn1), _bn_momentum(bn2)) if training else (0.0, 0.0)
        xh1, da1, rs1 = ops.linear_bn_gelu_fwd(x, lin1.weight.detach(), lin1.bias.detach() if lin1.bias is not None else None,
                                               bn1.weight.detach(), bn1.bias.detach(), bn1.running_mean, bn1.running_var,
                                               m1, bn1.eps, stats1, mask, a1, save)
        out = torch.empty((B, lin2.weight.shape[0]), device=dev, dtype=torch.float32)
        xh2, da2, rs2 = ops.linear_bn_gelu_fwd(a1, lin2.weight.detach(), lin2.bias.detach() if lin2.bias is not None else None,
                                               bn2.weight.detach(), bn2.bias.detach(), bn2.running_mean, bn2.running_var,
                                               m2, bn2.eps, stats2, None, out, save)
        if training:
            with torch.no_grad():
                bn1.num_batches_tracked += 1
                bn2.num_batches_tracked += 1
        if save:
            ctx.mods = (lin1, bn1, lin2, bn2)
            ctx.stats = (stats1, stats2)
            ctx.save_for_backward(x, xh1, da1, rs1, a1, xh2, da2, rs2)
            ctx.nparams = len(params)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout: Tensor):
        x, xh1, da1, rs1, a1, xh2, da2, rs2 = ctx.saved_tensors
        lin1, bn1, lin2, bn2 = ctx.mods

        def g(p: Optional[nn.Parameter]):
            return _grad_buffer(p) if p is not None and p.requires_grad else None

        dout = dout.float().contiguous()
        dh2 = ops.linear_bn_gelu_bwd(dout, a1, xh2, da2, rs2, bn2.weight.detach(), ctx.stats[1], g(lin2.weight),
                                     g(lin2.bias), g(bn2.weight), g(bn2.bias))
        # gradient w.r.t. stage 1's output (dropout and GELU' are folded into its saved dact)
        d_a1 = torch.empty_like(a1)
        ops.gemm(dh2, lin2.weight.detach(), None, d_a1, None, _K, _MN, _E["none"], 1, 0)
        ops.linear_bn_gelu_bwd(d_a1, x, xh1, da1, rs1, bn1.weight.detach(), ctx.stats[0], g(lin1.weight),
                               g(lin1.bias), g(bn1.weight), g(bn1.bias))
        return (None,) * 8 + (None,) * ctx.nparams


def metadata_embedding(metadata: Tensor, net: nn.Sequential, training: bool) -> Tensor:
    """emb = net(metadata) for the reference's ``Sequential(Linear, BatchNorm1d, GELU, Dropout, Linear,
    BatchNorm1d, GELU)`` (model.py:41-57)."""
    if not metadata.is_cuda:
        raise FedVitError("fedvit_b200 metadata branch runs on CUDA (sm_100a) only — no CPU/MPS fallback on this path")
    if len(net) != 7:
        raise FedVitError("metadata_embedding: expected the reference's 7-module Sequential")
    lin1, bn1, act1, drop, lin2, bn2, act2 = (net[i] for i in range(7))
    ok = (isinstance(lin1, nn.Linear) and isinstance(bn1, nn.BatchNorm1d) and isinstance(act1, nn.GELU)
          and isinstance(drop, nn.Dropout) and isinstance(lin2, nn.Linear) and isinstance(bn2, nn.BatchNorm1d)
          and isinstance(act2, nn.GELU) and bn1.affine and bn2.affine and bn1.track_running_stats and bn2.track_running_stats
          and getattr(act1, "approximate", "none") == "none" and getattr(act2, "approximate", "none") == "none")
    if not ok:
        raise FedVitError("metadata_embedding: expected Sequential(Linear, BatchNorm1d, GELU(erf), Dropout, Linear, "
                          "BatchNorm1d, GELU(erf)) with affine, running-stat BatchNorm")
    if metadata.dim() != 2 or metadata.shape[1] != lin1.weight.shape[1]:
        raise ValueError(f"metadata must be [B, {lin1.weight.shape[1]}], got {tuple(metadata.shape)}")
    params = [p for p in net.parameters()]
    save = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _MetaFunction.apply(metadata, lin1, bn1, float(drop.p), lin2, bn2, bool(training), save, *params)
