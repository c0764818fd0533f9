"""Training utilities of the path — the reference's ``utils.py`` names that ``train.py`` imports
for it (EMA, clip_grad_norm, WarmupCosineScheduler, seed_everything, get_device, load_config,
checkpoint save/load), re-implemented over the flat arena.

Out of scope here: auto batch-size probing, the
sklearn metric tables — they are host-side data/driver code, not the hot path.
"""
from __future__ import annotations

import math
import os
import random
from typing import Dict, Iterable, Optional

import numpy as np
import torch
import torch.nn as nn
from torch.optim.lr_scheduler import LRScheduler

from . import ops
from .arena import FlatArena, arena_of


def seed_everything(seed: int = 42) -> None:
    """Seeds python / numpy / torch (reference utils.py:25-33)."""
    os.environ["PYTHONHASHSEED"] = str(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def get_device(device_str: str = "auto") -> torch.device:
    """CUDA only: the MPS / CPU branches of reference utils.py:36-43 are out of scope by design."""
    if device_str in ("auto", "cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("fedvit_b200 needs a CUDA (sm_100a) device; no MPS/CPU fallback on this path")
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device_str)
    if dev.type != "cuda":
        raise RuntimeError(f"device {device_str!r} is not supported: CUDA (sm_100a) only")
    return dev


def load_config(path: str) -> dict:
    import yaml

    with open(path) as f:
        return yaml.safe_load(f)


# ----------------------------------------------------------------------------------------------
# EMA
# ----------------------------------------------------------------------------------------------
class EMA:
    """Exponential moving average of the trainable parameters — API of reference utils.py:50-105
    (``update`` / ``apply_shadow`` / ``restore`` / ``state_dict`` / ``load_state_dict``).

    The shadow is ONE flat fp32 buffer mirroring the parameter arena, so ``update`` is a single
    HBM sweep (12 B/param) instead of a Python loop of ~2x150 launches — or no extra sweep at all
    once attached to a ``FusedAdamW`` (``attach``), which folds it into the optimiser kernel.
    ``apply_shadow`` / ``restore`` are two flat copies."""

    def __init__(self, model: nn.Module, decay: float = 0.9995) -> None:
        self.model = model
        self.decay = decay
        arena = arena_of(model)
        if arena is None:
            arena = FlatArena(model)
        self.arena = arena
        self.flat = arena.params.clone()
        self._backup: Optional[torch.Tensor] = None
        self._fused_updates = 0
        self._seen_fused = 0

    def attach(self, optimizer) -> "EMA":
        optimizer.ema = self
        return self

    @torch.no_grad()
    def update(self) -> None:
        if self._fused_updates > self._seen_fused:  # the optimiser sweep already did this step's update
            self._seen_fused = self._fused_updates
            return
        ops.ema_update(self.flat, self.arena.params, self.decay)

    @torch.no_grad()
    def apply_shadow(self) -> None:
        self._backup = self.arena.params.clone()
        self.arena.params.copy_(self.flat)
        self._touch()

    @torch.no_grad()
    def restore(self) -> None:
        if self._backup is not None:
            self.arena.params.copy_(self._backup)
            self._backup = None
            self._touch()

    def _touch(self) -> None:
        if self.arena.lp is not None:
            self.arena.refresh_lp(force=True)

    @property
    def shadow(self) -> Dict[str, torch.Tensor]:
        return {n: self.arena.view(self.flat, n) for n, p in zip(self.arena.names, self.arena._params)
                if p.requires_grad}

    def state_dict(self) -> dict:
        return {"shadow": {k: v.detach().cpu().clone() for k, v in self.shadow.items()}, "decay": self.decay}

    def load_state_dict(self, sd: dict) -> None:
        for k, v in sd["shadow"].items():
            self.arena.view(self.flat, k).copy_(v)
        self.decay = sd.get("decay", self.decay)


# ----------------------------------------------------------------------------------------------
# schedule, clipping
# ----------------------------------------------------------------------------------------------
class WarmupCosineScheduler(LRScheduler):
    """Linear warm-up over ``warmup_epochs`` then cosine decay to ``min_lr``; stepped once per
    epoch — per FedAvg round in the federated loop (reference utils.py:171-185, train.py:297)."""

    def __init__(self, optimizer, warmup_epochs: int, total_epochs: int, min_lr: float = 1e-6,
                 last_epoch: int = -1) -> None:
        self.warmup_epochs = warmup_epochs
        self.total_epochs = total_epochs
        self.min_lr = min_lr
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        e = self.last_epoch
        if e < self.warmup_epochs:
            f = e / max(1, self.warmup_epochs)
            return [b * f for b in self.base_lrs]
        t = (e - self.warmup_epochs) / max(1, self.total_epochs - self.warmup_epochs)
        c = 0.5 * (1.0 + math.cos(math.pi * t))
        return [self.min_lr + (b - self.min_lr) * c for b in self.base_lrs]


def clip_grad_norm(parameters: Iterable[nn.Parameter], max_norm: float = 1.0, optimizer=None) -> torch.Tensor:
    """Global L2 norm over ALL given gradients (incl. the never-stepped cls_token / pos_embed, as
    in the reference: utils.py:192-193 on ``model.parameters()``, train.py:157) and clipping by
    ``min(1, max_norm / (norm + 1e-6))``. Returns the norm as a device scalar (no host sync).

    With a ``FusedAdamW`` passed as ``optimizer`` the rescale is not a pass of its own: the
    coefficient is applied when the optimiser sweep reads the gradients."""
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return torch.zeros(())
    arena = getattr(params[0], "_fv_arena", (None,))[0]
    in_arena = arena is not None and all(
        getattr(p, "_fv_arena", (None,))[0] is arena and p.grad.data_ptr() == arena.grad_view(p).data_ptr()
        for p in params)
    if in_arena and len(params) == sum(1 for q in arena._params if q.grad is not None):
        sumsq = torch.zeros(1, device=arena.device, dtype=torch.float32)
        ops.sumsq(arena.grads, sumsq, False)
        if optimizer is not None and hasattr(optimizer, "defer_clip"):
            optimizer.defer_clip(sumsq, max_norm)
        else:
            ops.scale_by_clip(arena.grads, sumsq, float(max_norm))
        return sumsq.sqrt().squeeze(0)
    # gradients living outside one arena: per-tensor sums of squares into one device scalar
    dev = params[0].grad.device
    sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
    flats = []
    for p in params:
        g = p.grad.contiguous().view(-1)
        if g.numel() % 4 or g.data_ptr() % 16:
            pad = torch.zeros((g.numel() + 3) // 4 * 4, device=dev, dtype=torch.float32)
            pad[: g.numel()] = g
            g = pad
        flats.append((p, g))
        ops.sumsq(g, sumsq, True)
    coef = (max_norm / (sumsq.sqrt() + 1e-6)).clamp(max=1.0)
    for p, _ in flats:
        p.grad.mul_(coef.to(p.grad.dtype))
    return sumsq.sqrt().squeeze(0)


# ----------------------------------------------------------------------------------------------
# checkpoint (same dict layout as reference utils.py:287-308, so files interchange)
# ----------------------------------------------------------------------------------------------
def save_checkpoint(model, optimizer, scheduler, ema, epoch, metric, path, config=None) -> None:
    torch.save({
        "epoch": epoch,
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict() if optimizer else None,
        "scheduler_state_dict": scheduler.state_dict() if scheduler else None,
        "ema_state_dict": ema.state_dict() if ema else None,
        "best_metric": metric,
        "config": config,
    }, path)


def load_checkpoint(path, model, optimizer=None, scheduler=None, ema=None, device=None):
    ckpt = torch.load(path, map_location=device or "cpu", weights_only=False)
    model.load_state_dict(ckpt["model_state_dict"])
    if optimizer and ckpt.get("optimizer_state_dict"):
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    if scheduler and ckpt.get("scheduler_state_dict"):
        scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    if ema and ckpt.get("ema_state_dict"):
        ema.load_state_dict(ckpt["ema_state_dict"])
    return ckpt


# ================================================================================================
# MixUp / CutMix on the device (scope row f3) — reference utils.py:112-170, same names
# ================================================================================================
class MixUp:
    """``mixed = lam * images + (1 - lam) * images[idx]`` (reference utils.py:112-121) as ONE pass
    of ``fv_mix_batch`` instead of three ATen kernels + a gather; same random draws in the same
    order (np.random.beta, torch.randperm on the images' device), bit-identical output."""

    def __init__(self, alpha: float = 0.4):
        self.alpha = alpha

    def __call__(self, images: torch.Tensor, labels: torch.Tensor):
        lam = np.random.beta(self.alpha, self.alpha) if self.alpha > 0 else 1.0
        idx = torch.randperm(images.size(0), device=images.device)
        mixed = ops.mix_batch(images.float().contiguous(), idx, float(lam), ops.MIX_MIXUP, [0, 0, 0, 0])
        return mixed, labels, labels[idx], lam


class CutMix:
    """Reference utils.py:124-150: a random box (rows x1:x2 of dim 2, columns y1:y2 of dim 3 — the
    reference's naming) is pasted from the permuted batch; lam is recomputed from the box area."""

    def __init__(self, alpha: float = 1.0, prob: float = 0.7):
        self.alpha = alpha
        self.prob = prob

    @staticmethod
    def _rand_bbox(size, lam):
        W, H = size[2], size[3]
        cut = np.sqrt(1. - lam)
        cw, ch = int(W * cut), int(H * cut)
        cx, cy = np.random.randint(W), np.random.randint(H)
        x1, y1 = np.clip(cx - cw // 2, 0, W), np.clip(cy - ch // 2, 0, H)
        x2, y2 = np.clip(cx + cw // 2, 0, W), np.clip(cy + ch // 2, 0, H)
        return int(x1), int(y1), int(x2), int(y2)

    def __call__(self, images: torch.Tensor, labels: torch.Tensor):
        if np.random.rand() > self.prob:
            return images, labels, labels, 1.0
        lam = np.random.beta(self.alpha, self.alpha)
        idx = torch.randperm(images.size(0), device=images.device)
        x1, y1, x2, y2 = self._rand_bbox(images.size(), lam)
        mixed = ops.mix_batch(images.float().contiguous(), idx, 1.0, ops.MIX_CUTMIX, [x1, y1, x2, y2])
        lam = 1 - ((x2 - x1) * (y2 - y1) / (images.size(-1) * images.size(-2)))
        return mixed, labels, labels[idx], lam


class MixupCutmix:
    """Randomly choose MixUp or CutMix each batch (reference utils.py:153-164)."""

    def __init__(self, mixup_alpha=0.4, cutmix_alpha=1.0, cutmix_prob=0.7):
        self.mixup = MixUp(alpha=mixup_alpha)
        self.cutmix = CutMix(alpha=cutmix_alpha, prob=1.0)
        self.cutmix_prob = cutmix_prob

    def __call__(self, images, labels):
        if np.random.rand() < self.cutmix_prob:
            return self.cutmix(images, labels)
        return self.mixup(images, labels)


def mixup_criterion(criterion, logits, labels_a, labels_b, lam):
    """Reference utils.py:167-168."""
    return lam * criterion(logits, labels_a) + (1 - lam) * criterion(logits, labels_b)


# ================================================================================================
# evaluation, plain and with test-time augmentation (scope row f4, first half) — reference
# utils.py:200-280, same names / signatures / returned keys
# ================================================================================================
@torch.no_grad()
def evaluate_with_tta(model: nn.Module, loader, device: torch.device, use_metadata: bool = True,
                      use_amp: bool = True):
    """TTA: the loader yields ``images`` of shape (B, T, C, H, W) (T augmented views, built on the
    host); the views run as ONE forward of batch B*T on the sm_100a kernels and their logits are
    averaged (reference utils.py:200-230). Returns (preds, labels, logits[B_total, C])."""
    model.eval()
    all_preds, all_labels, all_logits = [], [], []
    for batch in loader:
        images = batch["images"]
        labels = batch["label"]
        B, T = images.shape[:2]
        flat = images.reshape(-1, *images.shape[2:]).to(device, non_blocking=True)
        meta = None
        if use_metadata and "metadata" in batch:
            meta = batch["metadata"].to(device, non_blocking=True)
            meta = meta.unsqueeze(1).expand(-1, T, -1).reshape(B * T, -1)
        with torch.amp.autocast(device_type=device.type, enabled=use_amp and device.type == "cuda",
                                dtype=torch.bfloat16):
            logits_flat = model(flat, metadata=meta)["logits"]
        logits = logits_flat.float().view(B, T, -1).mean(dim=1)
        all_preds.extend(logits.argmax(dim=1).cpu().tolist())
        all_labels.extend(torch.as_tensor(labels).tolist())
        all_logits.append(logits.cpu().numpy())
    return all_preds, all_labels, np.concatenate(all_logits, axis=0)


@torch.no_grad()
def evaluate(model: nn.Module, loader, device: torch.device, use_metadata: bool = True, use_amp: bool = True) -> Dict:
    """Standard (no-TTA) evaluation with the fused cross-entropy kernel (reference utils.py:237-280):
    loss, accuracy, balanced accuracy, macro-F1, confusion matrix, per-class recall, predictions."""
    model.eval()
    loss_sum = torch.zeros((), device=device, dtype=torch.float32)
    preds, gold = [], []
    num_classes = None
    for batch in loader:
        images = batch["image"].to(device, non_blocking=True)
        labels = batch["label"].to(device, non_blocking=True)
        meta = batch.get("metadata")
        if meta is not None:
            meta = meta.to(device, non_blocking=True)
        with torch.amp.autocast(device_type=device.type, enabled=use_amp and device.type == "cuda",
                                dtype=torch.bfloat16):
            logits = model(images, metadata=meta if use_metadata else None)["logits"]
        loss, _ = ops.ce_loss(logits.float().contiguous(), labels)
        loss_sum += loss.reshape(()) * images.size(0)
        num_classes = logits.shape[1]
        preds.append(logits.argmax(1))
        gold.append(labels)
    p = torch.cat(preds).cpu().numpy() if preds else np.zeros(0, np.int64)
    y = torch.cat(gold).cpu().numpy() if gold else np.zeros(0, np.int64)
    total = max(len(y), 1)
    nc = int(num_classes or 1)
    cm = np.zeros((nc, nc), dtype=np.int64)
    np.add.at(cm, (y, p), 1)
    support = cm.sum(1)
    recall = [float(cm[i, i] / support[i]) if support[i] > 0 else 0.0 for i in range(nc)]
    present = support > 0
    tp = np.diag(cm).astype(np.float64)
    predicted = cm.sum(0).astype(np.float64)
    prec = np.divide(tp, predicted, out=np.zeros_like(tp), where=predicted > 0)
    rec = np.divide(tp, support.astype(np.float64), out=np.zeros_like(tp), where=present)
    den = prec + rec
    f1 = np.divide(2 * prec * rec, den, out=np.zeros_like(tp), where=den > 0)
    return {
        "loss": float(loss_sum.item()) / total,
        "accuracy": float(tp.sum() / total),
        "balanced_accuracy": float(rec[present].mean()) if present.any() else 0.0,
        # sklearn f1_score(average="macro", zero_division=0): mean over the labels seen in y or p
        "macro_f1": float(f1[present | (predicted > 0)].mean()) if (present | (predicted > 0)).any() else 0.0,
        "confusion_matrix": cm,
        "per_class_recall": recall,
        "all_preds": p.tolist(),
        "all_labels": y.tolist(),
    }
