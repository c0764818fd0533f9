"""One client-local optimisation step / epoch — restatement of reference train.py:95-168 with
utils.py:50-105 (EMA), :171-185 (warmup-cosine) and :192-193 (clip). fp32, CPU or any device.

TEST INFRASTRUCTURE (see oracle/__init__.py). MixUp/CutMix and GradScaler are outside the parity
configurations (mixup.alpha 0, cutmix.prob 0; AMP is a no-op off-CUDA, train.py:110).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, List, Optional

import torch


class OracleEMA:
    """utils.py:50-105: shadow = decay*shadow + (1-decay)*param over requires_grad parameters."""

    def __init__(self, model, decay=0.9995):
        self.model, self.decay = model, decay
        self.shadow = {n: p.data.clone() for n, p in model.named_parameters() if p.requires_grad}
        self.backup: Dict[str, torch.Tensor] = {}

    @torch.no_grad()
    def update(self):
        for n, p in self.model.named_parameters():
            if p.requires_grad:
                self.shadow[n].mul_(self.decay).add_(p.data, alpha=1.0 - self.decay)   # utils.py:81


def warmup_cosine_lr(base_lr: float, epoch: int, warmup_epochs: int, total_epochs: int, min_lr: float) -> float:
    """utils.py:179-185."""
    if epoch < warmup_epochs:
        return base_lr * (epoch / max(1, warmup_epochs))
    prog = (epoch - warmup_epochs) / max(1, total_epochs - warmup_epochs)
    return min_lr + (base_lr - min_lr) * 0.5 * (1 + math.cos(math.pi * prog))


def local_epoch(model, batches: Iterable[dict], criterion: Callable, optimizer, grad_clip: float = 1.0,
                accum_steps: int = 1, ema: Optional[OracleEMA] = None, use_meta: bool = False) -> float:
    """train.py:108-168. Returns the sample-weighted mean loss, as the reference does."""
    model.train()
    batches = list(batches)
    running, total = 0.0, 0
    optimizer.zero_grad(set_to_none=True)                                              # train.py:128
    for step, batch in enumerate(batches):
        images, labels = batch["image"], batch["label"]
        meta = batch.get("metadata") if use_meta else None
        logits = model(images, metadata=meta)["logits"]                                # train.py:145-146
        loss = criterion(logits, labels) / accum_steps                                 # train.py:150-151
        loss.backward()                                                                # train.py:153
        if (step + 1) % accum_steps == 0 or (step + 1) == len(batches):                # train.py:155
            torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip)              # train.py:157
            optimizer.step()                                                           # train.py:158
            optimizer.zero_grad(set_to_none=True)                                      # train.py:160
            if ema is not None:
                ema.update()                                                           # train.py:161-162
        bs = images.size(0)
        running += loss.item() * accum_steps * bs                                      # train.py:164
        total += bs
    return running / max(total, 1)
