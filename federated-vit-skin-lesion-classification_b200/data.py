"""Synthetic client shards with the reference's batch contract.

The reference's ``data.py`` (ISIC-2019 CSV parsing, PIL augmentation, samplers) is CPU image I/O
and out of scope. What the hot path needs from it is the batch format it yields
(data.py:226-234, consumed at train.py:132-136): a dict with ``"image"`` (B, C, H, W) float32,
``"label"`` (B,) int64 and optionally ``"metadata"`` (B, 13). This module produces that from a
seeded generator: images ~ N(0,1) (the range ``Normalize`` leaves, data.py:149-150),
a 4th mask plane in {-1,+1} on the masked path (data.py:153-154), labels uniform or
Dirichlet(alpha)-skewed per client for the non-IID configuration.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional

import numpy as np
import torch


def client_label_probs(num_clients: int, num_classes: int, partition: str = "iid",
                       alpha: float = 0.5, seed: int = 42) -> np.ndarray:
    if partition == "iid":
        return np.full((num_clients, num_classes), 1.0 / num_classes)
    if partition == "dirichlet":
        rng = np.random.default_rng(seed)
        return rng.dirichlet(np.full(num_classes, alpha), size=num_clients)
    raise ValueError(f"unknown federated.partition {partition!r} (iid | dirichlet)")


class SyntheticClientLoader:
    """Iterable of ``{"image", "label"[, "metadata"]}`` batches for one client.

    ``n_samples`` is the client's shard size n_k (``drop_last`` as the reference, data.py:468, so
    an epoch is n_k // batch_size steps). ``pool`` bounds how many distinct images are materialised
    (the epoch cycles through them); tensors live in pinned host memory unless ``device`` is given,
    in which case the whole pool is HBM-resident and batches are views (no copies in the loop).
    """

    def __init__(self, client_id: int, n_samples: int, batch_size: int, image_size: int,
                 channels: int = 3, num_classes: int = 7, label_probs: Optional[np.ndarray] = None,
                 metadata_dim: int = 0, pool: Optional[int] = None,
                 device: Optional[torch.device] = None, pin: bool = True) -> None:
        self.client_id = client_id
        self.n_samples = int(n_samples)
        self.batch_size = int(batch_size)
        self.steps = self.n_samples // self.batch_size
        if self.steps < 1:
            raise ValueError(f"client {client_id}: n_samples={n_samples} < batch_size={batch_size}")
        pool = self.steps * self.batch_size if pool is None else max(self.batch_size, min(pool, self.n_samples))
        pool = pool // self.batch_size * self.batch_size
        g = torch.Generator().manual_seed(1000 + client_id)
        img = torch.randn(pool, channels, image_size, image_size, generator=g)
        if channels == 4:  # lesion-mask plane: 0/255 PNG -> (x - 0.5) / 0.5
            img[:, 3] = (torch.bernoulli(torch.full((pool, image_size, image_size), 0.3), generator=g) - 0.5) / 0.5
        if label_probs is None:
            lab = torch.randint(0, num_classes, (pool,), generator=g)
        else:
            lab = torch.multinomial(torch.as_tensor(label_probs, dtype=torch.float64), pool,
                                    replacement=True, generator=g)
        meta = torch.rand(pool, metadata_dim, generator=g) if metadata_dim else None
        if device is not None:
            img, lab = img.to(device), lab.to(device)
            meta = meta.to(device) if meta is not None else None
        elif pin and torch.cuda.is_available():
            img, lab = img.pin_memory(), lab.pin_memory()
            meta = meta.pin_memory() if meta is not None else None
        self.images, self.labels, self.meta = img, lab, meta
        self.pool_batches = pool // self.batch_size

    def __len__(self) -> int:
        return self.steps

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        bs = self.batch_size
        for s in range(self.steps):
            i = (s % self.pool_batches) * bs
            batch = {"image": self.images[i:i + bs], "label": self.labels[i:i + bs]}
            if self.meta is not None:
                batch["metadata"] = self.meta[i:i + bs]
            yield batch

    def bytes_per_step(self) -> int:
        b = self.images[: self.batch_size].numel() * 4 + self.batch_size * 8
        if self.meta is not None:
            b += self.meta[: self.batch_size].numel() * 4
        return b


def client_sizes(config: dict) -> List[int]:
    """n_k per client from ``federated.samples_per_client`` (int, or a list for unequal shards)."""
    fed = config.get("federated", {})
    k = int(fed.get("num_clients", 1))
    spc = fed.get("samples_per_client", 64)
    if isinstance(spc, (list, tuple)):
        if len(spc) != k:
            raise ValueError("federated.samples_per_client list must have num_clients entries")
        return [int(s) for s in spc]
    return [int(spc)] * k


IMAGENET_MEAN = (0.485, 0.456, 0.406)  # reference data.py:33-34
IMAGENET_STD = (0.229, 0.224, 0.225)


class DeviceBatchAssembler:
    """Batch assembly on the GPU (scope row f3): a loader that yields raw ``uint8`` pixels
    (``image_u8`` [B,3,H,W] or [B,H,W,3], optional ``mask_u8`` [B,H,W]) instead of normalised fp32
    tensors ships a quarter of the bytes over PCIe; ``fv_assemble_batch`` then does what the
    reference Dataset does per sample on the host — ``TF.to_tensor`` + ``TF.normalize`` + mask to
    +-1 + 4-channel concat (reference data.py:148-155, 222-224) — in one pass, writing the NCHW fp32
    batch the patch-embedding TMA map reads. Batches that already carry ``image`` pass through.

        for batch in loader:                       # host dicts, uint8
            batch = assembler(batch)               # device dict with "image" fp32 [B,C,H,W]
    """

    def __init__(self, device: torch.device, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> None:
        self.device = torch.device(device)
        self.mean = [float(v) for v in mean]
        self.std = [float(v) for v in std]

    def __call__(self, batch: Dict) -> Dict:
        from . import ops

        if "image_u8" not in batch:
            return batch
        out = {k: v for k, v in batch.items() if k not in ("image_u8", "mask_u8")}
        img = batch["image_u8"].to(self.device, non_blocking=True).contiguous()
        mask = batch.get("mask_u8")
        if mask is not None:
            mask = mask.to(self.device, non_blocking=True).contiguous()
        out["image"] = ops.assemble_batch(img, mask, self.mean, self.std, None, 1.0, ops.MIX_NONE, [0, 0, 0, 0])
        return out
