"""CPU restatement of the reference's batch assembly and MixUp / CutMix. TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).

    assemble()   data.py:148-155 (TF.to_tensor + TF.normalize(IMAGENET_MEAN, IMAGENET_STD);
                 mask: TF.to_tensor then (m - 0.5) / 0.5) and data.py:222-224 (torch.cat to 4 ch)
    mixup()      utils.py:112-121  MixUp.__call__      mixed = lam * x + (1 - lam) * x[idx]
    cutmix()     utils.py:124-150  CutMix.__call__     box from _rand_bbox, pasted from x[idx]
numpy, fp32 step by step so every rounding matches the ATen expression the reference evaluates.
Pinned against the reference's own classes in tests/golden/mix.npz (tests/golden/make_golden_mix.py).
"""
from __future__ import annotations

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # reference data.py:33-34
IMAGENET_STD = (0.229, 0.224, 0.225)


def assemble(img_u8: np.ndarray, mask_u8=None, nhwc: bool = False) -> np.ndarray:
    """uint8 [B,3,H,W] (or [B,H,W,3]) (+ uint8 [B,H,W]) -> fp32 [B,3|4,H,W]."""
    x = img_u8.transpose(0, 3, 1, 2) if nhwc else img_u8
    x = x.astype(np.float32) / np.float32(255.0)                                      # TF.to_tensor
    mean = np.asarray(IMAGENET_MEAN, np.float32).reshape(1, 3, 1, 1)
    std = np.asarray(IMAGENET_STD, np.float32).reshape(1, 3, 1, 1)
    x = ((x - mean) / std).astype(np.float32)                                         # TF.normalize
    if mask_u8 is None:
        return x
    m = mask_u8.astype(np.float32) / np.float32(255.0)
    m = ((m - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)[:, None]
    return np.concatenate([x, m], axis=1)


def mixup(x: np.ndarray, idx: np.ndarray, lam: float) -> np.ndarray:
    lam32, oml32 = np.float32(lam), np.float32(1.0 - lam)
    return (lam32 * x).astype(np.float32) + (oml32 * x[idx]).astype(np.float32)


def rand_bbox(size, lam, cx, cy):
    """utils.py:129-137 with the two np.random.randint draws passed in (cx, cy)."""
    W, H = size[2], size[3]
    cut = np.sqrt(1.0 - lam)
    cw, ch = int(W * cut), int(H * cut)
    x1, y1 = np.clip(cx - cw // 2, 0, W), np.clip(cy - ch // 2, 0, H)
    x2, y2 = np.clip(cx + cw // 2, 0, W), np.clip(cy + ch // 2, 0, H)
    return int(x1), int(y1), int(x2), int(y2)


def cutmix(x: np.ndarray, idx: np.ndarray, box) -> tuple:
    x1, y1, x2, y2 = box
    out = x.copy()
    out[:, :, x1:x2, y1:y2] = x[idx][:, :, x1:x2, y1:y2]
    lam = 1 - ((x2 - x1) * (y2 - y1) / (x.shape[-1] * x.shape[-2]))
    return out, lam
