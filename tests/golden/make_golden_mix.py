#!/usr/bin/env python
"""Generate tests/golden/mix.npz from the reference's OWN utils.MixUp / utils.CutMix and the
torchvision calls its Dataset makes (build container only: imports /root/reference/utils.py).

    python tests/golden/make_golden_mix.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torchvision.transforms.functional as TF
from PIL import Image

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import ref_bridge  # noqa: E402


def main() -> None:
    _, _, ru = ref_bridge.load_reference()
    out = {}
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 4, 32, 32, generator=g)
    y = torch.tensor([0, 3, 6, 2, 2, 5])
    out["x"], out["y"] = x.numpy(), y.numpy()

    # MixUp (utils.py:112-121): np.random.beta for lam, torch.randperm for the partners
    np.random.seed(11)
    torch.manual_seed(11)
    mixed, la, lb, lam = ru.MixUp(alpha=0.4)(x, y)
    np.random.seed(11)
    torch.manual_seed(11)
    lam_again = np.random.beta(0.4, 0.4)
    idx = torch.randperm(6)
    assert lam == lam_again and torch.equal(lb, y[idx])
    out["mixup/out"], out["mixup/lam"], out["mixup/idx"] = mixed.numpy(), np.float64(lam), idx.numpy()

    # CutMix (utils.py:124-150): rand (prob gate), beta, randperm, two randints
    np.random.seed(12)
    torch.manual_seed(12)
    mixed, la, lb, lam = ru.CutMix(alpha=1.0, prob=1.0)(x, y)
    np.random.seed(12)
    torch.manual_seed(12)
    _gate = np.random.rand()
    lam0 = np.random.beta(1.0, 1.0)
    idx = torch.randperm(6)
    cx, cy = np.random.randint(32), np.random.randint(32)
    assert torch.equal(lb, y[idx])
    out["cutmix/out"], out["cutmix/lam_out"], out["cutmix/idx"] = mixed.numpy(), np.float64(lam), idx.numpy()
    out["cutmix/lam0"], out["cutmix/cx"], out["cutmix/cy"] = np.float64(lam0), np.int64(cx), np.int64(cy)

    # Dataset tensor work (data.py:148-155, 222-224) on uint8 pixels through PIL, as the reference does
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(5, 24, 32, 3), dtype=np.uint8)      # NHWC like PIL
    mask = (rng.random((5, 24, 32)) < 0.3).astype(np.uint8) * 255
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    ts = []
    for i in range(5):
        it = TF.normalize(TF.to_tensor(Image.fromarray(img[i])), mean, std)
        mt = (TF.to_tensor(Image.fromarray(mask[i], mode="L")) - 0.5) / 0.5
        ts.append(torch.cat([it, mt], dim=0))
    out["asm/img_u8"], out["asm/mask_u8"] = img, mask
    out["asm/out"] = torch.stack(ts).numpy()
    np.savez_compressed(HERE / "mix.npz", **out)
    print("wrote", HERE / "mix.npz", {k: getattr(v, "shape", ()) for k, v in out.items()})


if __name__ == "__main__":
    main()
