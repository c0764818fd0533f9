#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's OWN code (build container only).

Imports the unmodified /root/reference/{model,losses,utils}.py through oracle/ref_bridge.py (with
oracle/timm answering ``import timm``) and records, for seeded inputs on CPU in fp32:
  * reference ``build_model`` -> state_dict, logits (``ISICClassifier.forward``, model.py:178-207)
  * reference ``build_loss`` -> loss value and d loss / d logits (losses.py:41-67)
  * gradients of every parameter, the clipped global norm (utils.py:192-193)
  * parameters after 2 steps of torch.optim.AdamW over the reference's own LLRD groups
    (model.py:228-270, train.py:253-261) with clip 1.0, and the reference ``EMA`` shadow after them
  * the same for the 4-channel (lesion-mask) variant (model.py:150-166)
The toy size (``vit_micro_patch16_32``: D=64, L=2, 1 head, 32 px) keeps fixtures a few hundred kB;
the arithmetic exercised is the same as ViT-Tiny/Base.

    python tests/golden/make_golden.py        # rewrites the .npz files next to this script
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import ref_bridge  # noqa: E402


def config(masked: bool) -> dict:
    return {
        "model": {"backbone": "vit_micro_patch16_32", "num_classes": 7, "image_size": 32,
                  "pretrained": False, "drop_path_rate": 0.0, "metadata": {"enabled": False},
                  "classifier": {"hidden_dim": 512, "dropout": 0.0}},
        "data": {"use_segmentation_mask": masked},
        "loss": {"asymmetric": {"gamma_neg": 4, "gamma_pos": 1, "clip": 0.05}},
    }


def run(masked: bool) -> dict:
    rm, rl, ru = ref_bridge.load_reference()
    torch.manual_seed(1234 + int(masked))
    cfg = config(masked)
    model = rm.build_model(cfg)
    # give cls_token / biases / norms non-trivial values so every term is exercised
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1 or n.endswith("cls_token"):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    model.train()
    out = {f"state/{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
    c = 4 if masked else 3
    x = torch.randn(6, c, 32, 32, generator=g)
    y = torch.tensor([0, 3, 6, 2, 2, 5])
    out["x"], out["y"] = x.numpy(), y.numpy()
    crit = rl.build_loss(cfg)
    ema = ru.EMA(model, decay=0.9)
    opt = torch.optim.AdamW(rm.get_layerwise_lr_groups(model, base_lr=1e-3, decay_rate=0.75, weight_decay=1e-2),
                            weight_decay=1e-2)
    for step in range(2):
        opt.zero_grad(set_to_none=True)
        logits = model(x)["logits"]
        logits.retain_grad()
        loss = crit(logits, y)
        loss.backward()
        if step == 0:
            out["logits"] = logits.detach().numpy().copy()
            out["loss"] = np.float32(loss.item())
            out["dlogits"] = logits.grad.numpy().copy()
            for n, p in model.named_parameters():
                out[f"grad/{n}"] = p.grad.numpy().copy()
        norm = ru.clip_grad_norm(model.parameters(), 1.0)
        if step == 0:
            out["grad_norm"] = np.float32(float(norm))
        opt.step()
        ema.update()
    for n, p in model.named_parameters():
        out[f"after2/{n}"] = p.detach().numpy().copy()
    for n, v in ema.shadow.items():
        out[f"ema2/{n}"] = v.numpy().copy()
    model.eval()
    with torch.no_grad():
        out["eval_logits_after2"] = model(x)["logits"].numpy().copy()
    return out


def loss_kats() -> dict:
    _, rl, _ = ref_bridge.load_reference()
    crit = rl.build_loss({})
    out = {}
    lg = torch.tensor([[2.0, 0.0, -1.0], [0.5, 0.5, 0.5]])
    out["kat1_logits"], out["kat1_targets"] = lg.numpy(), np.array([0, 2])
    out["kat1_loss"] = np.float32(crit(lg, torch.tensor([0, 2])).item())
    g = torch.Generator().manual_seed(0)
    lg = torch.randn(4, 7, generator=g, dtype=torch.float64).float().requires_grad_(True)
    t = torch.tensor([0, 3, 6, 2])
    l = crit(lg, t)
    l.backward()
    out["kat2_logits"], out["kat2_targets"] = lg.detach().numpy(), t.numpy()
    out["kat2_loss"], out["kat2_dlogits"] = np.float32(l.item()), lg.grad.numpy()
    # saturated / clamped region: big margins make p hit the clip and eps branches
    lg = torch.tensor([[30.0, -30.0, 0.0, 1.0], [-20.0, 25.0, 24.0, -5.0], [0.0, 0.0, 0.0, 0.0]], requires_grad=True)
    t = torch.tensor([1, 1, 3])
    l = crit(lg, t)
    l.backward()
    out["kat3_logits"], out["kat3_targets"] = lg.detach().numpy(), t.numpy()
    out["kat3_loss"], out["kat3_dlogits"] = np.float32(l.item()), lg.grad.numpy()
    return out


if __name__ == "__main__":
    if not ref_bridge.available():
        sys.exit("/root/reference is not mounted: fixtures can only be regenerated in the build container")
    np.savez_compressed(HERE / "micro_rgb.npz", **run(False))
    np.savez_compressed(HERE / "micro_masked.npz", **run(True))
    np.savez_compressed(HERE / "asl_kats.npz", **loss_kats())
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")
