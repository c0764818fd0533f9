import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_rgb():
    return dict(np.load(GOLDEN / "micro_rgb.npz"))


@pytest.fixture(scope="session")
def golden_masked():
    return dict(np.load(GOLDEN / "micro_masked.npz"))


@pytest.fixture(scope="session")
def asl_kats():
    return dict(np.load(GOLDEN / "asl_kats.npz"))


def micro_config(masked: bool = False, **over) -> dict:
    cfg = {
        "seed": 42,
        "model": {"backbone": "vit_micro_patch16_32", "num_classes": 7, "image_size": 32,
                  "pretrained": False, "drop_path_rate": 0.0, "metadata": {"enabled": False},
                  "classifier": {"hidden_dim": 512, "dropout": 0.0}},
        "data": {"use_segmentation_mask": masked},
        "training": {"use_amp": False, "grad_clip": 1.0, "gradient_accumulation_steps": 1, "batch_size": 6,
                     "optimizer": {"lr": 1e-3, "weight_decay": 1e-2}, "llrd": {"enabled": True, "decay_rate": 0.75}},
        "augmentation": {"mixup": {"alpha": 0.0}, "cutmix": {"prob": 0.0}},
        "loss": {"asymmetric": {"gamma_neg": 4, "gamma_pos": 1, "clip": 0.05}},
    }
    for k, v in over.items():
        cfg[k] = v
    return cfg


def state_from_golden(g: dict) -> dict:
    return {k[len("state/"):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith("state/")}


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
