"""The one call the reference makes into timm — ``timm.create_model(name, pretrained=...,
num_classes=0, drop_path_rate=...)`` at model.py:112-117 — answered by the B200-native backbone.

``model.py`` in this package imports ``create_model`` from here; a maintainer of the reference
would change exactly one line (``import timm`` -> ``import fedvit_b200.timm_b200 as timm``), see
INTEGRATION.md.
"""
from .vit import VisionTransformer, create_model, parse_vit_name  # noqa: F401


def list_models(pattern: str = "") -> list:
    from .vit import _ARCH

    names = [f"{a}_patch16_{s}" for a in _ARCH if a != "vit_micro" for s in (224, 384)]
    needle = pattern.replace("*", "")
    return [n for n in names if needle in n]
