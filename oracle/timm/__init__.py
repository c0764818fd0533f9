"""Minimal stand-in for the ``timm`` package: just ``create_model`` for the ViT family.

TEST INFRASTRUCTURE (see oracle/__init__.py). The reference calls
``timm.create_model(name, pretrained=..., num_classes=0, drop_path_rate=...)`` at model.py:112-117;
timm (``timm>=0.9``, requirements.txt:3) is not installed in this image, so this restates the
public algorithm of ``timm.models.vision_transformer`` with timm's parameter names.
"""
from .models.vision_transformer import VisionTransformer, create_model, list_models  # noqa: F401

__version__ = "0.9.oracle"
