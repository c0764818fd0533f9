"""Skin-lesion classifier on the B200-native ViT backbone — the reference's ``model.py`` surface.

Same public names, constructor arguments, ``forward(x, metadata=None) -> {"logits"}`` contract,
state_dict layout (``backbone.*`` / ``metadata_branch.net.*`` / ``classifier.*``) and config keys
as reference model.py:27-60 (MetadataBranch), :67-280 (ISICClassifier), :287-324 (factories), so
``train.py`` and checkpoints written by either side interchange. What differs is below the seam:
the backbone is ``fedvit_b200.vit.VisionTransformer`` (hand-written sm_100a kernels, manual
backward) instead of a timm module running on ATen/cuDNN/cuBLAS.

Scope notes: only the ViT family is on this path — the reference's default
SwinV2 backbone name raises. The classifier head (``head.py``) and the 13-d metadata MLP incl. its
BatchNorm1d (``metadata.py``) run on libfedvit kernels as well (row f2 of the scope table).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import timm_b200 as timm
from .head import classifier_head
from .metadata import metadata_embedding

_DEFAULT_BACKBONE = "swinv2_large_window12to24_192to384.ms_in22k_ft_in1k"  # reference model.py:89


class MetadataBranch(nn.Module):
    """age(1) + sex one-hot(3) + site one-hot(9) -> ``output_dim`` embedding.
    Two Linear+BatchNorm1d+GELU stages with dropout after the first (reference model.py:27-60)."""

    def __init__(self, input_dim: int = 13, hidden_dim: int = 256, output_dim: int = 128,
                 dropout: float = 0.4) -> None:
        super().__init__()
        self.output_dim = output_dim
        stages = [
            nn.Linear(input_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(hidden_dim, output_dim), nn.BatchNorm1d(output_dim), nn.GELU(),
        ]
        self.net = nn.Sequential(*stages)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # the Sequential's modules hold the parameters / running statistics (reference state_dict keys
        # metadata_branch.net.{0,1,4,5}.*); the arithmetic is two fused kernels (metadata.py)
        return metadata_embedding(x, self.net, self.training)


class ISICClassifier(nn.Module):
    """backbone features (+ optional metadata embedding) -> Linear-GELU-Dropout-Linear logits."""

    def __init__(
        self,
        backbone_name: str = _DEFAULT_BACKBONE,
        num_classes: int = 8,
        image_size: int = 384,
        in_channels: int = 4,
        pretrained: bool = True,
        drop_path_rate: float = 0.4,
        metadata_enabled: bool = True,
        meta_input_dim: int = 13,
        meta_hidden_dim: int = 256,
        meta_output_dim: int = 128,
        meta_dropout: float = 0.4,
        cls_hidden_dim: int = 512,
        cls_dropout: float = 0.5,
    ) -> None:
        super().__init__()
        self.metadata_enabled = metadata_enabled
        self.num_classes = num_classes
        self.image_size = image_size
        self.in_channels = in_channels

        self.backbone = timm.create_model(backbone_name, pretrained=pretrained, num_classes=0,
                                          drop_path_rate=drop_path_rate)
        self.backbone_dim = self.backbone.num_features
        if self.backbone.patch_embed.img_size[0] != image_size:
            raise ValueError(f"model.image_size={image_size} but {backbone_name} is built for "
                             f"{self.backbone.patch_embed.img_size[0]} px")
        if in_channels != 3:
            self._modify_input_channels(in_channels, pretrained)

        width = self.backbone_dim
        if metadata_enabled:
            self.metadata_branch = MetadataBranch(meta_input_dim, meta_hidden_dim, meta_output_dim, meta_dropout)
            width += meta_output_dim
        self.classifier = nn.Sequential(
            nn.Linear(width, cls_hidden_dim), nn.GELU(), nn.Dropout(cls_dropout),
            nn.Linear(cls_hidden_dim, num_classes),
        )
        self._init_classifier()

    # -- construction helpers -------------------------------------------------------------------
    def _modify_input_channels(self, in_channels: int, pretrained: bool) -> None:
        """Swap the patch projection for one taking ``in_channels`` planes (RGB + lesion mask).
        With pretrained weights the RGB filters are kept and the extra plane gets their mean
        (reference model.py:150-166); the fused patch GEMM reads whatever conv is installed."""
        old = self.backbone.patch_embed.proj
        new = nn.Conv2d(in_channels, old.out_channels, kernel_size=old.kernel_size, stride=old.stride,
                        padding=old.padding, bias=old.bias is not None)
        if pretrained:
            with torch.no_grad():
                new.weight[:, :3].copy_(old.weight)
                new.weight[:, 3:].copy_(old.weight.mean(dim=1, keepdim=True))
                if old.bias is not None:
                    new.bias.copy_(old.bias)
        self.backbone.patch_embed.proj = new

    def _init_classifier(self) -> None:
        for layer in self.classifier:
            if isinstance(layer, nn.Linear):
                nn.init.trunc_normal_(layer.weight, std=0.02)
                nn.init.zeros_(layer.bias)

    # -- forward --------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, metadata: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        feats = self.backbone(x)  # (B, D) fp32, from the sm_100a kernels
        if self.metadata_enabled:
            if metadata is None:  # keep the classifier width fixed (reference model.py:198-203)
                emb = feats.new_zeros(feats.size(0), self.metadata_branch.output_dim)
            else:
                emb = self.metadata_branch(metadata)
            feats = torch.cat([feats, emb.to(feats.dtype)], dim=1)
        # the two head GEMMs (+ GELU, dropout) run on the libfedvit kernels too (head.py, scope row f2)
        return {"logits": classifier_head(feats, self.classifier, self.training)}

    # -- freezing / optimiser groups --------------------------------------------------------------
    def freeze_backbone(self) -> None:
        self.backbone.requires_grad_(False)

    def unfreeze_backbone(self) -> None:
        self.backbone.requires_grad_(True)

    def _head_params(self) -> List[nn.Parameter]:
        ps = list(self.classifier.parameters())
        if self.metadata_enabled:
            ps += list(self.metadata_branch.parameters())
        return ps

    def get_head_parameters(self) -> List[Dict]:
        return [{"params": self._head_params()}]

    def get_layerwise_lr_groups(self, base_lr: float = 1e-4, decay_rate: float = 0.75,
                                weight_decay: float = 1e-5) -> List[Dict]:
        """Layer-wise lr decay exactly as the reference builds it (model.py:228-270): L+3 groups —
        patch_embed at lr*d^(L+1), block i at lr*d^(L-i), final norm at lr, head at 10*lr; weight
        decay on every tensor. cls_token / pos_embed are direct parameters of the backbone and so
        fall in no group: they are never stepped (SURVEY.md quirk list) — kept for parity."""
        blocks = list(self.backbone.blocks)
        depth = len(blocks)

        def group(params, lr):
            return {"params": list(params), "lr": lr, "weight_decay": weight_decay}

        groups = [group(self.backbone.patch_embed.parameters(), base_lr * decay_rate ** (depth + 1))]
        groups += [group(b.parameters(), base_lr * decay_rate ** (depth - i)) for i, b in enumerate(blocks)]
        groups.append(group(self.backbone.norm.parameters(), base_lr))
        groups.append(group(self._head_params(), base_lr * 10))
        return groups

    def count_parameters(self) -> Dict[str, int]:
        def n(m):
            return sum(p.numel() for p in m.parameters())

        out = {"total": n(self), "backbone": n(self.backbone), "classifier": n(self.classifier)}
        if self.metadata_enabled:
            out["metadata"] = n(self.metadata_branch)
        return out


def get_layerwise_lr_groups(model: ISICClassifier, base_lr: float = 1e-4, decay_rate: float = 0.75,
                            weight_decay: float = 1e-5) -> List[Dict]:
    return model.get_layerwise_lr_groups(base_lr, decay_rate, weight_decay)


def count_parameters(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())


def build_model(config: dict) -> ISICClassifier:
    """Config -> model with the reference's keys and defaults (model.py:302-324): ``model.*``,
    ``model.metadata.*``, ``model.classifier.*`` and ``data.use_segmentation_mask`` (4-channel input)."""
    m = config.get("model", {})
    meta = m.get("metadata", {})
    head = m.get("classifier", {})
    masked = config.get("data", {}).get("use_segmentation_mask", False)
    net = ISICClassifier(
        backbone_name=m.get("backbone", _DEFAULT_BACKBONE),
        num_classes=m.get("num_classes", 8),
        image_size=m.get("image_size", 384),
        in_channels=4 if masked else 3,
        pretrained=m.get("pretrained", True),
        drop_path_rate=float(m.get("drop_path_rate", 0.4)),
        metadata_enabled=meta.get("enabled", True),
        meta_input_dim=int(meta.get("input_dim", 13)),
        meta_hidden_dim=int(meta.get("hidden_dim", 256)),
        meta_output_dim=int(meta.get("output_dim", 128)),
        meta_dropout=float(meta.get("dropout", 0.4)),
        cls_hidden_dim=int(head.get("hidden_dim", 512)),
        cls_dropout=float(head.get("dropout", 0.5)),
    )
    # additive key (not in the reference's config.yaml): run the last block's token-wise tail on the cls
    # rows only — see VisionTransformer.cls_only_last_block; off unless asked for
    net.backbone.cls_only_last_block = bool(m.get("cls_only_last_block", False))
    return net
