"""ISICClassifier — restatement of reference model.py (head, metadata MLP, forward, LLRD groups).

TEST INFRASTRUCTURE (see oracle/__init__.py). The backbone comes from oracle/timm (the reference
gets it from timm.create_model, model.py:112-117).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .timm.models.vision_transformer import create_model


class OracleMetadataMLP(nn.Module):
    """model.py:27-60: Linear-BN-GELU-Dropout-Linear-BN-GELU over the 13-d metadata vector."""

    def __init__(self, input_dim=13, hidden_dim=256, output_dim=128, dropout=0.4):
        super().__init__()
        self.output_dim = output_dim
        self.net = nn.Sequential(
            nn.Linear(input_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(hidden_dim, output_dim), nn.BatchNorm1d(output_dim), nn.GELU(),
        )

    def forward(self, x):
        return self.net(x)


class OracleClassifier(nn.Module):
    """model.py:67-207. Attribute names (backbone / metadata_branch / classifier) match the
    reference so state_dicts interchange."""

    def __init__(self, backbone_name, num_classes=8, image_size=384, in_channels=4, pretrained=False,
                 drop_path_rate=0.4, metadata_enabled=True, meta_input_dim=13, meta_hidden_dim=256,
                 meta_output_dim=128, meta_dropout=0.4, cls_hidden_dim=512, cls_dropout=0.5):
        super().__init__()
        self.metadata_enabled = metadata_enabled
        self.backbone = create_model(backbone_name, pretrained=pretrained, num_classes=0,
                                     drop_path_rate=drop_path_rate)                 # model.py:112-117
        self.backbone_dim = self.backbone.num_features                               # model.py:119
        if in_channels != 3:                                                         # model.py:123-124
            old = self.backbone.patch_embed.proj
            new = nn.Conv2d(in_channels, old.out_channels, kernel_size=old.kernel_size,
                            stride=old.stride, padding=old.padding, bias=old.bias is not None)
            if pretrained:                                                           # model.py:159-164
                with torch.no_grad():
                    new.weight[:, :3] = old.weight
                    new.weight[:, 3:] = old.weight.mean(dim=1, keepdim=True)
                    if old.bias is not None:
                        new.bias.copy_(old.bias)
            self.backbone.patch_embed.proj = new
        if metadata_enabled:                                                         # model.py:127-136
            self.metadata_branch = OracleMetadataMLP(meta_input_dim, meta_hidden_dim, meta_output_dim, meta_dropout)
            cin = self.backbone_dim + meta_output_dim
        else:
            cin = self.backbone_dim
        self.classifier = nn.Sequential(                                             # model.py:139-144
            nn.Linear(cin, cls_hidden_dim), nn.GELU(), nn.Dropout(cls_dropout), nn.Linear(cls_hidden_dim, num_classes))
        for m in self.classifier.modules():                                          # model.py:168-173
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def forward(self, x, metadata: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        f = self.backbone(x)                                                         # model.py:193
        if self.metadata_enabled:                                                    # model.py:195-204
            if metadata is not None:
                e = self.metadata_branch(metadata)
            else:
                e = torch.zeros(f.size(0), self.metadata_branch.output_dim, device=f.device, dtype=f.dtype)
            f = torch.cat([f, e], dim=1)
        return {"logits": self.classifier(f)}                                        # model.py:206-207


def llrd_groups(model: OracleClassifier, base_lr=1e-4, decay_rate=0.75, weight_decay=1e-5) -> List[Dict]:
    """model.py:228-270: patch_embed lr*d^(L+1), block i lr*d^(L-i), norm lr, head 10*lr; wd on
    every group; cls_token / pos_embed land in NO group."""
    bb = model.backbone
    blocks = list(bb.blocks)
    n = len(blocks)
    groups = [{"params": list(bb.patch_embed.parameters()), "lr": base_lr * decay_rate ** (n + 1),
               "weight_decay": weight_decay}]
    for i, blk in enumerate(blocks):
        groups.append({"params": list(blk.parameters()), "lr": base_lr * decay_rate ** (n - i),
                       "weight_decay": weight_decay})
    groups.append({"params": list(bb.norm.parameters()), "lr": base_lr, "weight_decay": weight_decay})
    head = list(model.classifier.parameters())
    if model.metadata_enabled:
        head += list(model.metadata_branch.parameters())
    groups.append({"params": head, "lr": base_lr * 10, "weight_decay": weight_decay})
    return groups


def model_from_config(config: dict) -> OracleClassifier:
    """model.py:302-324 build_model."""
    m, d = config.get("model", {}), config.get("data", {})
    meta, cls = m.get("metadata", {}), m.get("classifier", {})
    return OracleClassifier(
        backbone_name=m.get("backbone"), num_classes=m.get("num_classes", 8),
        image_size=m.get("image_size", 384), in_channels=4 if d.get("use_segmentation_mask", False) else 3,
        pretrained=m.get("pretrained", True), drop_path_rate=float(m.get("drop_path_rate", 0.4)),
        metadata_enabled=meta.get("enabled", True), meta_input_dim=int(meta.get("input_dim", 13)),
        meta_hidden_dim=int(meta.get("hidden_dim", 256)), meta_output_dim=int(meta.get("output_dim", 128)),
        meta_dropout=float(meta.get("dropout", 0.4)), cls_hidden_dim=int(cls.get("hidden_dim", 512)),
        cls_dropout=float(cls.get("dropout", 0.5)))
