"""Plain-PyTorch restatement of timm's ``VisionTransformer`` (timm >= 0.9 semantics).

TEST INFRASTRUCTURE (see oracle/__init__.py). Every transformer FLOP of the reference lives in
this third-party class (reference call site model.py:112-117, forward at model.py:193); SURVEY.md
§8.1 lists the algorithm. What matters for parity:

    x -> patch_embed.proj (Conv2d C_in->D, k=s=16, bias) -> flatten(2).transpose(1,2)
      -> cat(cls_token, x) + pos_embed -> pos_drop(p=0)
      -> L x [ x += drop_path(attn(norm1(x))) ; x += drop_path(mlp(norm2(x))) ]   (pre-norm)
      -> norm -> x[:, 0]  (global_pool='token'; fc_norm, head = Identity for num_classes=0)
    LayerNorm eps 1e-6; exact (erf) GELU; attention = softmax(q k^T / sqrt(hd)) v, qkv_bias=True;
    no LayerScale; stochastic depth rate_i = linspace(0, drop_path_rate, L)[i].

state_dict keys are timm's: cls_token, pos_embed, patch_embed.proj.{weight,bias},
blocks.{i}.{norm1,norm2}.{weight,bias}, blocks.{i}.attn.{qkv,proj}.{weight,bias},
blocks.{i}.mlp.{fc1,fc2}.{weight,bias}, norm.{weight,bias}.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F


class DropPath(nn.Module):
    """Per-sample stochastic depth: keep with prob 1-p, rescale kept samples by 1/(1-p)."""

    def __init__(self, drop_prob: float = 0.0) -> None:
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask


class PatchEmbed(nn.Module):
    def __init__(self, img_size: int, patch_size: int, in_chans: int, embed_dim: int) -> None:
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=True)
        self.norm = nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _, _, h, w = x.shape
        assert (h, w) == self.img_size, f"input {h}x{w} != model {self.img_size}"
        x = self.proj(x).flatten(2).transpose(1, 2)  # B, N-1, D
        return self.norm(x)


class Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int) -> None:
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        # == F.scaled_dot_product_attention(q, k, v) (no mask, dropout 0), written out
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = attn.softmax(dim=-1)
        x = attn @ v
        x = x.transpose(1, 2).reshape(b, n, c)
        return self.proj_drop(self.proj(x))


class Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int) -> None:
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class Block(nn.Module):
    def __init__(self, dim: int, num_heads: int, mlp_ratio: float, drop_path: float) -> None:
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.ls1 = nn.Identity()
        self.drop_path1 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = nn.Identity()
        self.drop_path2 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.drop_path1(self.ls1(self.attn(self.norm1(x))))
        x = x + self.drop_path2(self.ls2(self.mlp(self.norm2(x))))
        return x


def _trunc_normal_(t: torch.Tensor, std: float) -> None:
    nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0)  # timm: absolute cut-offs at +-2


class VisionTransformer(nn.Module):
    def __init__(
        self,
        img_size: int = 224,
        patch_size: int = 16,
        in_chans: int = 3,
        num_classes: int = 1000,
        embed_dim: int = 768,
        depth: int = 12,
        num_heads: int = 12,
        mlp_ratio: float = 4.0,
        drop_path_rate: float = 0.0,
    ) -> None:
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_prefix_tokens = 1
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches + 1
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, n, embed_dim) * 0.02)
        self.pos_drop = nn.Dropout(0.0)
        self.norm_pre = nn.Identity()
        dpr = [r.item() for r in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio, dpr[i]) for i in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.init_weights()

    def init_weights(self) -> None:
        _trunc_normal_(self.pos_embed, 0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, 0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        x = self.pos_drop(x + self.pos_embed)
        x = self.norm_pre(x)
        x = self.blocks(x)
        return self.norm(x)

    def forward_head(self, x: torch.Tensor) -> torch.Tensor:
        x = x[:, 0]  # global_pool == 'token'
        x = self.fc_norm(x)
        return self.head(self.head_drop(x))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward_head(self.forward_features(x))


# name -> (embed_dim, depth, heads); patch 16 everywhere, image size parsed from the name
_ARCH: Dict[str, tuple] = {
    "vit_micro": (64, 2, 1),      # test-only toy (not a timm model): fixtures small enough to commit
    "vit_tiny": (192, 12, 3),
    "vit_small": (384, 12, 6),
    "vit_base": (768, 12, 12),
    "vit_large": (1024, 24, 16),
}


def parse_vit_name(name: str):
    """'vit_base_patch16_224.augreg_in21k' -> (768, 12, 12, 16, 224)."""
    base = name.split(".")[0]
    parts = base.split("_")
    if len(parts) != 4 or parts[0] != "vit" or not parts[2].startswith("patch"):
        raise ValueError(f"not a ViT model name this path covers: {name!r}")
    arch = "_".join(parts[:2])
    if arch not in _ARCH:
        raise ValueError(f"unknown ViT size in {name!r}; known: {sorted(_ARCH)}")
    patch = int(parts[2][len("patch"):])
    img = int(parts[3])
    if patch != 16:
        raise ValueError("only patch16 models are covered")
    d, l, h = _ARCH[arch]
    return d, l, h, patch, img


def list_models(pattern: str = "") -> list:
    names = [f"{a}_patch16_{s}" for a in _ARCH for s in (224, 384) if a != "vit_micro"] + ["vit_micro_patch16_32"]
    return [n for n in names if pattern.replace("*", "") in n]


def create_model(model_name: str, pretrained: bool = False, num_classes: int = 1000,
                 drop_path_rate: float = 0.0, in_chans: int = 3, **kwargs) -> VisionTransformer:
    if pretrained:
        raise RuntimeError(
            "pretrained weights cannot be downloaded here (no network); set model.pretrained: false"
        )
    d, l, h, patch, img = parse_vit_name(model_name)
    return VisionTransformer(img_size=img, patch_size=patch, in_chans=in_chans, num_classes=num_classes,
                             embed_dim=d, depth=l, num_heads=h, drop_path_rate=drop_path_rate)
