"""timm-compatible ``VisionTransformer`` whose forward AND backward run on the libfedvit kernels.

This is the object the reference obtains from ``timm.create_model(...)`` at model.py:112-117 and
calls at model.py:193. It keeps everything model.py touches: ``num_features``,
a re-assignable ``patch_embed.proj`` (``nn.Conv2d``), iterable ``blocks``, ``norm``, and timm's
state_dict keys — so checkpoints interchange and ``_modify_input_channels`` /
``get_layerwise_lr_groups`` work unchanged.

The ``nn.Linear`` / ``nn.LayerNorm`` / ``nn.Conv2d`` children are *parameter holders only*; their
own forwards are never called. ``forward`` is one ``torch.autograd.Function`` over the whole
backbone with a hand-written backward:

  forward  per block: LN -> qkv GEMM(+bias) -> flash attention -> proj GEMM(+bias+residual)
                      -> LN -> fc1 GEMM(+bias, GELU, also emits GELU') -> fc2 GEMM(+bias+residual)
  backward per block: dgrad GEMMs (fc2's fused with GELU'), wgrad GEMMs (split-K, accumulate into
                      the fp32 gradient arena), column-sum bias grads, fused LN backward that also
                      adds the residual-path gradient and emits the bf16 copy the next GEMM reads.

Arithmetic modes (chosen per call, like the reference's ``torch.amp.autocast`` at train.py:144):
  * autocast active  -> bf16 operands on tcgen05 tensor cores, fp32 accumulate, fp32 residual
    stream / LN statistics / softmax (SURVEY.md Appendix B semantics);
  * otherwise        -> fp32 everywhere (FFMA GEMM kernel) — the 1e-4 parity configuration.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from ._lib import FedVitError

_K, _MN = ops.MAJOR_K, ops.MAJOR_MN
_E = ops.EPI

_ARCH = {
    "vit_micro": (64, 2, 1),  # test-only toy size (matches oracle/timm)
    "vit_tiny": (192, 12, 3),
    "vit_small": (384, 12, 6),
    "vit_base": (768, 12, 12),
    "vit_large": (1024, 24, 16),
}


def parse_vit_name(name: str) -> Tuple[int, int, int, int, int]:
    """'vit_base_patch16_224.augreg_in21k' -> (embed_dim, depth, heads, patch, img)."""
    parts = name.split(".")[0].split("_")
    if len(parts) != 4 or parts[0] != "vit" or not parts[2].startswith("patch"):
        raise ValueError(
            f"{name!r}: this path covers timm ViT names 'vit_<size>_patch16_<img>' "
            f"(sizes: {sorted(_ARCH)}); other backbones (e.g. SwinV2) are out of scope"
        )
    arch = "_".join(parts[:2])
    if arch not in _ARCH:
        raise ValueError(f"unknown ViT size in {name!r}; known: {sorted(_ARCH)}")
    patch, img = int(parts[2][5:]), int(parts[3])
    if patch != 16:
        raise ValueError("only patch16 ViTs are covered")
    if img % 16:
        raise ValueError("image size must be a multiple of 16")
    return (*_ARCH[arch], patch, img)


# ----------------------------------------------------------------------------------------------
# parameter holders with timm's attribute / key names
# ----------------------------------------------------------------------------------------------
class PatchEmbed(nn.Module):
    def __init__(self, img_size: int, patch_size: int, in_chans: int, embed_dim: int) -> None:
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=True)


class Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int) -> None:
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)


class Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int) -> None:
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden, bias=True)
        self.fc2 = nn.Linear(hidden, dim, bias=True)


class Block(nn.Module):
    def __init__(self, dim: int, num_heads: int, mlp_ratio: float, drop_path: float) -> None:
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.drop_path_rate = float(drop_path)


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _grad_buffer(p: nn.Parameter) -> Tensor:
    """Where this parameter's gradient is accumulated. Creates (zeroed) ``p.grad`` on first use;
    inside a FlatArena that is the parameter's slice of the flat gradient buffer."""
    a = getattr(p, "_fv_arena", None)
    if a is not None:
        a[0].grads_clean = False  # something is about to be accumulated into the flat gradient buffer
    if p.grad is None:
        if a is not None and a[0].owns(p):
            g = a[0].grad_view(p)
            g.zero_()
            p.grad = g
        else:
            p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


def _split_k_for(out_rows: int, out_cols: int, k: int, sms: int = 148) -> int:
    """K slices for a weight-gradient GEMM (small M x N, huge K = tokens): the smallest split whose
    item count fills whole waves of the persistent grid best, with at least 16 k-blocks per slice.
    The weight gradient runs on the CTA-pair kernel: 256 x 256 tiles on sms // 2 pairs (fc1 at
    ViT-B: 36 tiles -> 2 slices = 72 items on 74 pairs; qkv: 27 tiles -> 8 slices = 2.92 waves)."""
    tiles = ((out_rows + 255) // 256) * ((out_cols + 255) // 256)
    units = max(1, sms // 2)
    kb = (k + 63) // 64
    best, best_eff = 1, 0.0
    for s in range(1, 17):
        if s > 1 and kb // s < 16:
            break
        t = tiles * s
        eff = t / (((t + units - 1) // units) * units)
        if eff > best_eff + 0.01:
            best, best_eff = s, eff
    return best


def _to_bf16(t: Tensor) -> Tensor:
    t = t.contiguous()
    out = torch.empty(t.shape, device=t.device, dtype=torch.bfloat16)
    ops.cast_bf16(t, out)
    return out


def weight_operand(p: nn.Parameter, lp: bool) -> Tensor:
    """GEMM operand for a weight: the fp32 master (fp32 mode) or its bf16 shadow (the FlatArena's
    bf16 copy, rewritten by the optimiser sweep, when the parameter lives in one)."""
    w = p.detach()
    if w.dim() == 4:
        w = w.reshape(w.shape[0], -1)
    if not lp:
        return w
    a = getattr(p, "_fv_arena", None)
    if a is not None and a[0].lp is not None and a[0].owns(p):
        v = a[0].lp_view(p)
        return v.reshape(v.shape[0], -1) if v.dim() == 4 else v
    return _to_bf16(w)


class _Saved:
    __slots__ = ("lp", "B", "N", "img", "patches", "blocks", "cls_rows", "meanf", "rstdf", "cls_only")


class VisionTransformer(nn.Module):
    def __init__(self, img_size: int = 224, patch_size: int = 16, in_chans: int = 3,
                 num_classes: int = 0, embed_dim: int = 768, depth: int = 12, num_heads: int = 12,
                 mlp_ratio: float = 4.0, drop_path_rate: float = 0.0) -> None:
        super().__init__()
        if embed_dim // num_heads != 64:
            raise ValueError("head_dim must be 64 (true for every timm ViT size)")
        if num_classes != 0:
            raise ValueError("the reference builds the backbone with num_classes=0 (model.py:115)")
        self.num_classes = 0
        self.num_features = self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.num_prefix_tokens = 1
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.num_tokens = self.patch_embed.num_patches + 1
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_tokens, embed_dim))
        dpr = [r.item() for r in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio, dpr[i]) for i in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        # Opt-in (``model.cls_only_last_block``): the head reads the cls token only, and everything after the
        # last block's attention is token-wise, so the other 196 rows of that block's proj / LN2 / MLP
        # (forward) and of their gradients (backward: exactly zero) never reach the loss. With the switch on
        # those kernels run on the B cls rows instead of B*N — same logits, same parameter gradients
        # (up to summation order), ~4 % fewer FLOPs per ViT-B/16 step. Off by default: the dense schedule is
        # what timm executes.
        self.cls_only_last_block = False
        self.init_weights()

    def init_weights(self) -> None:
        # timm's init for pretrained=False: trunc_normal(.02) Linear weights and
        # pos_embed, zero biases, cls_token ~ N(0, 1e-6); the patch conv keeps PyTorch's default.
        nn.init.trunc_normal_(self.pos_embed, std=0.02, a=-2.0, b=2.0)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02, a=-2.0, b=2.0)
                nn.init.zeros_(m.bias)

    # ------------------------------------------------------------------------------------------
    def _drop_path_scale(self, rate: float, batch: int, device) -> Optional[Tensor]:
        """Per-sample stochastic-depth factor mask / keep_prob (timm DropPath, scale_by_keep=True),
        or None when the branch is always kept (eval mode or rate 0)."""
        if not self.training or rate <= 0.0:
            return None
        keep = 1.0 - rate
        mask = torch.empty(batch, device=device, dtype=torch.float32).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return mask

    def _bb_params(self) -> List[nn.Parameter]:
        return list(self.parameters())

    def _w(self, p: nn.Parameter, lp: bool) -> Tensor:
        return weight_operand(p, lp)

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise FedVitError(
                "fedvit_b200 backbone runs on CUDA (sm_100a) only — no CPU/MPS fallback on this path"
            )
        pe = self.patch_embed
        if tuple(x.shape[-2:]) != pe.img_size:
            raise ValueError(f"input {tuple(x.shape[-2:])} != model image size {pe.img_size}")
        if x.shape[1] != pe.proj.weight.shape[1]:
            raise ValueError(f"input has {x.shape[1]} channels, patch_embed.proj expects {pe.proj.weight.shape[1]}")
        lp = torch.is_autocast_enabled("cuda")
        if lp:
            a = getattr(self.cls_token, "_fv_arena", None)
            if a is not None and a[0].lp is not None and a[0].owns(self.cls_token):
                a[0].refresh_lp()
        params = self._bb_params()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if need_grad:
            a = getattr(self.cls_token, "_fv_arena", None)
            if a is not None:  # autograd-side modules (metadata MLP) accumulate into arena views too
                a[0].grads_clean = False
            return _VitFunction.apply(self, x.float(), lp, *params)
        feats, _ = self._forward_impl(x.float(), lp, save=False)
        return feats

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def _attention_fwd(self, qkv: Tensor, B: int, N: int, lp: bool):
        H, scale = self.num_heads, 64 ** -0.5
        if lp:
            o, lse = ops.attention_fwd(qkv, B, N, H, scale)
            return o, lse
        # fp32 parity path: S = Q K^T, P = softmax(scale S), O = P V as strided-batched FFMA GEMMs
        D = H * 64
        P = torch.empty((B, H, N, N), device=qkv.device, dtype=torch.float32)
        o = torch.empty((B * N, D), device=qkv.device, dtype=torch.float32)
        for h in range(H):
            q, k, v = (qkv[:, s * D + h * 64:] for s in range(3))
            ops.bgemm_f32(q, [3 * D, 1, N * 3 * D], k, [3 * D, 1, N * 3 * D], P[:, h], [N, H * N * N],
                          N, N, 64, B, 1.0, False)
        P = ops.softmax_rows(P, scale)
        for h in range(H):
            v = qkv[:, 2 * D + h * 64:]
            ops.bgemm_f32(P[:, h], [N, 1, H * N * N], v, [1, 3 * D, N * 3 * D], o[:, h * 64:], [D, N * D],
                          N, 64, N, B, 1.0, False)
        return o, P

    def _attention_bwd(self, qkv: Tensor, o: Tensor, do: Tensor, aux: Tensor, B: int, N: int, lp: bool) -> Tensor:
        H, scale = self.num_heads, 64 ** -0.5
        if lp:
            return ops.attention_bwd(qkv, o, do, aux, B, N, H, scale)
        D = H * 64
        P = aux
        dqkv = torch.empty_like(qkv)
        dP = torch.empty_like(P)
        for h in range(H):
            v = qkv[:, 2 * D + h * 64:]
            dO = do[:, h * 64:]
            # dV[key,d] = sum_q P[q,key] dO[q,d]
            ops.bgemm_f32(P[:, h], [1, N, H * N * N], dO, [1, D, N * D], dqkv[:, 2 * D + h * 64:],
                          [3 * D, N * 3 * D], N, 64, N, B, 1.0, False)
            # dP[q,key] = sum_d dO[q,d] V[key,d]
            ops.bgemm_f32(dO, [D, 1, N * D], v, [3 * D, 1, N * 3 * D], dP[:, h], [N, H * N * N],
                          N, N, 64, B, 1.0, False)
        dS = ops.softmax_rows_bwd(P, dP, scale)
        for h in range(H):
            q, k = qkv[:, h * 64:], qkv[:, D + h * 64:]
            # dQ[q,d] = sum_key dS[q,key] K[key,d]
            ops.bgemm_f32(dS[:, h], [N, 1, H * N * N], k, [1, 3 * D, N * 3 * D], dqkv[:, h * 64:],
                          [3 * D, N * 3 * D], N, 64, N, B, 1.0, False)
            # dK[key,d] = sum_q dS[q,key] Q[q,d]
            ops.bgemm_f32(dS[:, h], [1, N, H * N * N], q, [1, 3 * D, N * 3 * D], dqkv[:, D + h * 64:],
                          [3 * D, N * 3 * D], N, 64, N, B, 1.0, False)
        return dqkv

    def _forward_impl(self, img: Tensor, lp: bool, save: bool):
        B = img.shape[0]
        N, D = self.num_tokens, self.embed_dim
        M = B * N
        dev = img.device
        act = torch.bfloat16 if lp else torch.float32
        pe = self.patch_embed

        x = torch.empty((M, D), device=dev, dtype=torch.float32)
        patches = None
        if lp:
            # im2col-free: the 16x16xC patches are fetched by a 5-D TMA map straight out of the NCHW
            # image (tf32 tensor cores); nothing [B*196, C*256]-shaped is materialised in the forward
            ops.patch_embed(img, pe.proj.weight.detach(), pe.proj.bias.detach(),
                            self.pos_embed.detach().view(N, D), x)
        else:
            patches = ops.patchify(img, False)  # fp32 parity path: explicit patch rows + FFMA GEMM
            ops.gemm(patches, self._w(pe.proj.weight, lp), pe.proj.bias.detach(), x,
                     self.pos_embed.detach().view(N, D), _K, _K, _E["patch"], 1, N - 1)
        ops.cls_pos_rows(self.cls_token.detach(), self.pos_embed.detach(), x, B, N, D)

        st = None
        if save:
            st = _Saved()
            st.lp, st.B, st.N = lp, B, N
            st.patches = patches
            st.img = img
            st.blocks = []
        nblk = len(self.blocks)
        cls_only = False
        for bi, blk in enumerate(self.blocks):
            n1, n2, at, mlp = blk.norm1, blk.norm2, blk.attn, blk.mlp
            h, mean1, rstd1 = ops.layernorm_fwd(x, n1.weight.detach(), n1.bias.detach(), n1.eps, lp)
            qkv = torch.empty((M, 3 * D), device=dev, dtype=act)
            ops.gemm(h, self._w(at.qkv.weight, lp), at.qkv.bias.detach(), qkv, None, _K, _K, _E["none"], 1, 0)
            o, aux = self._attention_fwd(qkv, B, N, lp)
            s1 = self._drop_path_scale(blk.drop_path_rate, B, dev)
            s2 = self._drop_path_scale(blk.drop_path_rate, B, dev)
            cls_only = self.cls_only_last_block and bi == nblk - 1
            rows, per = (B, 1) if cls_only else (M, N)  # rows the token-wise tail runs on, rows per sample
            if cls_only:
                o_in = o.view(B, N, D)[:, 0].contiguous()
                x_in = x.view(B, N, D)[:, 0].contiguous()
            else:
                o_in, x_in = o, x
            x1 = torch.empty((rows, D), device=dev, dtype=torch.float32)
            ops.linear_residual(o_in, self._w(at.proj.weight, lp), at.proj.bias.detach(), x_in, s1, per, x1)
            h2, mean2, rstd2 = ops.layernorm_fwd(x1, n2.weight.detach(), n2.bias.detach(), n2.eps, lp)
            hid = mlp.fc1.weight.shape[0]
            a = torch.empty((rows, hid), device=dev, dtype=act)
            if save:
                u = torch.empty((rows, hid), device=dev, dtype=act)  # gelu'(fc1 output), all the backward needs
                ops.gemm_gelu(h2, self._w(mlp.fc1.weight, lp), mlp.fc1.bias.detach(), a, u)
            else:  # forward only: no derivative output
                u = None
                ops.gemm_gelu_fwd(h2, self._w(mlp.fc1.weight, lp), mlp.fc1.bias.detach(), a)
            x2 = torch.empty((rows, D), device=dev, dtype=torch.float32)
            ops.linear_residual(a, self._w(mlp.fc2.weight, lp), mlp.fc2.bias.detach(), x1, s2, per, x2)
            if save:
                st.blocks.append((x, h, mean1, rstd1, qkv, o, aux, x1, h2, mean2, rstd2, u, a, s1, s2))
            x = x2

        cls_rows = x if cls_only else x.view(B, N, D)[:, 0].contiguous()
        nf = self.norm
        feats, meanf, rstdf = ops.layernorm_fwd(cls_rows, nf.weight.detach(), nf.bias.detach(), nf.eps, False)
        if save:
            st.cls_rows, st.meanf, st.rstdf = cls_rows, meanf, rstdf
            st.cls_only = cls_only
        return feats, st

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def _linear_bwd(self, dy: Tensor, x_in: Tensor, lin: nn.Linear, lp: bool, need_dx: bool,
                    dgelu_aux: Optional[Tensor] = None) -> Optional[Tensor]:
        """Gradients of y = x W^T + b given dy [M, out]: accumulates dW, db; returns dx (or None)."""
        M = dy.shape[0]
        out_f, in_f = lin.weight.shape
        want_b = lin.bias is not None and lin.bias.requires_grad
        if lp and lin.weight.requires_grad:
            # bias gradient rides along in the weight-gradient kernel (no extra pass over dy)
            ops.wgrad(dy, x_in, _grad_buffer(lin.weight), _grad_buffer(lin.bias) if want_b else None,
                      _split_k_for(out_f, in_f, M))
        else:
            if lin.weight.requires_grad:
                ops.gemm(dy, x_in, None, _grad_buffer(lin.weight), None, _MN, _MN, _E["accum"], 1, 0)
            if want_b:
                ops.colsum(dy, _grad_buffer(lin.bias), True)
        if not need_dx:
            return None
        dx = torch.empty((M, in_f), device=dy.device, dtype=dy.dtype)
        if dgelu_aux is not None:
            ops.gemm(dy, self._w(lin.weight, lp), None, dx, dgelu_aux, _K, _MN, _E["dgelu"], 1, 0)
        else:
            ops.gemm(dy, self._w(lin.weight, lp), None, dx, None, _K, _MN, _E["none"], 1, 0)
        return dx

    def _ln_bwd(self, dy: Tensor, x: Tensor, ln: nn.LayerNorm, mean: Tensor, rstd: Tensor,
                dres: Optional[Tensor], lp: bool, branch_scale: Optional[Tensor] = None,
                rows_per_sample: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        """Returns (dx fp32, the gradient the preceding sub-layer's branch GEMMs consume): dx in the
        GEMM operand type, times that branch's stochastic-depth factor when it has one."""
        if ln.weight.requires_grad:
            dg, db = _grad_buffer(ln.weight), _grad_buffer(ln.bias)
        else:
            dg = torch.zeros_like(ln.weight)
            db = torch.zeros_like(ln.bias)
        rows_per_sample = self.num_tokens if rows_per_sample is None else rows_per_sample
        rows_per = rows_per_sample if branch_scale is not None else 0
        dx, dx_lp = ops.layernorm_bwd(dy, x, ln.weight.detach(), mean, rstd, dres, dg, db, lp,
                                      branch_scale if lp else None, rows_per)
        if lp:
            return dx, dx_lp
        if branch_scale is not None:  # fp32 parity path: explicit scaled copy
            return dx, (dx.view(-1, rows_per_sample, dx.shape[1]) * branch_scale.view(-1, 1, 1)).view_as(dx)
        return dx, dx

    def _backward_impl(self, st: _Saved, dfeats: Tensor) -> None:
        lp, B, N = st.lp, st.B, st.N
        D = self.embed_dim
        M = B * N
        dev = dfeats.device
        dcls, _ = self._ln_bwd(dfeats.float().contiguous(), st.cls_rows, self.norm, st.meanf, st.rstdf, None, False)
        last_s2 = st.blocks[-1][-1]
        if st.cls_only:
            # the last block's token-wise tail ran on the cls rows only: so does its backward
            dx = dcls
            dyf = dcls * last_s2.view(B, 1) if last_s2 is not None else dcls
            dy = _to_bf16(dyf) if lp else dyf
        else:
            # one pass writes both the fp32 residual-path gradient and the operand of the last block's
            # fc2 gradients (bf16 under autocast; times that block's stochastic-depth factor): zero
            # everywhere except the cls rows
            dx = torch.empty((M, D), device=dev, dtype=torch.float32)
            if lp or last_s2 is not None:
                dy = torch.empty((M, D), device=dev, dtype=torch.bfloat16 if lp else torch.float32)
                ops.cls_grad_rows(dcls, last_s2, dx, dy, B, N, D)
            else:
                ops.cls_grad_rows(dcls, None, dx, None, B, N, D)
                dy = dx

        nblk = len(st.blocks)
        for i in range(nblk - 1, -1, -1):
            blk = self.blocks[i]
            (x, h, mean1, rstd1, qkv, o, aux, x1, h2, mean2, rstd2, u, a, s1, s2) = st.blocks[i]
            prev_s2 = st.blocks[i - 1][-1] if i > 0 else None
            cls_only = st.cls_only and i == nblk - 1
            # MLP branch: x2 = x1 + s2 * fc2(gelu(fc1(LN2(x1))))   (dy already carries s2)
            du = self._linear_bwd(dy, a, blk.mlp.fc2, lp, True, dgelu_aux=u)
            dh2 = self._linear_bwd(du, h2, blk.mlp.fc1, lp, True)
            dx1, dy1 = self._ln_bwd(dh2, x1, blk.norm2, mean2, rstd2, dx, lp, branch_scale=s1,
                                    rows_per_sample=1 if cls_only else None)
            # attention branch: x1 = x + s1 * proj(attn(qkv(LN1(x))))
            o_in = o.view(B, N, D)[:, 0].contiguous() if cls_only else o
            do = self._linear_bwd(dy1, o_in, blk.attn.proj, lp, True)
            if cls_only:  # back to all tokens: gradients of the other rows are exactly zero up to here
                do_c, dx1_c = do, dx1
                do = torch.zeros((M, D), device=dev, dtype=do_c.dtype)
                do.view(B, N, D)[:, 0] = do_c
                dx1 = torch.zeros((M, D), device=dev, dtype=torch.float32)
                dx1.view(B, N, D)[:, 0] = dx1_c
            dqkv = self._attention_bwd(qkv, o, do, aux, B, N, lp)
            dh = self._linear_bwd(dqkv, h, blk.attn.qkv, lp, True)
            dx, dy = self._ln_bwd(dh, x, blk.norm1, mean1, rstd1, dx1, lp, branch_scale=prev_s2)

        # embedding: x0[b] = cat(cls, patches[b] W^T + b) + pos
        proj = self.patch_embed.proj
        if lp and st.patches is None and N == self.patch_embed.num_patches + 1:
            # bf16 path after the im2col-free forward. The stream's gradient is contracted as it lies in memory:
            # the patch rows are rebuilt with a zero row in every image's cls slot (fv_patchify_rows), so the
            # weight gradient needs no copy of dy's patch rows (a strided 77 MB bf16 copy, 85 us per step), and
            # ONE column sum over the batch, [N, D], carries the three small gradients: pos_embed (all of it),
            # cls_token (row 0), the conv bias (rows 1 .. N - 1 summed) — from the fp32 gradient.
            if self.pos_embed.requires_grad or self.cls_token.requires_grad or proj.bias.requires_grad:
                tok = torch.empty(N * D, device=dev, dtype=torch.float32)
                ops.colsum(dx.view(B, N * D), tok, False)
                if self.pos_embed.requires_grad:
                    _grad_buffer(self.pos_embed).view(N * D).add_(tok)
                if self.cls_token.requires_grad:
                    _grad_buffer(self.cls_token).view(D).add_(tok[:D])
                if proj.bias.requires_grad:
                    _grad_buffer(proj.bias).add_(tok.view(N, D)[1:].sum(0))
            if proj.weight.requires_grad:
                gw = _grad_buffer(proj.weight)
                ops.wgrad(dy, ops.patchify(st.img, True, 1), gw.view(D, -1), None, _split_k_for(D, gw.numel() // D, M))
            return
        if self.pos_embed.requires_grad:
            ops.colsum(dx.view(B, N * D), _grad_buffer(self.pos_embed).view(N * D), True)
        if self.cls_token.requires_grad:
            ops.colsum(dx.view(B, N * D)[:, :D], _grad_buffer(self.cls_token).view(D), True)
        if proj.weight.requires_grad or proj.bias.requires_grad:
            dpatch = dy.view(B, N, D)[:, 1:].reshape(B * (N - 1), D)
            if st.patches is None:  # the forward was im2col-free; the weight gradient wants patch rows
                st.patches = ops.patchify(st.img, lp)
            if lp and proj.weight.requires_grad:
                gw = _grad_buffer(proj.weight)
                ops.wgrad(dpatch, st.patches, gw.view(D, -1),
                          _grad_buffer(proj.bias) if proj.bias.requires_grad else None,
                          _split_k_for(D, gw.numel() // D, B * (N - 1)))
            else:
                if proj.weight.requires_grad:
                    gw = _grad_buffer(proj.weight)
                    ops.gemm(dpatch, st.patches, None, gw.view(D, -1), None, _MN, _MN, _E["accum"], 1, 0)
                if proj.bias.requires_grad:
                    ops.colsum(dpatch, _grad_buffer(proj.bias), True)


class _VitFunction(torch.autograd.Function):
    """Whole-backbone autograd node. Parameter gradients are accumulated by the kernels straight
    into ``p.grad`` (the FlatArena gradient buffer when there is one), so ``backward`` returns
    ``None`` for them: no per-tensor AccumulateGrad copies, and gradient accumulation across
    micro-steps (reference train.py:151-155) is the kernels' own ``+=``."""

    @staticmethod
    def forward(ctx, vit: VisionTransformer, img: Tensor, lp: bool, *params):
        feats, st = vit._forward_impl(img, lp, save=True)
        ctx.vit = vit
        ctx.st = st
        ctx.nparams = len(params)
        return feats

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dfeats: Tensor):
        st, ctx.st = ctx.st, None
        if st is None:
            raise RuntimeError("fedvit backbone: backward called twice on the same graph")
        ctx.vit._backward_impl(st, dfeats)
        return (None, None, None) + (None,) * ctx.nparams


def create_model(model_name: str, pretrained: bool = False, num_classes: int = 0,
                 drop_path_rate: float = 0.0, in_chans: int = 3, **kwargs) -> VisionTransformer:
    """Signature of ``timm.create_model`` as the reference calls it (model.py:112-117)."""
    if kwargs:
        raise TypeError(f"unsupported create_model arguments: {sorted(kwargs)}")
    if pretrained:
        raise RuntimeError(
            "pretrained=True needs timm's weight download, which this environment cannot do; "
            "set model.pretrained: false and load a timm ViT state_dict instead (keys are identical)"
        )
    d, l, h, patch, img = parse_vit_name(model_name)
    return VisionTransformer(img_size=img, patch_size=patch, in_chans=in_chans, num_classes=num_classes,
                             embed_dim=d, depth=l, num_heads=h, drop_path_rate=drop_path_rate)
