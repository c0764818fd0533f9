"""``torch.library`` custom ops (namespace ``fedvit``) over the libfedvit C ABI.

Each op is a thin shim: check shapes/dtypes, allocate outputs with ``torch.empty`` (so PyTorch's
caching allocator and stream semantics own every buffer), pass raw device pointers plus the
current CUDA stream through ctypes into ``include/fedvit.h``. There is no eager/CPU fallback: a
CPU tensor or a missing library raises.

The ops are forward primitives *and* backward primitives (``layernorm_bwd``, ``attention_bwd``,
the dgrad / wgrad GEMM modes). Autograd is wired one level up, in ``vit.py`` and ``losses.py``,
where the saved-activation lifetime is managed explicitly.
"""
from __future__ import annotations

import functools
import os
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from ._lib import LIB, FedVitError

F32, BF16 = 0, 1
EPI = {"none": 0, "residual": 1, "gelu": 2, "dgelu": 3, "accum": 4, "patch": 5}
MAJOR_K, MAJOR_MN = 0, 1


class _Op:
    """Python entry point of one library op. Calling it runs the implementation directly (argument
    checks + one ctypes call on the current stream). The same function is registered as the
    ``torch.library`` custom op ``torch.ops.fedvit.<name>`` (``.op``): that is the path for FakeTensor
    tracing / export, but each trip through the dispatcher costs ~20 us of host time in Python — 5 ms of a
    250-launch training step, more than the GPU needs for a ViT-Tiny step — so the framework's own callers
    (vit.py, head.py, losses.py, optim.py) take the direct route. ``FEDVIT_DISPATCH=1`` sends every call
    through the dispatcher instead (A/B)."""

    def __init__(self, name: str, fn, mutates_args) -> None:
        self.fn = fn
        self.op = torch.library.custom_op(name, mutates_args=mutates_args)(fn)
        self._via_dispatcher = os.environ.get("FEDVIT_DISPATCH", "0") == "1"
        functools.update_wrapper(self, fn)

    def __call__(self, *args, **kwargs):
        if self._via_dispatcher:
            return self.op(*args, **kwargs)
        return self.fn(*args, **kwargs)

    def register_fake(self, fake):
        return self.op.register_fake(fake)


def _op(name: str, mutates_args):
    return lambda fn: _Op(name, fn, mutates_args)


def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise FedVitError(f"unsupported dtype {t.dtype}")


def _stream(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise FedVitError(
                "fedvit ops run on CUDA (sm_100a) only — there is no CPU fallback on this path"
            )


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
# bench.py's roofline leg: when set to a list, every tensor-core GEMM launch is bracketed by CUDA
# events on its own stream and (start, end, algorithmic flops) is appended here.
GEMM_TRACE: Optional[list] = None


def _gemm_impl(a, b, bias, out, aux, a_major, b_major, epilogue, split_k, tokens_per_img) -> None:
    _need_cuda(a, b, bias, out, aux)
    if a.dim() != 2 or b.dim() != 2 or out.dim() != 2:
        raise FedVitError("gemm: operands must be 2-D")
    if a.stride(1) != 1 or b.stride(1) != 1 or out.stride(1) != 1:
        raise FedVitError("gemm: innermost stride must be 1")
    m, k = (a.shape[0], a.shape[1]) if a_major == MAJOR_K else (a.shape[1], a.shape[0])
    n, kb = (b.shape[0], b.shape[1]) if b_major == MAJOR_K else (b.shape[1], b.shape[0])
    if k != kb:
        raise FedVitError(f"gemm: contraction mismatch {k} vs {kb}")
    if epilogue == EPI["patch"]:
        if out.shape[1] != n:
            raise FedVitError("gemm: patch epilogue output width mismatch")
    elif tuple(out.shape) != (m, n):
        raise FedVitError(f"gemm: out is {tuple(out.shape)}, expected {(m, n)}")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != n):
        raise FedVitError("gemm: bias must be fp32 [N]")
    ldaux = aux.stride(0) if aux is not None else 0
    if a.dtype == torch.bfloat16:
        if b.dtype != torch.bfloat16:
            raise FedVitError("gemm: mixed operand dtypes")
        trace = GEMM_TRACE
        if trace is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        LIB.call(
            "fv_gemm_bf16", a.data_ptr(), a_major, a.stride(0), b.data_ptr(), b_major, b.stride(0),
            _ptr(bias), out.data_ptr(), _dt(out), out.stride(0), _ptr(aux), ldaux, m, n, k,
            epilogue, split_k, tokens_per_img, _stream(a),
        )
        if trace is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            trace.append((e0, e1, 2.0 * m * n * k))
    else:
        if a.dtype != torch.float32 or b.dtype != torch.float32 or out.dtype != torch.float32:
            raise FedVitError("gemm: fp32 path needs fp32 a, b, out")
        ars, acs = (a.stride(0), 1) if a_major == MAJOR_K else (1, a.stride(0))
        brs, bcs = (b.stride(0), 1) if b_major == MAJOR_K else (1, b.stride(0))
        LIB.call(
            "fv_gemm_f32", a.data_ptr(), ars, acs, 0, b.data_ptr(), brs, bcs, 0, _ptr(bias),
            out.data_ptr(), out.stride(0), 0, _ptr(aux), ldaux, m, n, k, 1, 1.0, epilogue,
            tokens_per_img, _stream(a),
        )


@_op("fedvit::gemm", mutates_args=("out",))
def gemm(
    a: Tensor,
    b: Tensor,
    bias: Optional[Tensor],
    out: Tensor,
    aux: Optional[Tensor],
    a_major: int,
    b_major: int,
    epilogue: int,
    split_k: int,
    tokens_per_img: int,
) -> None:
    """out[M,N] = op(a)[M,K] @ op(b)[N,K]^T with a fused epilogue (see include/fedvit.h).

    ``a_major`` / ``b_major``: 0 = the operand is stored [rows, K]; 1 = stored [K, rows].
    ``aux`` is read-only here (residual, saved pre-activation, pos_embed); the GELU epilogue, which
    writes a second output, is ``gemm_gelu``. bf16 operands run on tcgen05 tensor cores, fp32
    operands on the FFMA parity kernel.
    """
    if epilogue == EPI["gelu"]:
        raise FedVitError("gemm: use gemm_gelu for the GELU epilogue")
    _gemm_impl(a, b, bias, out, aux, a_major, b_major, epilogue, split_k, tokens_per_img)


@_op("fedvit::gemm_gelu", mutates_args=("out", "pre"))
def gemm_gelu(a: Tensor, b: Tensor, bias: Optional[Tensor], out: Tensor, pre: Tensor) -> None:
    """u = a @ b^T + bias ; out = gelu_erf(u) ; pre = gelu_erf'(u)  (fc1 of the MLP).

    The second output is the GELU derivative at the pre-activation, not the pre-activation itself:
    it shares the cdf/pdf evaluation with the activation, and it is all the backward needs — the
    fc2 dgrad epilogue (``EPI["dgelu"]``) just multiplies by it."""
    _gemm_impl(a, b, bias, out, pre, MAJOR_K, MAJOR_K, EPI["gelu"], 1, 0)


@_op("fedvit::gemm_gelu_fwd", mutates_args=("out",))
def gemm_gelu_fwd(a: Tensor, b: Tensor, bias: Optional[Tensor], out: Tensor) -> None:
    """out = gelu_erf(a @ b^T + bias) without the derivative output — the forward-only (eval /
    no-grad) form of :func:`gemm_gelu`: half the epilogue's stores."""
    _gemm_impl(a, b, bias, out, None, MAJOR_K, MAJOR_K, EPI["gelu"], 1, 0)


@_op("fedvit::linear_residual", mutates_args=("out",))
def linear_residual(a: Tensor, w: Tensor, bias: Optional[Tensor], residual: Tensor, row_scale: Optional[Tensor],
                    rows_per_scale: int, out: Tensor) -> None:
    """out = residual + row_scale[row // rows_per_scale] * (a @ w^T + bias) — a block's branch +
    residual with per-sample stochastic depth (``row_scale=None``: plain residual add)."""
    _need_cuda(a, w, bias, residual, row_scale, out)
    m, k = a.shape
    n = w.shape[0]
    if w.shape[1] != k or tuple(out.shape) != (m, n) or tuple(residual.shape) != (m, n):
        raise FedVitError("linear_residual: shape mismatch")
    if row_scale is not None and (row_scale.dtype != torch.float32 or rows_per_scale <= 0
                                  or row_scale.numel() * rows_per_scale < m):
        raise FedVitError("linear_residual: row_scale must be fp32 with one entry per rows_per_scale rows")
    name = "fv_linear_residual_bf16" if a.dtype == torch.bfloat16 else "fv_linear_residual_f32"
    if a.dtype != w.dtype or a.dtype not in (torch.bfloat16, torch.float32):
        raise FedVitError("linear_residual: operands must both be bf16 or both fp32")
    trace = GEMM_TRACE if a.dtype == torch.bfloat16 else None
    if trace is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
    LIB.call(name, a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), residual.data_ptr(),
             residual.stride(0), _ptr(row_scale), rows_per_scale, out.data_ptr(), out.stride(0), m, n, k,
             _stream(a))
    if trace is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        trace.append((e0, e1, 2.0 * m * n * k))


@_op("fedvit::wgrad", mutates_args=("dw", "dbias"))
def wgrad(dy: Tensor, x: Tensor, dw: Tensor, dbias: Optional[Tensor], split_k: int) -> None:
    """dw[out,in] += dy^T @ x and dbias[out] += dy.sum(0), bf16 operands, one tensor-core kernel."""
    _need_cuda(dy, x, dw, dbias)
    if dy.dtype != torch.bfloat16 or x.dtype != torch.bfloat16 or dw.dtype != torch.float32:
        raise FedVitError("wgrad: bf16 dy/x and fp32 dw required")
    if dy.dim() != 2 or x.dim() != 2 or dy.shape[0] != x.shape[0] or dy.stride(1) != 1 or x.stride(1) != 1:
        raise FedVitError("wgrad: dy [tokens,out], x [tokens,in] row-major")
    if tuple(dw.shape) != (dy.shape[1], x.shape[1]) or dw.stride(1) != 1:
        raise FedVitError("wgrad: dw must be [out, in]")
    if dbias is not None and (dbias.dtype != torch.float32 or dbias.numel() != dy.shape[1] or not dbias.is_contiguous()):
        raise FedVitError("wgrad: dbias must be contiguous fp32 [out]")
    trace = GEMM_TRACE
    if trace is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
    LIB.call("fv_wgrad_bf16", dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dw.data_ptr(),
             dw.stride(0), _ptr(dbias), dy.shape[0], dy.shape[1], x.shape[1], split_k, _stream(dy))
    if trace is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        trace.append((e0, e1, 2.0 * dy.shape[0] * dy.shape[1] * x.shape[1]))


@_op("fedvit::bgemm_f32", mutates_args=("out",))
def bgemm_f32(
    a: Tensor, a_strides: List[int], b: Tensor, b_strides: List[int], out: Tensor,
    out_strides: List[int], m: int, n: int, k: int, batch: int, alpha: float, accumulate: bool,
) -> None:
    """Strided-batched fp32 product for the parity attention path.

    out[z](i,j) (+)= alpha * sum_t a[z](i,t) * b[z](j,t); strides are (row, col, batch) for a/b and
    (row, batch) for out, all in elements from the tensors' data pointers.
    """
    _need_cuda(a, b, out)
    LIB.call(
        "fv_gemm_f32", a.data_ptr(), a_strides[0], a_strides[1], a_strides[2], b.data_ptr(),
        b_strides[0], b_strides[1], b_strides[2], None, out.data_ptr(), out_strides[0],
        out_strides[1], None, 0, m, n, k, batch, alpha, EPI["accum"] if accumulate else EPI["none"],
        0, _stream(a),
    )


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
@_op("fedvit::layernorm_fwd", mutates_args=())
def layernorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, out_bf16: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda(x, gamma, beta)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 2:
        raise FedVitError("layernorm_fwd: x must be a contiguous fp32 [rows, cols] tensor")
    rows, cols = x.shape
    y = torch.empty((rows, cols), device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    mean = torch.empty((rows,), device=x.device, dtype=torch.float32)
    rstd = torch.empty((rows,), device=x.device, dtype=torch.float32)
    LIB.call("fv_layernorm_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
             _dt(y), mean.data_ptr(), rstd.data_ptr(), rows, cols, eps, _stream(x))
    return y, mean, rstd


@layernorm_fwd.register_fake
def _(x, gamma, beta, eps, out_bf16):
    y = x.new_empty(x.shape, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    return y, x.new_empty((x.shape[0],)), x.new_empty((x.shape[0],))


@_op("fedvit::layernorm_bwd", mutates_args=("dgamma", "dbeta"))
def layernorm_bwd(dy: Tensor, x: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor,
                  dres: Optional[Tensor], dgamma: Tensor, dbeta: Tensor, want_lp: bool,
                  lp_scale: Optional[Tensor] = None, rows_per_scale: int = 0) -> Tuple[Tensor, Tensor]:
    """dx = dres + LN'(dy); dgamma/dbeta are accumulated in place. Returns (dx fp32, dx bf16|empty).
    ``lp_scale`` (fp32, one value per ``rows_per_scale`` rows) multiplies the bf16 copy only
    (stochastic depth of the sub-layer that consumes it)."""
    _need_cuda(dy, x, gamma, mean, rstd, dres, dgamma, dbeta, lp_scale)
    rows, cols = x.shape
    dx = torch.empty_like(x)
    dx_lp = torch.empty((rows, cols) if want_lp else (0,), device=x.device, dtype=torch.bfloat16)
    LIB.call("fv_layernorm_bwd", dy.data_ptr(), _dt(dy), x.data_ptr(), gamma.data_ptr(),
             mean.data_ptr(), rstd.data_ptr(), _ptr(dres), dx.data_ptr(),
             dx_lp.data_ptr() if want_lp else None, dgamma.data_ptr(), dbeta.data_ptr(), rows, cols,
             _ptr(lp_scale) if want_lp else None, rows_per_scale, _stream(x))
    return dx, dx_lp


@layernorm_bwd.register_fake
def _(dy, x, gamma, mean, rstd, dres, dgamma, dbeta, want_lp, lp_scale=None, rows_per_scale=0):
    return torch.empty_like(x), x.new_empty(x.shape if want_lp else (0,), dtype=torch.bfloat16)


# ------------------------------------------------------------------------------------------------
# attention core (bf16 flash kernels)
# ------------------------------------------------------------------------------------------------
@_op("fedvit::attention_fwd", mutates_args=())
def attention_fwd(qkv: Tensor, batch: int, tokens: int, heads: int, scale: float) -> Tuple[Tensor, Tensor]:
    """qkv [B*N, 3*H*64] bf16 -> (out [B*N, H*64] bf16, lse [B, H, N] fp32)."""
    _need_cuda(qkv)
    if qkv.dtype != torch.bfloat16 or not qkv.is_contiguous() or qkv.shape != (batch * tokens, 3 * heads * 64):
        raise FedVitError("attention_fwd: qkv must be contiguous bf16 [B*N, 3*H*64]")
    out = torch.empty((batch * tokens, heads * 64), device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty((batch, heads, tokens), device=qkv.device, dtype=torch.float32)
    LIB.call("fv_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), BF16, batch, tokens,
             heads, scale, _stream(qkv))
    return out, lse


@attention_fwd.register_fake
def _(qkv, batch, tokens, heads, scale):
    return (qkv.new_empty((batch * tokens, heads * 64)),
            qkv.new_empty((batch, heads, tokens), dtype=torch.float32))


@_op("fedvit::attention_bwd", mutates_args=())
def attention_bwd(qkv: Tensor, out: Tensor, dout: Tensor, lse: Tensor, batch: int, tokens: int,
                  heads: int, scale: float) -> Tensor:
    _need_cuda(qkv, out, dout, lse)
    if dout.dtype != torch.bfloat16 or not dout.is_contiguous():
        raise FedVitError("attention_bwd: dout must be contiguous bf16")
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    ws_bytes = int(LIB.raw("fv_attention_bwd_workspace")(batch, tokens, heads))
    ws = torch.empty(ws_bytes, device=qkv.device, dtype=torch.uint8) if ws_bytes else None
    LIB.call("fv_attention_bwd", qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
             delta.data_ptr(), dqkv.data_ptr(), BF16, batch, tokens, heads, scale, _ptr(ws), ws_bytes, _stream(qkv))
    return dqkv


@attention_bwd.register_fake
def _(qkv, out, dout, lse, batch, tokens, heads, scale):
    return torch.empty_like(qkv)


@_op("fedvit::softmax_rows", mutates_args=())
def softmax_rows(s: Tensor, scale: float) -> Tensor:
    _need_cuda(s)
    p = torch.empty_like(s)
    cols = s.shape[-1]
    LIB.call("fv_softmax_rows", s.data_ptr(), p.data_ptr(), s.numel() // cols, cols, scale, _stream(s))
    return p


@softmax_rows.register_fake
def _(s, scale):
    return torch.empty_like(s)


@_op("fedvit::softmax_rows_bwd", mutates_args=())
def softmax_rows_bwd(p: Tensor, dp: Tensor, scale: float) -> Tensor:
    _need_cuda(p, dp)
    ds = torch.empty_like(p)
    cols = p.shape[-1]
    LIB.call("fv_softmax_rows_bwd", p.data_ptr(), dp.data_ptr(), ds.data_ptr(), p.numel() // cols,
             cols, scale, _stream(p))
    return ds


@softmax_rows_bwd.register_fake
def _(p, dp, scale):
    return torch.empty_like(p)


# ------------------------------------------------------------------------------------------------
# patch embedding helpers, column sums
# ------------------------------------------------------------------------------------------------
@_op("fedvit::patchify", mutates_args=())
def patchify(img: Tensor, out_bf16: bool, lead_rows: int = 0) -> Tensor:
    """NCHW fp32 image -> [B*(lead_rows + (H/16)*(W/16)), C*256] patch rows, (c, py, px) column order, with
    ``lead_rows`` zero rows in front of every image's patches (1 = a cls slot: the rows line up with the
    token rows of the residual stream)."""
    _need_cuda(img)
    if img.dtype != torch.float32 or img.dim() != 4:
        raise FedVitError("patchify: image must be fp32 NCHW")
    img = img.contiguous()
    b, c, h, w = img.shape
    out = torch.empty((b * (lead_rows + (h // 16) * (w // 16)), c * 256), device=img.device,
                      dtype=torch.bfloat16 if out_bf16 else torch.float32)
    LIB.call("fv_patchify_rows", img.data_ptr(), out.data_ptr(), _dt(out), b, c, h, w, lead_rows, _stream(img))
    return out


@patchify.register_fake
def _(img, out_bf16, lead_rows=0):
    b, c, h, w = img.shape
    return img.new_empty((b * (lead_rows + (h // 16) * (w // 16)), c * 256),
                         dtype=torch.bfloat16 if out_bf16 else torch.float32)


@_op("fedvit::patch_embed", mutates_args=("x",))
def patch_embed(img: Tensor, weight: Tensor, bias: Tensor, pos: Tensor, x: Tensor) -> None:
    """Rows 1..N-1 of the residual stream from the NCHW fp32 image, im2col-free (5-D TMA, tf32)."""
    _need_cuda(img, weight, bias, pos, x)
    if img.dtype != torch.float32 or weight.dtype != torch.float32 or x.dtype != torch.float32:
        raise FedVitError("patch_embed: fp32 image / weight / output")
    img = img.contiguous()
    b, c, h, w = img.shape
    d = weight.shape[0]
    if weight.numel() != d * c * 256 or not weight.is_contiguous():
        raise FedVitError("patch_embed: weight must be a contiguous [D, C, 16, 16] conv weight")
    LIB.call("fv_patch_embed_tf32", img.data_ptr(), weight.data_ptr(), bias.data_ptr(), pos.data_ptr(),
             x.data_ptr(), b, c, h, w, d, _stream(img))


@_op("fedvit::cls_pos_rows", mutates_args=("x",))
def cls_pos_rows(cls: Tensor, pos: Tensor, x: Tensor, batch: int, tokens: int, dim: int) -> None:
    _need_cuda(cls, pos, x)
    LIB.call("fv_cls_pos_rows", cls.data_ptr(), pos.data_ptr(), x.data_ptr(), batch, tokens, dim, _stream(x))


@_op("fedvit::cls_grad_rows", mutates_args=("dx", "dy"))
def cls_grad_rows(dcls: Tensor, row_scale: Optional[Tensor], dx: Optional[Tensor], dy: Optional[Tensor],
                  batch: int, tokens: int, dim: int) -> None:
    """Head of the backbone backward: ``dx`` / ``dy`` [B*N, D] are zero except row 0 of each image
    (= dcls, ``dy`` times the per-sample ``row_scale``) — one pass instead of zeros + assign + cast."""
    _need_cuda(dcls, row_scale, dx, dy)
    if dcls.dtype != torch.float32 or not dcls.is_contiguous() or tuple(dcls.shape) != (batch, dim):
        raise FedVitError("cls_grad_rows: dcls must be contiguous fp32 [B, D]")
    for t in (dx, dy):
        if t is not None and (not t.is_contiguous() or t.numel() != batch * tokens * dim):
            raise FedVitError("cls_grad_rows: outputs must be contiguous [B*N, D]")
    if dx is not None and dx.dtype != torch.float32:
        raise FedVitError("cls_grad_rows: dx is fp32")
    LIB.call("fv_cls_grad_rows", dcls.data_ptr(), _ptr(row_scale), _ptr(dx), _ptr(dy),
             _dt(dy) if dy is not None else F32, batch, tokens, dim, _stream(dcls))


@_op("fedvit::colsum", mutates_args=("out",))
def colsum(a: Tensor, out: Tensor, accumulate: bool) -> None:
    """out[c] (+)= sum_r a[r, c] — bias / pos_embed / cls_token gradients."""
    _need_cuda(a, out)
    if a.dim() != 2 or a.stride(1) != 1 or out.dtype != torch.float32:
        raise FedVitError("colsum: a must be 2-D row-major, out fp32")
    LIB.call("fv_colsum", a.data_ptr(), _dt(a), a.stride(0), out.data_ptr(), int(accumulate),
             a.shape[0], a.shape[1], _stream(a))


# ------------------------------------------------------------------------------------------------
# metadata branch stage: Linear + BatchNorm1d + GELU (+ dropout mask), scope row f2
# ------------------------------------------------------------------------------------------------
@_op("fedvit::linear_bn_gelu_fwd", mutates_args=("running_mean", "running_var", "y"))
def linear_bn_gelu_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor], gamma: Tensor, beta: Tensor,
                       running_mean: Tensor, running_var: Tensor, momentum: float, eps: float, training: bool,
                       drop_mask: Optional[Tensor], y: Tensor, save: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """y = drop_mask * gelu(batchnorm(x w^T + bias)) — one stage of reference model.py:27-60. ``y`` may be
    a column slice of a wider row-major buffer. Returns (xhat, dact, rstd) for the backward (``dact``
    empty unless ``save``); the running statistics are updated in place when ``training``."""
    _need_cuda(x, w, bias, gamma, beta, running_mean, running_var, drop_mask, y)
    if x.dim() != 2 or w.dim() != 2 or x.shape[1] != w.shape[1] or x.stride(1) != 1 or not w.is_contiguous():
        raise FedVitError("linear_bn_gelu_fwd: x [B, in] row-major, w [out, in] contiguous")
    b, k = x.shape
    f = w.shape[0]
    for t in (x, w, gamma, beta, running_mean, running_var, y):
        if t.dtype != torch.float32:
            raise FedVitError("linear_bn_gelu_fwd: fp32 tensors required")
    if tuple(y.shape) != (b, f) or y.stride(1) != 1:
        raise FedVitError("linear_bn_gelu_fwd: y must be [B, out] with unit inner stride")
    if drop_mask is not None and (tuple(drop_mask.shape) != (b, f) or not drop_mask.is_contiguous()
                                  or drop_mask.dtype != torch.float32):
        raise FedVitError("linear_bn_gelu_fwd: drop_mask must be contiguous fp32 [B, out]")
    if training and b < 2:
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(x.shape)}")
    xhat = torch.empty((b, f), device=x.device, dtype=torch.float32)
    dact = torch.empty((b, f) if save else (0,), device=x.device, dtype=torch.float32)
    rstd = torch.empty((f,), device=x.device, dtype=torch.float32)
    LIB.call("fv_linear_bn_gelu_fwd", x.data_ptr(), x.stride(0), w.data_ptr(), _ptr(bias), gamma.data_ptr(),
             beta.data_ptr(), running_mean.data_ptr(), running_var.data_ptr(), momentum, eps, int(training),
             _ptr(drop_mask), y.data_ptr(), y.stride(0), xhat.data_ptr(), dact.data_ptr() if save else None,
             rstd.data_ptr(), b, k, f, _stream(x))
    return xhat, dact, rstd


@linear_bn_gelu_fwd.register_fake
def _(x, w, bias, gamma, beta, running_mean, running_var, momentum, eps, training, drop_mask, y, save):
    b, f = x.shape[0], w.shape[0]
    return x.new_empty((b, f)), x.new_empty((b, f) if save else (0,)), x.new_empty((f,))


@_op("fedvit::linear_bn_gelu_bwd", mutates_args=("dw", "dbias", "dgamma", "dbeta"))
def linear_bn_gelu_bwd(dy: Tensor, x: Tensor, xhat: Tensor, dact: Tensor, rstd: Tensor, gamma: Tensor,
                       training: bool, dw: Optional[Tensor], dbias: Optional[Tensor], dgamma: Optional[Tensor],
                       dbeta: Optional[Tensor]) -> Tensor:
    """Backward of ``linear_bn_gelu_fwd``: returns dh [B, out] (gradient w.r.t. the Linear output) and
    accumulates dw / dbias / dgamma / dbeta in place."""
    _need_cuda(dy, x, xhat, dact, rstd, gamma, dw, dbias, dgamma, dbeta)
    b, f = xhat.shape
    k = x.shape[1]
    if dy.dtype != torch.float32 or tuple(dy.shape) != (b, f) or dy.stride(1) != 1 or x.stride(1) != 1:
        raise FedVitError("linear_bn_gelu_bwd: dy [B, out] fp32 with unit inner stride")
    if dw is not None and (tuple(dw.shape) != (f, k) or not dw.is_contiguous()):
        raise FedVitError("linear_bn_gelu_bwd: dw must be contiguous [out, in]")
    dh = torch.empty((b, f), device=dy.device, dtype=torch.float32)
    LIB.call("fv_linear_bn_gelu_bwd", dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), xhat.data_ptr(),
             dact.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), int(training), dh.data_ptr(), _ptr(dw), _ptr(dbias),
             _ptr(dgamma), _ptr(dbeta), b, k, f, _stream(dy))
    return dh


@linear_bn_gelu_bwd.register_fake
def _(dy, x, xhat, dact, rstd, gamma, training, dw, dbias, dgamma, dbeta):
    return dy.new_empty(xhat.shape)


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
@_op("fedvit::asl_loss", mutates_args=())
def asl_loss(logits: Tensor, targets: Tensor, gamma_neg: float, gamma_pos: float, clip: float,
             eps: float) -> Tuple[Tensor, Tensor]:
    """(mean asymmetric-focal loss [scalar], d loss / d logits [B,C]) in one fused pass. A target outside
    [0, C) — where the reference's F.one_hot raises (losses.py:47) — yields a NaN loss (a kernel cannot
    raise and a host-side range check would cost a device sync per step)."""
    _need_cuda(logits, targets)
    logits = logits.float().contiguous()
    targets = targets.to(torch.int64).contiguous()
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    dlogits = torch.empty_like(logits)
    LIB.call("fv_asl_loss", logits.data_ptr(), targets.data_ptr(), loss.data_ptr(),
             dlogits.data_ptr(), logits.shape[0], logits.shape[1], gamma_neg, gamma_pos, clip, eps,
             _stream(logits))
    return loss, dlogits


@asl_loss.register_fake
def _(logits, targets, gamma_neg, gamma_pos, clip, eps):
    return logits.new_empty((), dtype=torch.float32), logits.new_empty(logits.shape, dtype=torch.float32)


@_op("fedvit::ce_loss", mutates_args=())
def ce_loss(logits: Tensor, targets: Tensor) -> Tuple[Tensor, Tensor]:
    _need_cuda(logits, targets)
    logits = logits.float().contiguous()
    targets = targets.to(torch.int64).contiguous()
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    dlogits = torch.empty_like(logits)
    LIB.call("fv_ce_loss", logits.data_ptr(), targets.data_ptr(), loss.data_ptr(),
             dlogits.data_ptr(), logits.shape[0], logits.shape[1], _stream(logits))
    return loss, dlogits


@ce_loss.register_fake
def _(logits, targets):
    return logits.new_empty((), dtype=torch.float32), logits.new_empty(logits.shape, dtype=torch.float32)


# ------------------------------------------------------------------------------------------------
# flat-arena sweeps: grad norm, AdamW(+EMA+bf16 cast), EMA, FedAvg fold, cast
# ------------------------------------------------------------------------------------------------
@_op("fedvit::sumsq", mutates_args=("out",))
def sumsq(g: Tensor, out: Tensor, accumulate: bool) -> None:
    _need_cuda(g, out)
    LIB.call("fv_sumsq", g.data_ptr(), g.numel(), out.data_ptr(), int(accumulate), _stream(g))


@_op("fedvit::adamw_flat", mutates_args=("p", "g", "m", "v", "ema", "p_lp"))
def adamw_flat(p: Tensor, g: Tensor, m: Tensor, v: Tensor, seg_end: Tensor, seg_lr: Tensor,
               seg_wd: Tensor, sumsq_: Optional[Tensor], max_norm: float, beta1: float, beta2: float,
               eps: float, step: int, ema: Optional[Tensor], ema_decay: float,
               p_lp: Optional[Tensor], zero_grad: bool = False) -> None:
    """One sweep: clip coefficient, AdamW over the per-range lr / weight-decay table, optional EMA and
    bf16 re-cast; ``zero_grad`` leaves ``g`` zeroed behind the sweep (its last reader)."""
    _need_cuda(p, g, m, v, seg_end, seg_lr, seg_wd, sumsq_, ema, p_lp)
    LIB.call("fv_adamw_flat", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
             seg_end.data_ptr(), seg_lr.data_ptr(), seg_wd.data_ptr(), seg_end.numel(),
             _ptr(sumsq_), max_norm, beta1, beta2, eps, step, _ptr(ema), ema_decay, _ptr(p_lp),
             p.numel(), int(zero_grad), _stream(p))


@_op("fedvit::adamw_flat_dev", mutates_args=("p", "g", "m", "v", "ema", "p_lp"))
def adamw_flat_dev(p: Tensor, g: Tensor, m: Tensor, v: Tensor, seg_end: Tensor, seg_lr: Tensor,
                   seg_wd: Tensor, sumsq_: Optional[Tensor], max_norm: float, beta1: float, beta2: float,
                   eps: float, bias_corr: Tensor, ema: Optional[Tensor], ema_decay: float,
                   p_lp: Optional[Tensor], zero_grad: bool = False) -> None:
    """``adamw_flat`` with the step-dependent bias corrections ``[1 - beta1^t, sqrt(1 - beta2^t)]`` in a
    device tensor instead of the step number as a launch argument (CUDA-graph replay)."""
    _need_cuda(p, g, m, v, seg_end, seg_lr, seg_wd, sumsq_, bias_corr, ema, p_lp)
    if bias_corr.dtype != torch.float32 or bias_corr.numel() < 2:
        raise FedVitError("adamw_flat_dev: bias_corr must be fp32 with two entries")
    LIB.call("fv_adamw_flat_dev", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
             seg_end.data_ptr(), seg_lr.data_ptr(), seg_wd.data_ptr(), seg_end.numel(),
             _ptr(sumsq_), max_norm, beta1, beta2, eps, bias_corr.data_ptr(), _ptr(ema), ema_decay,
             _ptr(p_lp), p.numel(), int(zero_grad), _stream(p))


@_op("fedvit::adamw_tick", mutates_args=("step", "bias_corr"))
def adamw_tick(step: Tensor, bias_corr: Tensor, beta1: float, beta2: float) -> None:
    """Device-side step counter of the graph-replayed optimiser: ``step += 1`` and that step's bias
    corrections into ``bias_corr`` — captured in the graph right before ``adamw_flat_dev``."""
    _need_cuda(step, bias_corr)
    if step.dtype != torch.int64 or step.numel() != 1 or bias_corr.dtype != torch.float32 or bias_corr.numel() < 2:
        raise FedVitError("adamw_tick: step int64 [1], bias_corr fp32 [2]")
    LIB.call("fv_adamw_tick", step.data_ptr(), bias_corr.data_ptr(), beta1, beta2, _stream(step))


@_op("fedvit::scale_by_clip", mutates_args=("x",))
def scale_by_clip(x: Tensor, sumsq_: Tensor, max_norm: float) -> None:
    _need_cuda(x, sumsq_)
    LIB.call("fv_scale_inplace", x.data_ptr(), sumsq_.data_ptr(), max_norm, x.numel(), _stream(x))


@_op("fedvit::ema_update", mutates_args=("shadow",))
def ema_update(shadow: Tensor, p: Tensor, decay: float) -> None:
    _need_cuda(shadow, p)
    LIB.call("fv_ema_update", shadow.data_ptr(), p.data_ptr(), decay, p.numel(), _stream(p))


@_op("fedvit::fedavg_accum", mutates_args=("acc",))
def fedavg_accum(acc: Tensor, w: Tensor, weight: float, init: bool) -> None:
    """acc = (init ? 0 : acc) + weight * w over a flat fp32 arena."""
    _need_cuda(acc, w)
    if acc.dtype != torch.float32 or w.dtype != torch.float32 or acc.numel() != w.numel():
        raise FedVitError("fedavg_accum: fp32 arenas of equal length required")
    LIB.call("fv_fedavg_accum", acc.data_ptr(), w.data_ptr(), weight, int(init), w.numel(), _stream(w))


@_op("fedvit::fedavg_fold_into", mutates_args=("out", "out_lp"))
def fedavg_fold_into(acc: Optional[Tensor], w: Tensor, weight: float, out: Tensor, out_lp: Optional[Tensor]) -> None:
    """out = (acc if given else 0) + weight * w, same rounding as ``fedavg_accum``; ``out`` may be ``w``
    itself (the round's last fold lands in the parameter arena); optional bf16 copy of the result."""
    _need_cuda(acc, w, out, out_lp)
    if w.dtype != torch.float32 or out.dtype != torch.float32 or out.numel() != w.numel() or \
            (acc is not None and (acc.dtype != torch.float32 or acc.numel() != w.numel())):
        raise FedVitError("fedavg_fold_into: fp32 arenas of equal length required")
    if out_lp is not None and (out_lp.dtype != torch.bfloat16 or out_lp.numel() != w.numel()):
        raise FedVitError("fedavg_fold_into: out_lp must be a bf16 arena of the same length")
    LIB.call("fv_fedavg_fold_into", _ptr(acc), w.data_ptr(), weight, out.data_ptr(), _ptr(out_lp), w.numel(), _stream(w))


@_op("fedvit::cast_bf16", mutates_args=("dst",))
def cast_bf16(src: Tensor, dst: Tensor) -> None:
    _need_cuda(src, dst)
    LIB.call("fv_cast_f32_bf16", src.data_ptr(), dst.data_ptr(), src.numel(), _stream(src))


# ------------------------------------------------------------------------------------------------
# device-side batch assembly (scope row f3)
# ------------------------------------------------------------------------------------------------
MIX_NONE, MIX_MIXUP, MIX_CUTMIX = 0, 1, 2


def _mix_args(perm: Optional[Tensor], lam: float, mode: int, box, batch: int):
    if mode not in (MIX_NONE, MIX_MIXUP, MIX_CUTMIX):
        raise FedVitError("mix: mode must be 0 (none), 1 (mixup) or 2 (cutmix)")
    if mode != MIX_NONE:
        if perm is None or perm.dtype != torch.int64 or perm.numel() != batch or not perm.is_cuda:
            raise FedVitError("mix: perm must be a CUDA int64 tensor with one partner index per sample")
        perm = perm.contiguous()
    x1, y1, x2, y2 = (int(v) for v in box)
    # the two mixing weights exactly as torch forms them: the python scalars lam and (1 - lam), each
    # rounded to fp32 when it meets the fp32 tensor
    return perm, float(lam), float(1.0 - lam), x1, y1, x2, y2


@_op("fedvit::mix_batch", mutates_args=())
def mix_batch(x: Tensor, perm: Optional[Tensor], lam: float, mode: int, box: List[int]) -> Tensor:
    """MixUp (mode 1: ``lam*x + (1-lam)*x[perm]``) or CutMix (mode 2: rows box[0]:box[2] x columns
    box[1]:box[3] taken from ``x[perm]``) of an fp32 NCHW batch in one pass — reference
    utils.py:112-150. Bit-identical to the reference's ATen expression."""
    _need_cuda(x, perm)
    if x.dim() != 4 or x.dtype != torch.float32 or not x.is_contiguous() or x.shape[3] % 4:
        raise FedVitError("mix_batch: x must be contiguous fp32 [B,C,H,W] with W % 4 == 0")
    b, c, h, w = x.shape
    perm, lam_, oml, x1, y1, x2, y2 = _mix_args(perm, lam, mode, box, b)
    out = torch.empty_like(x)
    LIB.call("fv_mix_batch", x.data_ptr(), _ptr(perm), lam_, oml, mode, x1, y1, x2, y2, out.data_ptr(), b, c, h, w,
             _stream(x))
    return out


@mix_batch.register_fake
def _(x, perm, lam, mode, box):
    return torch.empty_like(x)


@_op("fedvit::assemble_batch", mutates_args=())
def assemble_batch(img_u8: Tensor, mask_u8: Optional[Tensor], mean: List[float], std: List[float],
                   perm: Optional[Tensor], lam: float, mode: int, box: List[int]) -> Tensor:
    """uint8 images ([B,3,H,W] or [B,H,W,3]) + optional uint8 masks [B,H,W] -> normalised fp32 NCHW
    with 3 or 4 channels (reference data.py:148-155, 222-224), optionally MixUp / CutMix-ed in the
    same pass."""
    _need_cuda(img_u8, mask_u8, perm)
    if img_u8.dim() != 4 or img_u8.dtype != torch.uint8 or not img_u8.is_contiguous():
        raise FedVitError("assemble_batch: images must be a contiguous uint8 4-D tensor")
    nhwc = img_u8.shape[3] == 3 and img_u8.shape[1] != 3
    if not nhwc and img_u8.shape[1] != 3:
        raise FedVitError("assemble_batch: images must be [B,3,H,W] or [B,H,W,3]")
    b = img_u8.shape[0]
    h, w = (img_u8.shape[1], img_u8.shape[2]) if nhwc else (img_u8.shape[2], img_u8.shape[3])
    if w % 4:
        raise FedVitError("assemble_batch: width must be a multiple of 4")
    if mask_u8 is not None and (mask_u8.dtype != torch.uint8 or tuple(mask_u8.shape) != (b, h, w)
                                or not mask_u8.is_contiguous()):
        raise FedVitError("assemble_batch: masks must be contiguous uint8 [B,H,W]")
    if len(mean) != 3 or len(std) != 3:
        raise FedVitError("assemble_batch: mean / std need three values")
    perm, lam_, oml, x1, y1, x2, y2 = _mix_args(perm, lam, mode, box, b)
    import ctypes

    mean_c = (ctypes.c_float * 3)(*[float(v) for v in mean])
    std_c = (ctypes.c_float * 3)(*[float(v) for v in std])
    out = torch.empty((b, 4 if mask_u8 is not None else 3, h, w), device=img_u8.device, dtype=torch.float32)
    LIB.call("fv_assemble_batch", img_u8.data_ptr(), int(nhwc), _ptr(mask_u8), ctypes.addressof(mean_c),
             ctypes.addressof(std_c), _ptr(perm), lam_, oml, mode, x1, y1, x2, y2, out.data_ptr(), b, h, w,
             _stream(img_u8))
    return out


@assemble_batch.register_fake
def _(img_u8, mask_u8, mean, std, perm, lam, mode, box):
    nhwc = img_u8.shape[3] == 3 and img_u8.shape[1] != 3
    b = img_u8.shape[0]
    h, w = (img_u8.shape[1], img_u8.shape[2]) if nhwc else (img_u8.shape[2], img_u8.shape[3])
    return img_u8.new_empty((b, 4 if mask_u8 is not None else 3, h, w), dtype=torch.float32)
