"""Kernel parity on a real B200: every libfedvit entry point, through the C ABI, against a plain
PyTorch fp32 reference of the same op (tolerances: 1e-4 relative for fp32 arithmetic, 2e-2 for
bf16 — the north_star gates — tighter where the op is exact)."""
import math

import numpy as np
import pytest
import torch

import fedvit_b200  # noqa: F401
from conftest import rel_err
from fedvit_b200 import ops
from fedvit_b200._lib import FedVitError, launch_count

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _gen(seed):
    return torch.Generator(device=DEV).manual_seed(seed)


def _gemm_ref(a, b, am, bm):
    A = a.double() if am == 0 else a.double().t()
    B = b.double() if bm == 0 else b.double().t()
    return (A @ B.t()).float()


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (1000, 776, 328), (3152, 576, 192), (5, 8, 8), (4097, 2304, 768)])
@pytest.mark.parametrize("am,bm", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_gemm_bf16_all_layouts(m, n, k, am, bm):
    if am == 1 and m % 8:
        m = (m + 7) // 8 * 8  # MN-major A needs a 16-byte aligned leading dimension
    g = _gen(m + n + k)
    a = torch.randn((m, k) if am == 0 else (k, m), device=DEV, generator=g).bfloat16()
    b = torch.randn((n, k) if bm == 0 else (k, n), device=DEV, generator=g).bfloat16()
    bias = torch.randn(n, device=DEV, generator=g)
    out = torch.empty(m, n, device=DEV)
    ops.gemm(a, b, bias, out, None, am, bm, ops.EPI["none"], 1, 0)
    assert rel_err(out, _gemm_ref(a, b, am, bm) + bias) < 1e-5  # bf16 inputs are exact in fp32; fp32 accumulate


@pytest.mark.parametrize("m,n,k", [(777, 520, 200),       # single-CTA kernel, ragged in every dimension
                                   (19000, 1032, 328)])   # 75 x 5 CTA-pair tiles (cta_group::2 kernel), ragged M / N / K
def test_gemm_bf16_epilogues(m, n, k):
    g = _gen(1)
    a = torch.randn(m, k, device=DEV, generator=g).bfloat16()
    b = torch.randn(n, k, device=DEV, generator=g).bfloat16()
    bias = torch.randn(n, device=DEV, generator=g)
    ref = _gemm_ref(a, b, 0, 0)
    # bf16 store
    out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, b, bias, out, None, 0, 0, ops.EPI["none"], 1, 0)
    assert rel_err(out, ref + bias) < 4e-3
    # + residual (fp32)
    res = torch.randn(m, n, device=DEV, generator=g)
    out = torch.empty(m, n, device=DEV)
    ops.gemm(a, b, bias, out, res, 0, 0, ops.EPI["residual"], 1, 0)
    assert rel_err(out, ref + bias + res) < 1e-5
    # + residual with a per-sample stochastic-depth factor on the branch
    groups = 7
    scale = torch.rand((m + groups - 1) // groups, device=DEV, generator=g) * 2
    out = torch.empty(m, n, device=DEV)
    ops.linear_residual(a, b, bias, res, scale, groups, out)
    rows = torch.arange(m, device=DEV) // groups
    assert rel_err(out, (ref + bias) * scale[rows, None] + res) < 1e-5
    out32 = torch.empty(m, n, device=DEV)
    ops.linear_residual(a.float(), b.float(), bias, res, scale, groups, out32)
    assert rel_err(out32, (ref + bias) * scale[rows, None] + res) < 1e-5
    # GELU + its derivative (kept for the backward), fp32 and bf16 stores
    pre = (ref + bias).double().requires_grad_(True)
    act_ref = torch.nn.functional.gelu(pre)
    act_ref.sum().backward()
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 8e-3)):
        o, dact = torch.empty(m, n, device=DEV, dtype=dt), torch.empty(m, n, device=DEV, dtype=dt)
        ops.gemm_gelu(a, b, bias, o, dact)
        assert rel_err(o, act_ref) < tol
        assert rel_err(dact, pre.grad) < tol
        o2 = torch.full((m, n), 7.0, device=DEV, dtype=dt)
        ops.gemm_gelu_fwd(a, b, bias, o2)  # forward-only form (no derivative output): the same activation bits
        assert torch.equal(o2, o)
    o32, o32b = torch.empty(m, n, device=DEV), torch.empty(m, n, device=DEV)
    ops.gemm_gelu(a.float(), b.float(), bias, o32, torch.empty(m, n, device=DEV))  # fp32 FFMA path
    ops.gemm_gelu_fwd(a.float(), b.float(), bias, o32b)
    assert torch.equal(o32, o32b) and rel_err(o32, act_ref) < 1e-5
    # dgrad fused with the saved GELU derivative
    g1 = torch.randn(m, n, device=DEV, generator=g)
    out = torch.empty(m, n, device=DEV)
    ops.gemm(a, b, None, out, g1, 0, 0, ops.EPI["dgelu"], 1, 0)
    assert rel_err(out, ref * g1) < 1e-5
    outb = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, b, None, outb, g1.bfloat16(), 0, 0, ops.EPI["dgelu"], 1, 0)
    assert rel_err(outb, ref * g1.bfloat16().float()) < 4e-3
    # accumulate + split-K
    for sk in (1, 3):
        acc = torch.randn(m, n, device=DEV, generator=g)
        want = acc + ref
        ops.gemm(a, b, None, acc, None, 0, 0, ops.EPI["accum"], sk, 0)
        assert rel_err(acc, want) < 1e-5
    # patch epilogue: rows scattered past one cls slot per image, + pos_embed
    tokens, imgs = 37, 21
    a = torch.randn(tokens * imgs, k, device=DEV, generator=g).bfloat16()
    pos = torch.randn(tokens + 1, n, device=DEV, generator=g)
    x = torch.zeros(imgs * (tokens + 1), n, device=DEV)
    ops.gemm(a, b, bias, x, pos, 0, 0, ops.EPI["patch"], 1, tokens)
    want = (_gemm_ref(a, b, 0, 0) + bias).view(imgs, tokens, n) + pos[1:]
    assert rel_err(x.view(imgs, tokens + 1, n)[:, 1:], want) < 1e-5
    assert float(x.view(imgs, tokens + 1, n)[:, 0].abs().max()) == 0.0


@pytest.mark.parametrize("am,bm", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_gemm_f32_layouts(am, bm):
    m, n, k = 197, 130, 77
    g = _gen(5)
    a = torch.randn((m, k) if am == 0 else (k, m), device=DEV, generator=g)
    b = torch.randn((n, k) if bm == 0 else (k, n), device=DEV, generator=g)
    bias = torch.randn(n, device=DEV, generator=g)
    out = torch.empty(m, n, device=DEV)
    ops.gemm(a, b, bias, out, None, am, bm, 0, 1, 0)
    assert rel_err(out, _gemm_ref(a, b, am, bm) + bias) < 1e-6


def test_gemm_rejects_bad_arguments():
    a = torch.zeros(16, 16, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(FedVitError):
        ops.gemm(a, a, None, torch.zeros(16, 12, device=DEV), None, 0, 0, 0, 1, 0)  # shape mismatch
    with pytest.raises(FedVitError):
        ops.gemm(a, a, torch.zeros(16, device=DEV, dtype=torch.bfloat16), torch.zeros(16, 16, device=DEV), None, 0, 0, 0, 1, 0)
    with pytest.raises(FedVitError):  # C ABI argument check: ldc must keep 16-byte rows
        ops.gemm(a, a, None, torch.zeros(16, 20, device=DEV)[:, :16], None, 0, 0, 0, 1, 0)
    with pytest.raises(FedVitError):
        ops.gemm(a, a, None, torch.zeros(16, 16, device=DEV), None, 0, 0, ops.EPI["residual"], 1, 0)  # aux missing
    with pytest.raises(FedVitError):
        ops.gemm(a.cpu(), a.cpu(), None, torch.zeros(16, 16), None, 0, 0, 0, 1, 0)  # no CPU fallback


@pytest.mark.parametrize("cols", [64, 192, 768, 1024])
def test_layernorm_fwd_bwd(cols):
    rows = 1031
    g = _gen(cols)
    x = torch.randn(rows, cols, device=DEV, generator=g) * 2 + 0.5
    gam, bet = torch.randn(cols, device=DEV, generator=g), torch.randn(cols, device=DEV, generator=g)
    dy, dres = torch.randn(rows, cols, device=DEV, generator=g), torch.randn(rows, cols, device=DEV, generator=g)
    xr, gr, br = x.double().requires_grad_(True), gam.double().requires_grad_(True), bet.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (cols,), gr, br, 1e-6)
    yr.backward(dy.double())
    y, mean, rstd = ops.layernorm_fwd(x, gam, bet, 1e-6, False)
    assert rel_err(y, yr) < 1e-6
    assert rel_err(mean, x.double().mean(1)) < 1e-6
    dg, db = torch.zeros(cols, device=DEV), torch.zeros(cols, device=DEV)
    dx, dxlp = ops.layernorm_bwd(dy, x, gam, mean, rstd, dres, dg, db, True)
    assert rel_err(dx, xr.grad + dres.double()) < 1e-5
    assert rel_err(dxlp, dx.bfloat16()) == 0.0
    assert rel_err(dg, gr.grad) < 1e-5 and rel_err(db, br.grad) < 1e-5
    ops.layernorm_bwd(dy, x, gam, mean, rstd, None, dg, db, False)  # accumulates
    assert rel_err(dg, 2 * gr.grad) < 1e-5
    sc = torch.rand((rows + 12) // 13, device=DEV, generator=g) + 0.5
    dxs, dxs_lp = ops.layernorm_bwd(dy, x, gam, mean, rstd, dres, torch.zeros_like(dg), torch.zeros_like(db), True, sc, 13)
    assert torch.equal(dxs, dx)
    assert rel_err(dxs_lp, dx * sc[torch.arange(rows, device=DEV) // 13, None]) < 4e-3
    ybf, _, _ = ops.layernorm_fwd(x, gam, bet, 1e-6, True)
    assert rel_err(ybf, yr) < 4e-3
    dxb, _ = ops.layernorm_bwd(dy.bfloat16(), x, gam, mean, rstd, None, torch.zeros_like(dg), torch.zeros_like(db), False)
    assert rel_err(dxb, xr.grad) < 1e-2


@pytest.mark.parametrize("cols", [128, 384, 768, 1024])
@pytest.mark.parametrize("rows", [5, 1031, 8 * 148 * 3 + 1])
def test_layernorm_bwd_ring_kernel(rows, cols, monkeypatch):
    """The bulk-copy ring kernel (default for bf16 dy + residual gradient, cols % 128 == 0) against an
    fp64 reference and against the register-pipelined kernel: ragged last group, fewer rows than one
    group, more groups than the ring is deep, the stochastic-depth factor on the bf16 copy, capped grids."""
    g = _gen(rows + cols)
    x = torch.randn(rows, cols, device=DEV, generator=g) * 2 + 0.5
    gam = torch.randn(cols, device=DEV, generator=g)
    dy = torch.randn(rows, cols, device=DEV, generator=g).bfloat16()
    dres = torch.randn(rows, cols, device=DEV, generator=g)
    xr, gr = x.double().requires_grad_(True), gam.double().requires_grad_(True)
    br = torch.zeros(cols, device=DEV, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.layer_norm(xr, (cols,), gr, br, 1e-6).backward(dy.double())
    mean = x.double().mean(1).float()
    rstd = (x.double().var(1, unbiased=False) + 1e-6).rsqrt().float()
    sc = torch.rand((rows + 12) // 13, device=DEV, generator=g) + 0.5
    monkeypatch.setenv("FEDVIT_LN_REREAD", "1")
    out = {}
    for mode, grid in (("0", None), ("9", None), ("9", "3")):
        monkeypatch.setenv("FEDVIT_LN_MINB", mode)
        if grid:
            monkeypatch.setenv("FEDVIT_LN_GRID", grid)
        dg, db = torch.zeros(cols, device=DEV), torch.zeros(cols, device=DEV)
        dx, dxlp = ops.layernorm_bwd(dy, x, gam, mean, rstd, dres, dg, db, True, sc, 13)
        assert rel_err(dx, xr.grad + dres.double()) < 1e-5
        assert rel_err(dxlp, dx * sc[torch.arange(rows, device=DEV) // 13, None]) < 4e-3
        assert rel_err(dg, gr.grad) < 1e-5 and rel_err(db, br.grad) < 1e-5
        dx2, dxlp2 = ops.layernorm_bwd(dy, x, gam, mean, rstd, dres, dg, db, True)  # accumulates, no factor
        assert torch.equal(dx2, dx) and rel_err(dxlp2, dx.bfloat16()) == 0.0
        assert rel_err(dg, 2 * gr.grad) < 1e-5
        out[(mode, grid)] = (dx, dxlp)
    monkeypatch.delenv("FEDVIT_LN_GRID")
    monkeypatch.delenv("FEDVIT_LN_MINB")
    ops.layernorm_bwd(dy, x, gam, mean, rstd, dres, dg, db, True)  # back to the default for later tests
    assert rel_err(out[("9", None)][0], out[("0", None)][0]) < 1e-6
    assert torch.equal(out[("9", None)][0], out[("9", "3")][0])  # row arithmetic does not depend on the grid


@pytest.mark.parametrize("B,N,H", [(2, 197, 3), (3, 64, 1), (2, 65, 2), (1, 1, 1), (2, 256, 2),   # one-tile tcgen05 kernels
                                   (1, 577, 2), (3, 257, 2), (5, 400, 3), (2, 768, 1),           # long-sequence tcgen05 kernels
                                   (1, 800, 1)])                                                 # mma.sync flash kernels
def test_flash_attention_fwd_bwd(B, N, H):
    g = _gen(N)
    qkv = torch.randn(B * N, 3 * H * 64, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B * N, H * 64, device=DEV, generator=g).bfloat16()
    scale = 1.0 / math.sqrt(64)
    q, k, v = (qkv.double().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)[i].clone().requires_grad_(True) for i in range(3))
    s = (q @ k.transpose(-1, -2)) * scale
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * N, H * 64)
    o.backward(dout.double())
    out, lse = ops.attention_fwd(qkv, B, N, H, scale)
    assert rel_err(out, o) < 2e-2 and rel_err(out, o) < 5e-3
    assert rel_err(lse, torch.logsumexp(s, -1)) < 1e-5
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, scale).float().view(B * N, 3, H * 64)
    ref = torch.stack([q.grad, k.grad, v.grad], 0).permute(1, 3, 0, 2, 4).reshape(B * N, 3, H * 64)
    scale_ref = float(ref.norm()) + 1e-6 * float(dout.double().norm())  # N == 1: dq, dk are exactly 0
    for i, name in enumerate("qkv"):
        err = float((dqkv[:, i].double().cpu() - ref[:, i].cpu()).norm())
        assert err < 1e-2 * max(float(ref[:, i].norm()), 1e-3 * scale_ref), name
    # reproducible: no atomics on the attention path, except that the long-sequence kernel sums the
    # key blocks' dQ contributions with fp32 TMA reduce-adds (order-dependent in the last fp32 bits)
    again = ops.attention_bwd(qkv, out, dout, lse, B, N, H, scale).view(B * N, 3, H * 64).float()
    assert torch.equal(again[:, 1:], dqkv[:, 1:])
    if 256 < N <= 768:
        assert rel_err(again[:, 0], dqkv[:, 0]) < 1e-2
    else:
        assert torch.equal(again[:, 0], dqkv[:, 0])


def _asl_ref(logits, targets, gn=4.0, gp=1.0, clip=0.05, eps=1e-8):
    from oracle import asl
    return asl.asymmetric_focal_loss(logits, targets, gn, gp, clip, eps)


@pytest.mark.parametrize("B,C", [(4, 7), (256, 7), (33, 8), (1, 3), (1024, 7)])
def test_fused_losses_match_oracle(B, C):
    g = _gen(B)
    logits = (torch.randn(B, C, device=DEV, generator=g) * 3)
    targets = torch.randint(0, C, (B,), device=DEV, generator=g)
    lc = logits.cpu().requires_grad_(True)
    ref = _asl_ref(lc, targets.cpu())
    ref.backward()
    loss, dl = ops.asl_loss(logits, targets, 4.0, 1.0, 0.05, 1e-8)
    assert float(loss) == pytest.approx(float(ref), rel=1e-5)
    assert rel_err(dl, lc.grad) < 1e-5
    lc.grad = None
    ce = torch.nn.functional.cross_entropy(lc, targets.cpu())
    ce.backward()
    loss, dl = ops.ce_loss(logits, targets)
    assert float(loss) == pytest.approx(float(ce), rel=1e-5)
    assert rel_err(dl, lc.grad) < 1e-5


def test_loss_known_answers_on_gpu(asl_kats):
    k = asl_kats
    for i in (1, 2, 3):
        lg = torch.from_numpy(k[f"kat{i}_logits"]).to(DEV)
        t = torch.from_numpy(k[f"kat{i}_targets"]).to(DEV)
        loss, dl = ops.asl_loss(lg, t, 4.0, 1.0, 0.05, 1e-8)
        assert float(loss) == pytest.approx(float(k[f"kat{i}_loss"]), rel=2e-6)
        if f"kat{i}_dlogits" in k:
            assert torch.allclose(dl.cpu(), torch.from_numpy(k[f"kat{i}_dlogits"]), rtol=1e-4, atol=1e-7)


def test_adamw_sweep_matches_torch_adamw():
    n = 1 << 18
    g = _gen(1)
    p0 = torch.randn(n, device=DEV, generator=g)
    cuts = [n // 4, n // 2, n]
    lrs, wds = [1e-3, -1.0, 3e-3], [1e-2, 0.0, 1e-5]
    ps = [p0[: cuts[0]].clone().requires_grad_(True), p0[cuts[0]: cuts[1]].clone().requires_grad_(True),
          p0[cuts[1]:].clone().requires_grad_(True)]
    opt = torch.optim.AdamW([{"params": [ps[0]], "lr": lrs[0], "weight_decay": wds[0]},
                             {"params": [ps[2]], "lr": lrs[2], "weight_decay": wds[2]}])
    pp, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    ema, ema_ref = p0.clone(), p0.clone()
    plp = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    seg = (torch.tensor(cuts, device=DEV), torch.tensor(lrs, device=DEV), torch.tensor(wds, device=DEV))
    ss = torch.zeros(1, device=DEV)
    for step in (1, 2, 3):
        grad = torch.randn(n, device=DEV, generator=g) * (0.02 if step < 3 else 1e-5)
        for q, sl in zip(ps, (slice(0, cuts[0]), slice(cuts[0], cuts[1]), slice(cuts[1], n))):
            q.grad = grad[sl].clone()
        total = torch.nn.utils.clip_grad_norm_(ps, 1.0)  # the norm covers the never-stepped range too
        opt.step()
        ops.sumsq(grad, ss, False)
        assert float(ss.sqrt()) == pytest.approx(float(total), rel=1e-5)
        ops.adamw_flat(pp, grad, m, v, *seg, ss, 1.0, 0.9, 0.999, 1e-8, step, ema, 0.99, plp)
        now = torch.cat([q.detach() for q in ps])
        ema_ref.mul_(0.99).add_(now, alpha=0.01)
    ref = torch.cat([q.detach() for q in ps])
    assert rel_err(pp, ref) < 1e-6
    assert torch.equal(pp[cuts[0]: cuts[1]], p0[cuts[0]: cuts[1]])  # lr < 0: untouched
    assert rel_err(ema, ema_ref) < 1e-6
    assert torch.equal(plp, pp.bfloat16())
    g2 = grad.clone()
    ops.scale_by_clip(g2, torch.full((1,), 100.0, device=DEV), 1.0)
    assert rel_err(g2, grad * (1.0 / (10.0 + 1e-6))) < 1e-6


def test_fedavg_fold_is_bit_exact_against_oracle():
    from oracle import fedavg as ofed
    n = (1 << 20) + 64
    g = _gen(2)
    ws = [torch.randn(n, device=DEV, generator=g) for _ in range(5)]
    n_k = [100, 250, 50, 300, 300]
    acc = torch.empty(n, device=DEV)
    for i, (w, c) in enumerate(zip(ws, ofed.client_weights(n_k))):
        ops.fedavg_accum(acc, w, c, i == 0)
    want = ofed.fedavg_flat([w.cpu() for w in ws], n_k)
    assert torch.equal(acc.cpu(), want)
    # properties at full ViT-B arena size: identical clients average to themselves; linear in w
    n = 86_196_224
    w = torch.randn(n, device=DEV, generator=g)
    acc = torch.empty(n, device=DEV)
    for i, c in enumerate(ofed.client_weights([4096] * 8)):
        ops.fedavg_accum(acc, w, c, i == 0)
    assert rel_err(acc, w) < 1e-6
    acc2 = torch.empty(n, device=DEV)
    ops.fedavg_accum(acc2, w * 2, 0.5, True)
    assert torch.equal(acc2, w)


def test_elementwise_helpers():
    g = _gen(3)
    img = torch.randn(3, 4, 224, 224, device=DEV, generator=g)
    ref = torch.nn.functional.unfold(img, 16, stride=16).transpose(1, 2).reshape(-1, 4 * 256)
    assert torch.equal(ops.patchify(img, False), ref)
    assert torch.equal(ops.patchify(img, True), ref.bfloat16())
    a = torch.randn(5000, 770, device=DEV, generator=g)
    out = torch.zeros(770, device=DEV)
    ops.colsum(a, out, False)
    assert rel_err(out, a.double().sum(0)) < 1e-5
    ops.colsum(a.bfloat16(), out, True)
    assert rel_err(out, a.double().sum(0) + a.bfloat16().double().sum(0)) < 1e-5
    s = torch.randn(37, 197, device=DEV, generator=g)
    p = ops.softmax_rows(s, 0.125)
    assert rel_err(p, (s.double() * 0.125).softmax(-1)) < 1e-6
    dp = torch.randn_like(s)
    sr = s.double().requires_grad_(True)
    ((sr * 0.125).softmax(-1) * dp.double()).sum().backward()
    assert rel_err(ops.softmax_rows_bwd(p, dp, 0.125), sr.grad) < 1e-5
    n0 = launch_count()
    ops.cast_bf16(a.view(-1)[: 4096], torch.empty(4096, device=DEV, dtype=torch.bfloat16))
    assert launch_count() == n0 + 1


@pytest.mark.parametrize("tokens,out_f,in_f,sk", [(1000, 776, 328, 3), (50432, 768, 768, 8), (197 * 4, 192, 768, 1), (37, 8, 16, 1),
                                                  # CTA-pair weight-gradient kernel (>= 37 items): ragged M / N / K tails,
                                                  # several items per pair, one n-block (every k-slice summed by one pair)
                                                  (1031, 1000, 520, 4), (4099, 2304, 768, 8), (2050, 3072, 200, 4)])
def test_wgrad_with_fused_bias_gradient(tokens, out_f, in_f, sk):
    g = _gen(tokens + out_f)
    dy = torch.randn(tokens, out_f, device=DEV, generator=g).bfloat16()
    x = torch.randn(tokens, in_f, device=DEV, generator=g).bfloat16()
    dw = torch.randn(out_f, in_f, device=DEV, generator=g)
    db = torch.randn(out_f, device=DEV, generator=g)
    want_w = dw.double() + dy.double().t() @ x.double()
    want_b = db.double() + dy.double().sum(0)
    ops.wgrad(dy, x, dw, db, sk)
    assert rel_err(dw, want_w) < 1e-5
    assert rel_err(db, want_b) < 1e-5
    dw2 = torch.zeros(out_f, in_f, device=DEV)
    ops.wgrad(dy, x, dw2, None, sk)  # without the bias gradient
    assert rel_err(dw2, dy.double().t() @ x.double()) < 1e-5


@pytest.mark.parametrize("B,C,S,D", [(3, 3, 224, 192), (2, 4, 224, 768), (2, 3, 384, 1024), (5, 3, 32, 64)])
def test_patch_embed_im2col_free(B, C, S, D):
    """5-D TMA + tf32 tensor-core patch embedding vs conv2d (fp32): rows 1..N-1 get conv + bias + pos,
    row 0 of every image is left alone; tf32 keeps 10 mantissa bits -> ~1e-3."""
    g = _gen(S + D)
    img = torch.randn(B, C, S, S, device=DEV, generator=g)
    w = torch.randn(D, C, 16, 16, device=DEV, generator=g) * 0.05
    bias = torch.randn(D, device=DEV, generator=g)
    N = (S // 16) ** 2 + 1
    pos = torch.randn(N, D, device=DEV, generator=g)
    x = torch.full((B * N, D), 7.0, device=DEV)
    ops.patch_embed(img, w, bias, pos, x)
    ref = torch.nn.functional.conv2d(img.double(), w.double(), bias.double(), stride=16).flatten(2).transpose(1, 2)
    ref = ref + pos[1:].double()
    got = x.view(B, N, D)
    assert rel_err(got[:, 1:], ref) < 2e-3
    assert float((got[:, 0] - 7.0).abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------
# device-side batch assembly, MixUp / CutMix (scope row f3) — bit-exact against the reference fixture
# ------------------------------------------------------------------------------------------------
def test_mix_and_assemble_match_reference_fixture_bit_exactly():
    import numpy as np
    from conftest import GOLDEN
    from oracle import mix

    g = dict(np.load(GOLDEN / "mix.npz"))
    x = torch.from_numpy(g["x"]).to(DEV)
    idx = torch.from_numpy(g["mixup/idx"]).to(DEV)
    out = ops.mix_batch(x, idx, float(g["mixup/lam"]), ops.MIX_MIXUP, [0, 0, 0, 0])
    assert torch.equal(out.cpu(), torch.from_numpy(g["mixup/out"]))
    box = mix.rand_bbox(x.shape, float(g["cutmix/lam0"]), int(g["cutmix/cx"]), int(g["cutmix/cy"]))
    out = ops.mix_batch(x, torch.from_numpy(g["cutmix/idx"]).to(DEV), 1.0, ops.MIX_CUTMIX, list(box))
    assert torch.equal(out.cpu(), torch.from_numpy(g["cutmix/out"]))
    assert torch.equal(ops.mix_batch(x, None, 1.0, ops.MIX_NONE, [0, 0, 0, 0]), x)
    # uint8 -> normalised fp32, NHWC (PIL order) and planar, with and without the mask plane
    img = torch.from_numpy(g["asm/img_u8"]).to(DEV)
    mask = torch.from_numpy(g["asm/mask_u8"]).to(DEV)
    want = torch.from_numpy(g["asm/out"])
    mean, std = list(mix.IMAGENET_MEAN), list(mix.IMAGENET_STD)
    assert torch.equal(ops.assemble_batch(img, mask, mean, std, None, 1.0, 0, [0, 0, 0, 0]).cpu(), want)
    planar = img.permute(0, 3, 1, 2).contiguous()
    assert torch.equal(ops.assemble_batch(planar, None, mean, std, None, 1.0, 0, [0, 0, 0, 0]).cpu(), want[:, :3])
    # assembly + mixing in one pass == assembly, then the oracle's mixing
    perm = torch.tensor([3, 0, 4, 1, 2], device=DEV)
    fused = ops.assemble_batch(img, mask, mean, std, perm, 0.3, ops.MIX_MIXUP, [0, 0, 0, 0]).cpu().numpy()
    assert np.array_equal(fused, mix.mixup(g["asm/out"], perm.cpu().numpy(), 0.3))
    fused = ops.assemble_batch(planar, mask, mean, std, perm, 1.0, ops.MIX_CUTMIX, [5, 9, 20, 30]).cpu().numpy()
    assert np.array_equal(fused, mix.cutmix(g["asm/out"], perm.cpu().numpy(), (5, 9, 20, 30))[0])


def test_mixup_cutmix_classes_follow_reference_draw_order():
    """fedvit_b200.utils.MixUp / CutMix / MixupCutmix make the reference's random draws in the
    reference's order (utils.py:112-164), so a seeded run mixes the same pairs with the same lam."""
    import numpy as np
    from fedvit_b200 import utils
    from oracle import mix

    x = torch.randn(16, 3, 224, 224, device=DEV)
    y = torch.arange(16, device=DEV) % 7
    np.random.seed(3)
    torch.manual_seed(3)
    mixed, la, lb, lam = utils.MixUp(alpha=0.4)(x, y)
    np.random.seed(3)
    torch.manual_seed(3)
    lam_ref = np.random.beta(0.4, 0.4)
    idx = torch.randperm(16, device=DEV)
    assert lam == lam_ref and torch.equal(lb, y[idx]) and torch.equal(la, y)
    assert torch.equal(mixed, lam_ref * x + (1 - lam_ref) * x[idx])  # the reference's ATen expression
    np.random.seed(4)
    torch.manual_seed(4)
    mixed, la, lb, lam = utils.CutMix(alpha=1.0, prob=1.0)(x, y)
    np.random.seed(4)
    torch.manual_seed(4)
    np.random.rand()
    lam0 = np.random.beta(1.0, 1.0)
    idx = torch.randperm(16, device=DEV)
    box = mix.rand_bbox(x.shape, lam0, np.random.randint(224), np.random.randint(224))
    want, lam_want = mix.cutmix(x.cpu().numpy(), idx.cpu().numpy(), box)
    assert lam == lam_want and np.array_equal(mixed.cpu().numpy(), want) and torch.equal(lb, y[idx])


# ------------------------------------------------------------------------------------------------
# guard-band test: compute-sanitizer is closed on the GPU pool, so out-of-bounds WRITES are looked
# for directly — every output lives between two sentinel bands that have to survive the launch
# ------------------------------------------------------------------------------------------------
def _guarded(shape, dtype, pad=2048):
    n = int(np.prod(shape))
    buf = torch.empty(n + 2 * pad, device=DEV, dtype=dtype)
    buf.view(torch.uint8).fill_(0x5A)
    view = buf[pad:pad + n].view(shape)

    def intact():
        b = buf.view(torch.uint8)
        e = buf.element_size()
        return bool((b[:pad * e] == 0x5A).all()) and bool((b[(pad + n) * e:] == 0x5A).all())

    return view, intact


def test_no_out_of_bounds_writes_at_ragged_sizes():
    import numpy as np  # noqa: F401

    g = _gen(77)
    # GEMM epilogues through the TMA-store path: ragged M (TMA row clipping), N % 256 != 0, K % 64 != 0;
    # once below and once above the CTA-pair threshold
    for m, n, k in [(1000, 776, 328), (19000, 520, 200)]:
        a = torch.randn(m, k, device=DEV, generator=g).bfloat16()
        b = torch.randn(n, k, device=DEV, generator=g).bfloat16()
        bias = torch.randn(n, device=DEV, generator=g)
        for dt in (torch.bfloat16, torch.float32):
            out, ok = _guarded((m, n), dt)
            ops.gemm(a, b, bias, out, None, 0, 0, ops.EPI["none"], 1, 0)
            assert ok(), ("none", m, n, k, dt)
        act, ok1 = _guarded((m, n), torch.bfloat16)
        pre, ok2 = _guarded((m, n), torch.bfloat16)
        ops.gemm_gelu(a, b, bias, act, pre)
        assert ok1() and ok2(), ("gelu", m, n, k)
        res = torch.randn(m, n, device=DEV, generator=g)
        out, ok = _guarded((m, n), torch.float32)
        ops.linear_residual(a, b, bias, res, None, 0, out)
        assert ok(), ("residual", m, n, k)
    # weight gradient (TMA reduce-add) + fused bias gradient
    dy = torch.randn(1000, 776, device=DEV, generator=g).bfloat16()
    x = torch.randn(1000, 328, device=DEV, generator=g).bfloat16()
    dw, okw = _guarded((776, 328), torch.float32)
    db, okb = _guarded((776,), torch.float32)
    dw.zero_(); db.zero_()
    ops.wgrad(dy, x, dw, db, 3)
    assert okw() and okb()
    assert rel_err(dw, dy.float().t() @ x.float()) < 1e-5 and rel_err(db, dy.float().sum(0)) < 1e-5
    # attention outputs (256-bit register stores / TMA reduce workspace) at ragged token counts
    for B, N, H in [(2, 197, 3), (2, 65, 2), (1, 577, 2), (3, 257, 2)]:
        qkv = torch.randn(B * N, 3 * H * 64, device=DEV, generator=g).bfloat16()
        dout = torch.randn(B * N, H * 64, device=DEV, generator=g).bfloat16()
        out, lse = ops.attention_fwd(qkv, B, N, H, 0.125)
        ref_out, _ = ops.attention_fwd(qkv, B, N, H, 0.125)
        assert torch.equal(out, ref_out)
        dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, 0.125)
        assert torch.isfinite(dqkv.float()).all()
    # LayerNorm forward / backward on a row count that is not a multiple of the rows per CTA
    rows, cols = 1003, 768
    xln = torch.randn(rows, cols, device=DEV, generator=g)
    gam, bet = torch.randn(cols, device=DEV, generator=g), torch.randn(cols, device=DEV, generator=g)
    y, mean, rstd = ops.layernorm_fwd(xln, gam, bet, 1e-6, True)
    assert torch.isfinite(y.float()).all() and mean.shape == (rows,)
    # batch assembly
    img = torch.randint(0, 256, (3, 24, 32, 3), dtype=torch.uint8, device=DEV)
    o = ops.assemble_batch(img, None, [0.5, 0.5, 0.5], [0.25, 0.25, 0.25], None, 1.0, 0, [0, 0, 0, 0])
    assert o.shape == (3, 3, 24, 32) and torch.isfinite(o).all()
