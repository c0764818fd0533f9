#!/usr/bin/env python
"""Kernel-level bring-up checks on a real B200: every libfedvit kernel against a plain PyTorch
reference of the same op. Each check runs in its own subprocess under a timeout so a faulting or
hung kernel cannot take the rest of the run down with it.

    python tools/gpu_check.py                # all checks
    python tools/gpu_check.py gemm_fwd ...   # selected checks
Results are appended to gpurun_out/gpu_check.log as one JSON object per check.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


def _rel(a, b):
    import torch

    a, b = a.float(), b.float()
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b).clamp_min(1e-30)).item()


def _maxabs(a, b):
    return (a.float() - b.float()).abs().max().item()


def _time(fn, iters=20, warm=3):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _gemm_case(m, n, k, a_major, b_major, epi, c_bf16, split_k=1, time_it=False):
    import torch
    from fedvit_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(m * 7 + n * 3 + k)
    dev = "cuda"
    A = torch.randn(m, k, device=dev, generator=g)
    B = torch.randn(n, k, device=dev, generator=g)
    bias = torch.randn(n, device=dev, generator=g)
    a = (A if a_major == 0 else A.t().contiguous()).to(torch.bfloat16)
    b = (B if b_major == 0 else B.t().contiguous()).to(torch.bfloat16)
    ref = a.float() @ b.float().t() if (a_major, b_major) == (0, 0) else (
        (a.float() if a_major == 0 else a.float().t()) @ (b.float().t() if b_major == 0 else b.float()))
    odt = torch.bfloat16 if c_bf16 else torch.float32
    out = torch.zeros(m, n, device=dev, dtype=odt)
    aux = None
    if epi == "none":
        ref = ref + bias
    elif epi == "residual":
        aux = torch.randn(m, n, device=dev, generator=g)
        ref = ref + bias + aux
    elif epi == "gelu":
        aux = torch.zeros(m, n, device=dev, dtype=odt)
        u = (ref + bias).to(odt).float().requires_grad_(True)
        ref = torch.nn.functional.gelu(u)
        ref.sum().backward()
        pre = u.grad  # the second output is gelu'(u)
        ref = ref.detach()
    elif epi == "dgelu":
        aux = torch.randn(m, n, device=dev, generator=g).to(odt)
        ref = ref * aux.float()
        bias = None
    elif epi == "accum":
        out = torch.randn(m, n, device=dev, generator=g)
        ref = ref + out
        bias = None
    def call():
        if epi == "gelu":
            ops.gemm_gelu(a, b, bias, out, aux)
        else:
            ops.gemm(a, b, bias, out, aux, a_major, b_major, ops.EPI[epi], split_k, 0)

    call()
    torch.cuda.synchronize()
    res = {"rel": _rel(out, ref), "maxabs": _maxabs(out, ref)}
    if epi == "gelu":
        res["rel_pre"] = _rel(aux, pre)
    if time_it:
        ms = _time(call)
        res["ms"] = ms
        res["tflops"] = 2.0 * m * n * k / ms / 1e9
    return res


@check
def gemm_fwd_small():
    return _gemm_case(128, 256, 64, 0, 0, "none", False)


@check
def gemm_fwd_k256():
    return _gemm_case(128, 256, 256, 0, 0, "none", False)


@check
def gemm_fwd_multi_tile():
    return _gemm_case(1000, 776, 328, 0, 0, "none", True)


@check
def gemm_fwd_vitb_qkv():
    return _gemm_case(50432, 2304, 768, 0, 0, "none", True, time_it=True)


@check
def gemm_fwd_vitb_proj_residual():
    return _gemm_case(50432, 768, 768, 0, 0, "residual", False, time_it=True)


@check
def gemm_fwd_vitb_fc1_gelu():
    return _gemm_case(50432, 3072, 768, 0, 0, "gelu", True, time_it=True)


@check
def gemm_fwd_vitb_fc2_residual():
    return _gemm_case(50432, 768, 3072, 0, 0, "residual", False, time_it=True)


@check
def gemm_dgrad_small():
    return _gemm_case(128, 256, 64, 0, 1, "none", False)


@check
def gemm_dgrad_multi():
    return _gemm_case(1000, 776, 328, 0, 1, "none", True)


@check
def gemm_dgrad_vitb_fc2_dgelu():
    return _gemm_case(50432, 3072, 768, 0, 1, "dgelu", True, time_it=True)


@check
def gemm_dgrad_vitb_fc1():
    return _gemm_case(50432, 768, 3072, 0, 1, "none", True, time_it=True)


@check
def gemm_wgrad_small():
    return _gemm_case(128, 256, 64, 1, 1, "accum", False)


@check
def gemm_wgrad_multi():
    return _gemm_case(776, 328, 1000, 1, 1, "accum", False, split_k=3)


@check
def gemm_wgrad_vitb_fc1():
    return _gemm_case(3072, 768, 50432, 1, 1, "accum", False, split_k=2, time_it=True)


@check
def gemm_wgrad_vitb_proj():
    return _gemm_case(768, 768, 50432, 1, 1, "accum", False, split_k=8, time_it=True)


@check
def gemm_a_mn_only():
    return _gemm_case(256, 256, 128, 1, 0, "none", False)


@check
def gemm_f32_modes():
    import torch
    from fedvit_b200 import ops

    out = {}
    for (am, bm) in [(0, 0), (0, 1), (1, 1), (1, 0)]:
        m, n, k = 197, 130, 77
        g = torch.Generator(device="cuda").manual_seed(5)
        A = torch.randn(m, k, device="cuda", generator=g)
        B = torch.randn(n, k, device="cuda", generator=g)
        bias = torch.randn(n, device="cuda", generator=g)
        a = A if am == 0 else A.t().contiguous()
        b = B if bm == 0 else B.t().contiguous()
        c = torch.empty(m, n, device="cuda")
        ops.gemm(a, b, bias, c, None, am, bm, 0, 1, 0)
        ref = (A.double() @ B.double().t() + bias.double()).float()
        out[f"rel_{am}{bm}"] = _rel(c, ref)
    return out


@check
def layernorm():
    import torch
    from fedvit_b200 import ops

    res = {}
    for cols in (192, 768, 1024):
        rows = 4099
        g = torch.Generator(device="cuda").manual_seed(cols)
        x = torch.randn(rows, cols, device="cuda", generator=g) * 2 + 0.5
        gam = torch.randn(cols, device="cuda", generator=g)
        bet = torch.randn(cols, device="cuda", generator=g)
        dy = torch.randn(rows, cols, device="cuda", generator=g)
        dres = torch.randn(rows, cols, device="cuda", generator=g)
        xr = x.clone().requires_grad_(True)
        gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
        yr = torch.nn.functional.layer_norm(xr, (cols,), gr, br, 1e-6)
        yr.backward(dy)
        y, mean, rstd = ops.layernorm_fwd(x, gam, bet, 1e-6, False)
        res[f"fwd_{cols}"] = _rel(y, yr)
        dg, db = torch.zeros(cols, device="cuda"), torch.zeros(cols, device="cuda")
        dx, dxlp = ops.layernorm_bwd(dy, x, gam, mean, rstd, dres, dg, db, True)
        res[f"dx_{cols}"] = _rel(dx, xr.grad + dres)
        res[f"dxlp_{cols}"] = _rel(dxlp, (xr.grad + dres).bfloat16())
        res[f"dg_{cols}"] = _rel(dg, gr.grad)
        res[f"db_{cols}"] = _rel(db, br.grad)
        ybf, _, _ = ops.layernorm_fwd(x, gam, bet, 1e-6, True)
        res[f"fwd_bf16_{cols}"] = _rel(ybf, yr.bfloat16())
    rows, cols = 50432, 768
    x = torch.randn(rows, cols, device="cuda")
    gam, bet = torch.ones(cols, device="cuda"), torch.zeros(cols, device="cuda")
    ms = _time(lambda: ops.layernorm_fwd(x, gam, bet, 1e-6, True))
    res["fwd_ms"] = ms
    res["fwd_GBs"] = rows * cols * 6 / ms / 1e6
    y, mean, rstd = ops.layernorm_fwd(x, gam, bet, 1e-6, True)
    dg, db = torch.zeros(cols, device="cuda"), torch.zeros(cols, device="cuda")
    dres = torch.randn(rows, cols, device="cuda")  # its own buffer: aliasing x would halve the fp32 reads
    ms = _time(lambda: ops.layernorm_bwd(y, x, gam, mean, rstd, dres, dg, db, True))
    res["bwd_ms"] = ms
    res["bwd_GBs"] = rows * cols * (2 + 4 + 4 + 4 + 2) / ms / 1e6
    return res


@check
def attention():
    import math
    import torch
    from fedvit_b200 import ops

    res = {}
    for (B, N, H) in [(2, 197, 3), (1, 577, 2), (3, 64, 1), (2, 65, 2), (3, 257, 2), (2, 768, 1), (5, 400, 3)]:
        g = torch.Generator(device="cuda").manual_seed(N)
        qkv = (torch.randn(B * N, 3 * H * 64, device="cuda", generator=g)).bfloat16()
        dout = torch.randn(B * N, H * 64, device="cuda", generator=g).bfloat16()
        scale = 1.0 / math.sqrt(64)
        q, k, v = (qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)[i].clone().requires_grad_(True)
                   for i in range(3))
        s = (q @ k.transpose(-1, -2)) * scale
        p = s.softmax(-1)
        o = (p @ v).transpose(1, 2).reshape(B * N, H * 64)
        o.backward(dout.float())
        out, lse = ops.attention_fwd(qkv, B, N, H, scale)
        tag = f"{B}x{N}x{H}"
        res[f"fwd_{tag}"] = _rel(out, o)
        res[f"lse_{tag}"] = _rel(lse, torch.logsumexp(s, -1))
        dqkv = ops.attention_bwd(qkv, out, dout, lse, B, N, H, scale)
        ref = torch.stack([q.grad, k.grad, v.grad], 0).permute(1, 3, 0, 2, 4).reshape(B * N, 3 * H * 64)
        d = dqkv.float().view(B * N, 3, H * 64)
        r = ref.view(B * N, 3, H * 64)
        res[f"dq_{tag}"] = _rel(d[:, 0], r[:, 0])
        res[f"dk_{tag}"] = _rel(d[:, 1], r[:, 1])
        res[f"dv_{tag}"] = _rel(d[:, 2], r[:, 2])
    B, N, H = 256, 197, 12
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    dout = torch.randn(B * N, H * 64, device="cuda").bfloat16()
    out, lse = ops.attention_fwd(qkv, B, N, H, 0.125)
    ms = _time(lambda: ops.attention_fwd(qkv, B, N, H, 0.125))
    res["fwd_ms_vitb"] = ms
    res["fwd_tflops"] = 4.0 * B * H * N * N * 64 / ms / 1e9
    ms = _time(lambda: ops.attention_bwd(qkv, out, dout, lse, B, N, H, 0.125))
    res["bwd_ms_vitb"] = ms
    res["bwd_tflops_alg"] = 10.0 * B * H * N * N * 64 / ms / 1e9
    B, N, H = 64, 577, 16  # ViT-L/16 @ 384
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    dout = torch.randn(B * N, H * 64, device="cuda").bfloat16()
    out, lse = ops.attention_fwd(qkv, B, N, H, 0.125)
    res["fwd_ms_vitl384"] = _time(lambda: ops.attention_fwd(qkv, B, N, H, 0.125))
    res["bwd_ms_vitl384"] = _time(lambda: ops.attention_bwd(qkv, out, dout, lse, B, N, H, 0.125))
    return res


@check
def losses():
    import torch
    from fedvit_b200 import ops

    res = {}

    def asl_ref(logits, targets, gn=4.0, gp=1.0, clip=0.05, eps=1e-8):
        import torch.nn.functional as F

        probs = torch.softmax(logits, dim=1)
        oh = F.one_hot(targets, logits.size(1)).float()
        p_pos = probs.clamp(min=eps)
        p_neg = probs.clamp(max=1.0 - eps)
        if clip > 0:
            p_neg = (p_neg - clip).clamp(min=eps)
        lp = oh * torch.log(p_pos)
        ln = (1.0 - oh) * torch.log(1.0 - p_neg)
        wp = (1.0 - probs).clamp(min=0.0) ** gp
        wn = probs.clamp(min=0.0) ** gn
        return (-(wp * lp + wn * ln)).sum(1).mean()

    for (B, C) in [(4, 7), (256, 7), (33, 8), (1, 3)]:
        g = torch.Generator(device="cuda").manual_seed(B)
        logits = (torch.randn(B, C, device="cuda", generator=g) * 3).requires_grad_(True)
        targets = torch.randint(0, C, (B,), device="cuda", generator=g)
        ref = asl_ref(logits, targets)
        ref.backward()
        loss, dl = ops.asl_loss(logits.detach(), targets, 4.0, 1.0, 0.05, 1e-8)
        res[f"asl_{B}x{C}"] = abs(loss.item() - ref.item()) / abs(ref.item())
        res[f"asl_grad_{B}x{C}"] = _rel(dl, logits.grad)
        logits.grad = None
        ce = torch.nn.functional.cross_entropy(logits, targets)
        ce.backward()
        loss, dl = ops.ce_loss(logits.detach(), targets)
        res[f"ce_{B}x{C}"] = abs(loss.item() - ce.item()) / abs(ce.item())
        res[f"ce_grad_{B}x{C}"] = _rel(dl, logits.grad)
    # the survey's known-answer vectors (reference losses.py on CPU)
    lg = torch.tensor([[2.0, 0.0, -1.0], [0.5, 0.5, 0.5]], device="cuda")
    loss, _ = ops.asl_loss(lg, torch.tensor([0, 2], device="cuda"), 4.0, 1.0, 0.05, 1e-8)
    res["kat1"] = loss.item()
    return res


@check
def optimizer():
    import torch
    from fedvit_b200 import ops

    res = {}
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(1)
    p = torch.randn(n, device="cuda", generator=g)
    grad = torch.randn(n, device="cuda", generator=g) * 0.01
    seg_end = torch.tensor([n // 4, n // 2, n], device="cuda", dtype=torch.int64)
    lrs, wds = [1e-3, -1.0, 3e-3], [1e-2, 0.0, 1e-5]
    seg_lr = torch.tensor(lrs, device="cuda")
    seg_wd = torch.tensor(wds, device="cuda")
    # torch reference: three param groups; the middle one is not optimised
    ps = [p[: n // 4].clone().requires_grad_(True), p[n // 4: n // 2].clone(), p[n // 2:].clone().requires_grad_(True)]
    ps[0].grad, ps[2].grad = grad[: n // 4].clone(), grad[n // 2:].clone()
    opt = torch.optim.AdamW([{"params": [ps[0]], "lr": lrs[0], "weight_decay": wds[0]},
                             {"params": [ps[2]], "lr": lrs[2], "weight_decay": wds[2]}])
    total = torch.nn.utils.clip_grad_norm_([ps[0], ps[2]], 1.0)  # norm over optimised grads only here
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    ema = p.clone()
    plp = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    ss = torch.zeros(1, device="cuda")
    gq = grad.clone()
    gq[n // 4: n // 2] = 0  # unoptimised segment carries no gradient in this check
    pp = p.clone()
    for step in (1, 2, 3):
        ops.sumsq(gq, ss, False)
        ops.adamw_flat(pp, gq, m, v, seg_end, seg_lr, seg_wd, ss, 1.0, 0.9, 0.999, 1e-8, step, ema, 0.9995, plp)
        if step > 1:
            ps[0].grad, ps[2].grad = grad[: n // 4].clone(), grad[n // 2:].clone()
            torch.nn.utils.clip_grad_norm_([ps[0], ps[2]], 1.0)
        opt.step()
    ref = torch.cat([ps[0].detach(), ps[1], ps[2].detach()])
    res["sumsq_rel"] = abs(ss.sqrt().item() - total.item()) / total.item()
    res["p_rel"] = _rel(pp, ref)
    res["p_maxabs"] = _maxabs(pp, ref)
    res["lp_rel"] = _rel(plp, ref.bfloat16())
    res["mid_untouched"] = bool(torch.equal(pp[n // 4: n // 2], p[n // 4: n // 2]))
    # fedavg fold bit-exactness against the sequential fp32 oracle order
    ws = [torch.randn(n, device="cuda", generator=g) for _ in range(5)]
    nk = [100.0, 250.0, 50.0, 300.0, 300.0]
    tot = sum(nk)
    acc = torch.empty(n, device="cuda")
    oracle = None
    for i, (w, k) in enumerate(zip(ws, nk)):
        wt = torch.tensor(k / tot, dtype=torch.float32).item()
        ops.fedavg_accum(acc, w, wt, i == 0)
        term = w * torch.tensor(wt, device="cuda", dtype=torch.float32)
        oracle = term if oracle is None else oracle + term
    res["fedavg_bit_exact"] = bool(torch.equal(acc, oracle))
    # bandwidth at ViT-B size
    n = 86_196_224
    P, G, M, V = (torch.zeros(n, device="cuda") for _ in range(4))
    se = torch.tensor([n], device="cuda", dtype=torch.int64)
    sl, sw = torch.tensor([1e-4], device="cuda"), torch.tensor([1e-5], device="cuda")
    ms = _time(lambda: ops.adamw_flat(P, G, M, V, se, sl, sw, None, 0.0, 0.9, 0.999, 1e-8, 1, None, 0.0, None))
    res["adamw_ms"] = ms
    res["adamw_GBs"] = n * 28 / ms / 1e6
    ms = _time(lambda: ops.fedavg_accum(P, G, 0.5, False))
    res["fedavg_GBs"] = n * 12 / ms / 1e6
    return res


@check
def elementwise():
    import torch
    from fedvit_b200 import ops

    res = {}
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.randn(3, 4, 224, 224, device="cuda", generator=g)
    ref = torch.nn.functional.unfold(img, 16, stride=16).transpose(1, 2).reshape(-1, 4 * 256)
    res["patchify_f32_equal"] = bool(torch.equal(ops.patchify(img, False), ref))
    res["patchify_bf16_equal"] = bool(torch.equal(ops.patchify(img, True), ref.bfloat16()))
    a = torch.randn(5000, 770, device="cuda", generator=g)
    out = torch.zeros(770, device="cuda")
    ops.colsum(a, out, False)
    res["colsum_f32"] = _rel(out, a.sum(0))
    ab = a.bfloat16()
    ops.colsum(ab, out, True)
    res["colsum_bf16_acc"] = _rel(out, a.sum(0) + ab.float().sum(0))
    s = torch.randn(37, 197, device="cuda", generator=g)
    res["softmax"] = _rel(ops.softmax_rows(s, 0.125), (s * 0.125).softmax(-1))
    return res


def _run_one(name: str) -> dict:
    import torch

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    t0 = time.time()
    out = CHECKS[name]()
    out["_s"] = round(time.time() - t0, 2)
    return out


def main() -> int:
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        print("RESULT " + json.dumps(_run_one(sys.argv[2])))
        return 0
    names = sys.argv[1:] or list(CHECKS)
    outdir = ROOT / "gpurun_out"
    outdir.mkdir(exist_ok=True)
    bad = 0
    with open(outdir / "gpu_check.log", "a") as log:
        for name in names:
            try:
                r = subprocess.run([sys.executable, __file__, "--one", name], capture_output=True,
                                   text=True, timeout=float(os.environ.get("CHECK_TIMEOUT", "240")))
                line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                if r.returncode == 0 and line:
                    rec = {"check": name, "ok": True, **json.loads(line[-1][7:])}
                else:
                    rec = {"check": name, "ok": False, "rc": r.returncode,
                           "tail": (r.stdout + r.stderr)[-1500:]}
                    bad += 1
            except subprocess.TimeoutExpired:
                rec = {"check": name, "ok": False, "timeout": True}
                bad += 1
            s = json.dumps(rec)
            print(s, flush=True)
            log.write(s + "\n")
            log.flush()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
