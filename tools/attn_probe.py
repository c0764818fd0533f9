#!/usr/bin/env python
"""Attention kernels on a B200: parity of forward / backward against an fp64 reference over ragged shapes,
the second-generation backward (keys on lanes) against the first (FEDVIT_ATTN_BWD=v1), and CUDA-event
timings at the benchmark shape (256 x 197 x 12 heads) and the ViT-L shape.

    python tools/attn_probe.py [--iters 20]
"""
import argparse
import math
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fedvit_b200  # noqa: F401,E402
from fedvit_b200 import ops  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def reference(qkv, dout, B, N, H, scale):
    q, k, v = (qkv.double().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)[i].clone().requires_grad_(True) for i in range(3))
    s = (q @ k.transpose(-1, -2)) * scale
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * N, H * 64)
    o.backward(dout.double())
    ref = torch.stack([q.grad, k.grad, v.grad], 0).permute(1, 3, 0, 2, 4).reshape(B * N, 3, H * 64)
    return o, torch.logsumexp(s, -1), ref


def check(B, N, H, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed + N)
    qkv = torch.randn(B * N, 3 * H * 64, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B * N, H * 64, device=DEV, generator=g).bfloat16()
    scale = 1.0 / math.sqrt(64)
    o, lse_ref, ref = reference(qkv, dout, B, N, H, scale)
    os.environ["FEDVIT_ATTN_FWD"] = "v1"
    out1, lse1 = ops.attention_fwd(qkv, B, N, H, scale)
    os.environ["FEDVIT_ATTN_FWD"] = "v2"
    out, lse = ops.attention_fwd(qkv, B, N, H, scale)
    out_again, _ = ops.attention_fwd(qkv, B, N, H, scale)
    res = {"fwd1": rel(out1, o), "lse1": rel(lse1, lse_ref), "fwd": rel(out, o), "lse": rel(lse, lse_ref),
           "fwd.repro": bool(torch.equal(out, out_again))}
    for ver in ("v2", "v1"):
        os.environ["FEDVIT_ATTN_BWD"] = ver
        d = ops.attention_bwd(qkv, out, dout, lse, B, N, H, scale).float().view(B * N, 3, H * 64)
        torch.cuda.synchronize()
        floor = 1e-3 * (float(ref.norm()) + 1e-6 * float(dout.double().norm()))  # N == 1: dq, dk are exactly 0
        for i, name in enumerate("qkv"):
            den = max(float(ref[:, i].norm()), floor)
            res[f"{ver}.d{name}"] = float((d[:, i].double().cpu() - ref[:, i].cpu()).norm()) / den
        again = ops.attention_bwd(qkv, out, dout, lse, B, N, H, scale).float().view(B * N, 3, H * 64)
        res[f"{ver}.repro"] = bool(torch.equal(again, d))
    os.environ["FEDVIT_ATTN_BWD"] = "v2"
    bad = [k for k, v in res.items() if (isinstance(v, float) and not (v < 1e-2)) or v is False]
    print(f"B={B} N={N} H={H}: " + " ".join(f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}" for k, v in res.items()),
          "  <-- BAD " + ",".join(bad) if bad else "", flush=True)
    return not bad


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()  # 256 MB > L2: cold operands every iteration
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--skip-check", action="store_true")
    ap.add_argument("--one", action="store_true", help="time the benchmark shape only (ncu captures)")
    args = ap.parse_args()
    ok = True
    if not args.skip_check:
        for shape in [(1, 1, 1), (3, 64, 1), (2, 65, 2), (2, 128, 2), (4, 129, 2), (2, 197, 3), (2, 256, 2), (3, 200, 1),
                      (40, 197, 12), (300, 197, 3)]:
            ok &= check(*shape)
    scale = 0.125
    for (B, N, H) in ([(256, 197, 12)] if args.one else [(256, 197, 12), (64, 197, 12), (1024, 197, 3)]):
        g = torch.Generator(device=DEV).manual_seed(1)
        qkv = torch.randn(B * N, 3 * H * 64, device=DEV, generator=g).bfloat16()
        dout = torch.randn(B * N, H * 64, device=DEV, generator=g).bfloat16()
        out, lse = ops.attention_fwd(qkv, B, N, H, scale)
        fl = 4.0 * N * N * 64 * B * H
        line = f"B={B} N={N} H={H}:"
        for ver in ("v1", "v2"):
            os.environ["FEDVIT_ATTN_FWD"] = ver
            t_f = timeit(lambda: ops.attention_fwd(qkv, B, N, H, scale), args.iters)
            line += f" fwd {ver} {t_f:.1f} us ({fl / t_f / 1e6:.0f} TF/s) |"
        for ver in ("v1", "v2"):
            os.environ["FEDVIT_ATTN_BWD"] = ver
            t = timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, B, N, H, scale), args.iters)
            line += f" bwd {ver} {t:.1f} us ({2.5 * fl / t / 1e6:.0f} TF/s) |"
        print(line, flush=True)
    os.environ["FEDVIT_ATTN_BWD"] = "v2"
    print("PARITY", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
