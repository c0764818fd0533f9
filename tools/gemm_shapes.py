#!/usr/bin/env python
"""Every tensor-core GEMM of a ViT-B/16 training step (batch 256) timed alone, back to back, next
to cuBLAS (torch.matmul, bf16) on the same shape — tells whether a gap is the kernel's or the shape's.

    python tools/gemm_shapes.py > gpurun_out/gemm_shapes.txt
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    import torch

    import fedvit_b200  # noqa: F401
    from fedvit_b200 import ops
    from tools.gpu_check import _time

    M = 256 * 197
    dev = "cuda"
    bf = torch.bfloat16
    rows = []

    def rand(*s, dt=bf):
        return (torch.randn(*s, device=dev) * 0.5).to(dt)

    def fwd(name, n, k, epi):
        a, w, bias = rand(M, k), rand(n, k), rand(n, dt=torch.float32)
        if epi == "residual":
            out, aux = torch.empty(M, n, device=dev), rand(M, n, dt=torch.float32)
            f = lambda: ops.linear_residual(a, w, bias, aux, None, 0, out)
        elif epi == "gelu":
            out, aux = torch.empty(M, n, device=dev, dtype=bf), torch.empty(M, n, device=dev, dtype=bf)
            f = lambda: ops.gemm_gelu(a, w, bias, out, aux)
        else:
            out = torch.empty(M, n, device=dev, dtype=bf)
            f = lambda: ops.gemm(a, w, bias, out, None, 0, 0, ops.EPI["none"], 1, 0)
        ref = lambda: torch.matmul(a, w.t())
        rows.append((name, M, n, k, _time(f), _time(ref)))

    def dgrad(name, n, k, epi):  # dX[M, n] = dY[M, k] W[k, n]   (W stored [k, n]: MN-major B)
        dy, w = rand(M, k), rand(k, n)
        out = torch.empty(M, n, device=dev, dtype=bf)
        if epi == "dgelu":
            aux = rand(M, n)
            f = lambda: ops.gemm(dy, w, None, out, aux, 0, 1, ops.EPI["dgelu"], 1, 0)
        else:
            f = lambda: ops.gemm(dy, w, None, out, None, 0, 1, ops.EPI["none"], 1, 0)
        ref = lambda: torch.matmul(dy, w)
        rows.append((name, M, n, k, _time(f), _time(ref)))

    def wgrad(name, o, i, with_bias=True):
        from fedvit_b200.vit import _split_k_for

        dy, x = rand(M, o), rand(M, i)
        dw, db = torch.zeros(o, i, device=dev), (torch.zeros(o, device=dev) if with_bias else None)
        s = _split_k_for(o, i, M)
        f = lambda: ops.wgrad(dy, x, dw, db, s)
        ref = lambda: torch.matmul(dy.t(), x)
        rows.append((f"{name} split{s}", o, i, M, _time(f), _time(ref)))

    fwd("qkv fwd [none]", 2304, 768, "none")
    fwd("proj fwd [residual]", 768, 768, "residual")
    fwd("fc1 fwd [gelu]", 3072, 768, "gelu")
    fwd("fc2 fwd [residual]", 768, 3072, "residual")
    dgrad("fc2 dgrad [dgelu]", 3072, 768, "dgelu")
    dgrad("fc1 dgrad [none]", 768, 3072, "none")
    dgrad("proj dgrad [none]", 768, 768, "none")
    dgrad("qkv dgrad [none]", 768, 2304, "none")
    wgrad("qkv wgrad", 2304, 768)
    wgrad("proj wgrad", 768, 768)
    wgrad("fc1 wgrad", 3072, 768)
    wgrad("fc2 wgrad", 768, 3072)
    wgrad("fc1 wgrad, no dbias", 3072, 768, False)
    wgrad("fc2 wgrad, no dbias", 768, 3072, False)
    print(f"{'GEMM':28} {'m':>6} {'n':>5} {'k':>6} {'ours us':>9} {'TF/s':>6} {'cuBLAS us':>10} {'TF/s':>6}")
    tot_o = tot_c = 0.0
    for name, m, n, k, ms, ms_ref in rows:
        fl = 2.0 * m * n * k
        tot_o += ms
        tot_c += ms_ref
        print(f"{name:28} {m:6d} {n:5d} {k:6d} {ms * 1e3:9.1f} {fl / ms / 1e9:6.0f} {ms_ref * 1e3:10.1f} {fl / ms_ref / 1e9:6.0f}")
    print(f"sum per layer: ours {tot_o * 1e3:.0f} us, cuBLAS (no epilogue work at all) {tot_c * 1e3:.0f} us")


if __name__ == "__main__":
    main()
