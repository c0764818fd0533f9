#!/usr/bin/env python
"""Can the HBM-bound LayerNorm backward share the GPU with a tensor-bound weight-gradient GEMM?

Measures, at the ViT-B/16 batch-256 shapes (50432 rows x 768):
  1. the bulk-copy ring LayerNorm backward against the register-pipelined one (results + time),
  2. both kernels on a capped number of SMs (how much bandwidth S SMs can pull),
  3. the fc1 / qkv weight-gradient GEMMs on a capped number of SMs,
  4. GEMM on G SMs and LayerNorm backward on the other 148 - G, launched on two streams.

    python tools/overlap_probe.py > gpurun_out/overlap_probe.txt

Parts 3 and 4 cap the grid of the single-CTA GEMM kernel (FEDVIT_GEMM_GRID): run them with
FEDVIT_GEMM_DBG=256, which keeps the weight gradients off the CTA-pair kernel they use by default now.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    import torch

    import fedvit_b200  # noqa: F401
    from fedvit_b200 import ops

    os.environ["FEDVIT_LN_REREAD"] = "1"
    dev = "cuda"
    M, D = 256 * 197, 768
    bf = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(M, D, device=dev, generator=g)
    dy = (torch.randn(M, D, device=dev, generator=g) * 0.1).to(bf)
    dres = torch.randn(M, D, device=dev, generator=g) * 0.1
    gamma = torch.rand(D, device=dev, generator=g) + 0.5
    mean = x.mean(1)
    rstd = (x.var(1, unbiased=False) + 1e-6).rsqrt()
    scale = torch.rand(256, device=dev, generator=g) + 0.5

    def ln(mode, grid=0, sc=None):
        os.environ["FEDVIT_LN_MINB"] = str(mode)
        if grid:
            os.environ["FEDVIT_LN_GRID"] = str(grid)
        else:
            os.environ.pop("FEDVIT_LN_GRID", None)
        dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
        dx, dxlp = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dres, dg, db, True, sc, 197 if sc is not None else 0)
        return dx, dxlp, dg, db

    def timed(fn, iters=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3  # us

    # 1. results
    for sc in (None, scale):
        a = ln(0, 0, sc)
        b = ln(9, 0, sc)
        c = ln(9, 37, sc)
        for name, u, v, w in zip(("dx", "dx_lp", "dgamma", "dbeta"), a, b, c):
            e1 = (u.float() - v.float()).abs().max().item()
            e2 = (u.float() - w.float()).abs().max().item()
            print(f"ring vs pipe {name:7} scale={'yes' if sc is not None else 'no '} max|diff| {e1:.3e} (148 SMs) {e2:.3e} (37 SMs)"
                  f"  ref max {u.float().abs().max().item():.3e}")
    bytes_ln = M * D * 16
    print("\nLayerNorm backward alone: us (TB/s) by SM cap")
    sweep = (0, 136, 128, 120, 112, 108, 104, 100, 96, 88, 80) if "--ln-sweep" in sys.argv else (0, 120, 100, 74, 56, 48, 40, 36, 32, 24)
    for S in sweep:
        t0 = timed(lambda: ln(0, S))
        t9 = timed(lambda: ln(9, S))
        print(f"  SMs {S or 148:4d}   pipe {t0:7.1f} ({bytes_ln / t0 / 1e6:5.2f})   ring {t9:7.1f} ({bytes_ln / t9 / 1e6:5.2f})")

    if "--ln-sweep" in sys.argv:
        return
    # 3. weight-gradient GEMMs on capped grids
    def mk_wgrad(o, i):
        dyw = (torch.randn(M, o, device=dev, generator=g) * 0.5).to(bf)
        xw = (torch.randn(M, i, device=dev, generator=g) * 0.5).to(bf)
        dw, dbias = torch.zeros(o, i, device=dev), torch.zeros(o, device=dev)

        def run(G, split):
            if G:
                os.environ["FEDVIT_GEMM_GRID"] = str(G)
            else:
                os.environ.pop("FEDVIT_GEMM_GRID", None)
            ops.wgrad(dyw, xw, dw, dbias, split)
            os.environ.pop("FEDVIT_GEMM_GRID", None)

        return run

    shapes = {"fc1 3072x768": (mk_wgrad(3072, 768), [(0, 2), (120, 5), (112, 3), (108, 3), (104, 3), (96, 4), (90, 5)]),
              "qkv 2304x768": (mk_wgrad(2304, 768), [(0, 8), (120, 9), (112, 4), (108, 4), (108, 2), (96, 16), (90, 5)])}
    for name, (run, cases) in shapes.items():
        print(f"\nwgrad {name} alone: us by (SM cap, split-K)")
        for G, s in cases:
            t = timed(lambda: run(G, s))
            print(f"  SMs {G or 148:4d} split {s:2d}  {t:7.1f}")

    # 4. concurrent: GEMM on G SMs (launched first) + LayerNorm backward on the rest
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

    def both(run, G, s, mode, S):
        ev = torch.cuda.Event()
        ev.record(sa)
        sb.wait_event(ev)
        with torch.cuda.stream(sa):
            run(G, s)
        with torch.cuda.stream(sb):
            ln(mode, S)
        ev2 = torch.cuda.Event()
        ev2.record(sb)
        sa.wait_event(ev2)

    def timed_both(run, G, s, mode, S, iters=20):
        torch.cuda.synchronize()
        with torch.cuda.stream(sa):
            for _ in range(3):
                both(run, G, s, mode, S)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(sa)
            for _ in range(iters):
                both(run, G, s, mode, S)
            e1.record(sa)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    for name, (run, _) in shapes.items():
        seq = timed(lambda: (run(0, 2 if name.startswith("fc1") else 8), ln(0, 0)))
        seq9 = timed(lambda: (run(0, 2 if name.startswith("fc1") else 8), ln(9, 0)))
        print(f"\nwgrad {name} + LayerNorm backward: back to back {seq:7.1f} us (pipe) {seq9:7.1f} us (ring); two streams:")
        for G, s in ((120, 5 if name.startswith("fc1") else 9), (112, 3 if name.startswith("fc1") else 4),
                     (108, 3 if name.startswith("fc1") else 4), (104, 3 if name.startswith("fc1") else 4),
                     (96, 4 if name.startswith("fc1") else 16)):
            for mode in (9, 0):
                t = timed_both(run, G, s, mode, 148 - G)
                print(f"  GEMM on {G:3d} SMs (split {s:2d}) + LN {'ring' if mode == 9 else 'pipe'} on {148 - G:2d} SMs: {t:7.1f} us")


if __name__ == "__main__":
    main()
