#!/usr/bin/env python
"""The fused AdamW sweep timed alone on a B200 (ViT-B/16 arena: 86.2 M parameters), with and without the fused
gradient zeroing / bf16 re-cast / EMA, cold L2.   python tools/adamw_probe.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fedvit_b200  # noqa: F401,E402
from fedvit_b200 import ops  # noqa: E402

dev = "cuda"
n = 86_196_032
P, G, M, V, E = (torch.randn(n, device=dev) * 0.01 for _ in range(5))
V.abs_()
LP = torch.empty(n, device=dev, dtype=torch.bfloat16)
nseg = 17
ends = torch.tensor([n * (i + 1) // nseg // 64 * 64 if i + 1 < nseg else n for i in range(nseg)], device=dev, dtype=torch.int64)
lrs = torch.full((nseg,), 1e-4, device=dev)
wds = torch.full((nseg,), 1e-5, device=dev)
ss = torch.ones(1, device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


for zero in (False, True):
    for lp in (None, LP):
        for ema in (None, E):
            us = timeit(lambda: ops.adamw_flat(P, G, M, V, ends, lrs, wds, ss, 1.0, 0.9, 0.999, 1e-8, 5, ema, 0.9995, lp, zero))
            byts = n * (28 + (4 if zero else 0) + (2 if lp is not None else 0) + (8 if ema is not None else 0))
            print(f"zero_grad={zero!s:5} lp={'yes' if lp is not None else 'no ':3} ema={'yes' if ema is not None else 'no ':3}: {us:8.1f} us  {byts / us / 1e6:6.2f} TB/s")
print(f"sumsq: {timeit(lambda: ops.sumsq(G, ss, False)):8.1f} us")
