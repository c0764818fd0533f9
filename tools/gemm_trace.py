#!/usr/bin/env python
"""Per-tile timeline of one epilogue warp of the CTA-pair GEMM (measurement build:
`python tools/build_variants.py gemm_tc.cu gtrace:-DGEMM_TRACE`, then FEDVIT_LIB=<variant> python tools/gemm_trace.py [epilogue]).
epilogue: none | gelu | residual | dgelu (the ViT-B/16 batch-256 shapes that use it)."""
import ctypes
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fedvit_b200  # noqa: F401,E402
from fedvit_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "gelu"
M, dev, bf = 256 * 197, "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.05)
if which == "gelu":
    x, w, b = rnd(M, 768).to(bf), rnd(3072, 768).to(bf), rnd(3072)
    out, aux = torch.empty(M, 3072, device=dev, dtype=bf), torch.empty(M, 3072, device=dev, dtype=bf)
    fn = lambda: ops.gemm_gelu(x, w, b, out, aux)
elif which == "none":
    x, w, b = rnd(M, 768).to(bf), rnd(2304, 768).to(bf), rnd(2304)
    out = torch.empty(M, 2304, device=dev, dtype=bf)
    fn = lambda: ops.gemm(x, w, b, out, None, 0, 0, ops.EPI["none"], 1, 0)
elif which == "residual":
    x, w, b, r = rnd(M, 768).to(bf), rnd(768, 768).to(bf), rnd(768), rnd(M, 768)
    out = torch.empty(M, 768, device=dev)
    fn = lambda: ops.linear_residual(x, w, b, r, None, 0, out)
else:
    raise SystemExit("epilogue: none | gelu | residual")
for _ in range(3):
    fn()
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["FEDVIT_LIB"])
buf = (ctypes.c_longlong * (24 * 8))()
assert lib.fv_debug_read_gemm_trace(buf) == 0
base = buf[0]
print(f"epilogue warp 4 of CTA 0, {which}: cycles")
for t in range(24):
    ev = [buf[t * 8 + e] for e in range(7)]
    nxt = buf[(t + 1) * 8] if t + 1 < 24 else ev[2]
    print(f" tile {t:2d}: starts {ev[0] - base:8d} | accumulator wait +{ev[1] - ev[0]:6d} | drain +{ev[2] - ev[1]:6d} (of it: staging-tile waits {ev[3]:6d}, tcgen05.ld waits {ev[4]:5d}, arithmetic incl. the ld wait {ev[5]:6d}, stage_store {ev[6]:5d})")
