#!/usr/bin/env python
"""Per-cell timeline of the attention backward kernel's math warp 0 (measurement build:
`python tools/build_variants.py attention_tc.cu trace1:-DATC_TRACE`, then FEDVIT_LIB=<variant> python tools/attn_trace_bwd.py)."""
import ctypes
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fedvit_b200  # noqa: F401,E402
from fedvit_b200 import ops  # noqa: E402

B, N, H = 256, 197, 12
g = torch.Generator(device="cuda").manual_seed(1)
qkv = torch.randn(B * N, 3 * H * 64, device="cuda", generator=g).bfloat16()
dout = torch.randn(B * N, H * 64, device="cuda", generator=g).bfloat16()
out, lse = ops.attention_fwd(qkv, B, N, H, 0.125)
for _ in range(3):
    ops.attention_bwd(qkv, out, dout, lse, B, N, H, 0.125)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["FEDVIT_LIB"])
buf = (ctypes.c_longlong * (24 * 8))()
assert lib.fv_debug_read_trace_bwd(buf) == 0
base = buf[0]
print("math warp 0 of CTA 0: cell = (item, key block, query tile); cycles")
for it in range(24):
    ev = [buf[it * 8 + e] - base for e in range(7)]
    print(f" cell {it:2d} (kb {(it >> 1) & 1}, qt {it & 1}): starts {ev[0]:7d} | S wait +{ev[1] - ev[0]:5d} | loads (both chunks; chunk-0 math between) +{ev[2] - ev[1]:5d}"
          f" | rest of math +{ev[3] - ev[2]:5d} | MMA-2 wait +{ev[4] - ev[3]:5d} | P/dS stores +{ev[5] - ev[4]:5d} | drains +{ev[6] - ev[5]:5d} | cell {ev[6] - ev[0]:6d}")
