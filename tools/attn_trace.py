#!/usr/bin/env python
"""Per-tile timeline of the attention forward kernel from SM-clock timestamps (measurement build:
`python tools/build_variants.py attention_tc.cu trace:-DATC_TRACE`, then FEDVIT_LIB=<variant> python tools/attn_trace.py)."""
import ctypes
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fedvit_b200  # noqa: F401,E402
from fedvit_b200 import ops  # noqa: E402

VER = os.environ.setdefault("FEDVIT_ATTN_FWD", "v1")
B, N, H = 256, 197, 12
g = torch.Generator(device="cuda").manual_seed(1)
qkv = torch.randn(B * N, 3 * H * 64, device="cuda", generator=g).bfloat16()
for _ in range(3):
    ops.attention_fwd(qkv, B, N, H, 0.125)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["FEDVIT_LIB"])
buf = (ctypes.c_longlong * (4 * 32 * 8))()
assert (lib.fv_debug_read_trace3 if VER == "v3" else lib.fv_debug_read_trace)(buf, 4 * 32 * 8) == 0
names = ["S ready", "pass1 done", "P written", "O ready", "stored", "mma: S issue", "mma: PV issue"]
for c in range(4):
    base = buf[(c * 32 + 0) * 8 + 5]
    print(f"--- CTA slot {c} (smid {buf[(c * 32) * 8 + 7]}), cycles relative to its first S issue")
    for gt in range(12):
        ev = [buf[(c * 32 + gt) * 8 + e] - base for e in range(7)]
        if VER == "v3":  # the score MMA is issued during the previous tile: absolute times, period = P-written to P-written
            print(f" tile {gt:2d}: S issue {ev[5]:7d} | S ready {ev[0]:7d} | pass1 done {ev[1]:7d} (+{ev[1] - ev[0]:5d}) | P written {ev[2]:7d} (+{ev[2] - ev[1]:5d})"
                  f" | last PV issue {ev[6]:7d} | O ready {ev[3]:7d} (+{ev[3] - ev[2]:5d}) | stored {ev[4]:7d} (+{ev[4] - ev[3]:5d})")
            continue
        s_issue, pv_issue = ev[5], ev[6]
        print(f" tile {gt:2d}: S issue {s_issue:7d} | S ready +{ev[0] - s_issue:5d} | pass1 +{ev[1] - ev[0]:5d} | pass2 +{ev[2] - ev[1]:5d}"
              f" | PV issue +{pv_issue - ev[2]:5d} | O ready +{ev[3] - pv_issue:5d} | store +{ev[4] - ev[3]:5d} | tile total {ev[4] - s_issue:6d}")
