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
assert {"v3": "fv_debug_read_trace3", "v4": "fv_debug_read_trace4"}.get(VER, "fv_debug_read_trace") and getattr(lib, {"v3": "fv_debug_read_trace3", "v4": "fv_debug_read_trace4"}.get(VER, "fv_debug_read_trace"))(buf, 4 * 32 * 8) == 0
names = ["S ready", "pass1 done", "P written", "O ready", "stored", "mma: S issue", "mma: PV issue"]
for c in range(4):
    base = buf[(c * 32 + 0) * 8 + 5]
    print(f"--- CTA slot {c} (smid {buf[(c * 32) * 8 + 7]}), cycles relative to its first S issue")
    for gt in range(12):
        ev = [buf[(c * 32 + gt) * 8 + e] - base for e in range(7)]
        if VER == "v4":  # softmax warp 0: 0 arrives at the score wait, 1 scores there, 2 P written; epilogue warp 4: 3 O ready, 4 stored
            print(f" tile {gt:2d}: S issue {ev[5]:7d} | softmax waits from {ev[0]:7d} | pass {ev[1]:7d} .. {ev[2]:7d} ({ev[2] - ev[1]:5d})"
                  f" | PV issue {ev[6]:7d} | O ready {ev[3]:7d} (+{ev[3] - ev[6]:5d}) | stored {ev[4]:7d} (+{ev[4] - ev[3]:5d})")
            continue
        if VER == "v3":  # the score MMA is issued during the previous tile: absolute times, period = P-written to P-written
            print(f" tile {gt:2d}: S issue {ev[5]:7d} | S ready {ev[0]:7d} | pass1 done {ev[1]:7d} (+{ev[1] - ev[0]:5d}) | P written {ev[2]:7d} (+{ev[2] - ev[1]:5d})"
                  f" | last PV issue {ev[6]:7d} | O ready {ev[3]:7d} (+{ev[3] - ev[2]:5d}) | stored {ev[4]:7d} (+{ev[4] - ev[3]:5d})")
            continue
        s_issue, pv_issue = ev[5], ev[6]
        print(f" tile {gt:2d}: S issue {s_issue:7d} | S ready +{ev[0] - s_issue:5d} | pass1 +{ev[1] - ev[0]:5d} | pass2 +{ev[2] - ev[1]:5d}"
              f" | PV issue +{pv_issue - ev[2]:5d} | O ready +{ev[3] - pv_issue:5d} | store +{ev[4] - ev[3]:5d} | tile total {ev[4] - s_issue:6d}")

if VER == "v4":
    for c in range(4):
        k = [buf[(c * 32 + 31) * 8 + e] for e in range(4)]
        base = buf[(c * 32 + 0) * 8 + 5]
        print(f"--- CTA slot {c}: kernel entry {k[0] - base} .. exit {k[1] - base} cycles (relative to the first S issue) = {k[1] - k[0]} cycles"
              f" in {k[3] - k[2]} ns -> SM clock {(k[1] - k[0]) / max(k[3] - k[2], 1):.3f} GHz")
    cb = (ctypes.c_longlong * 40)()
    assert lib.fv_debug_read_chunks4(cb) == 0
    print("--- softmax warp 0 of CTA 0, tile 4, per 32-key chunk: cycles since the previous stamp")
    prev = cb[0]
    for c in range(6):
        ev = [cb[c * 5 + e] for e in range(4)]
        print(f" chunk {c}: load wait +{ev[0] - prev:5d} | next load issue +{ev[1] - ev[0]:5d} | exponentials +{ev[2] - ev[1]:5d} | P store +{ev[3] - ev[2]:5d}")
        prev = ev[3]
