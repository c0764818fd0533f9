#!/usr/bin/env python
"""Attention backward (ViT-B/16 batch 256: 3072 items of 197 tokens) with parts of the kernel switched
off (FEDVIT_ATTN_DBG, results wrong by construction) — where does its time go?

    python tools/attn_probe.py > gpurun_out/attn_probe.txt
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch

    import fedvit_b200  # noqa: F401
    from fedvit_b200 import ops
    from tools.gpu_check import _time

    B, N, H = 256, 197, 12
    if len(sys.argv) > 1:
        B, N, H = (int(v) for v in sys.argv[1:4])
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda", generator=g).bfloat16()
    dout = torch.randn(B * N, H * 64, device="cuda", generator=g).bfloat16()
    scale = 0.125
    out, lse = ops.attention_fwd(qkv, B, N, H, scale)
    names = {0: "full kernel", 1: "no softmax / dS arithmetic", 2: "no dV / dK / dQ MMAs", 4: "no drain stores",
             8: "no helper loads", 16: "no P / dS shared-memory stores", 17: "no arithmetic, no P / dS stores",
             19: "scores only (1 + 2 + 16)", 31: "skeleton (everything off)", 12: "no drain stores, no helper loads"}
    for dbg, name in names.items():
        os.environ["FEDVIT_ATTN_DBG"] = str(dbg)
        ms = _time(lambda: ops.attention_bwd(qkv, out, dout, lse, B, N, H, scale))
        print(f"dbg {dbg:2d}  {ms * 1e3:7.1f} us  {name}")
    os.environ.pop("FEDVIT_ATTN_DBG")


if __name__ == "__main__":
    main()
