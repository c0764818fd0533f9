#!/usr/bin/env python
"""Live per-kernel breakdown of one ViT-B/16 training step (bench.py's step): every C-ABI call is
bracketed by CUDA events inside normally running steps (no profiler, warm caches, real clocks).

    python tools/step_breakdown.py [--steps 3] [--batch 256] > gpurun_out/step_breakdown.txt
"""
from __future__ import annotations

import argparse
import collections
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

EPI = {0: "none", 1: "residual", 2: "gelu", 3: "dgelu", 4: "accum", 5: "patch"}


def tag(name, a):
    if name == "fv_gemm_bf16":
        m, n, k, epi = a[12], a[13], a[14], a[15]
        return f"gemm[{EPI.get(epi, epi)}] {m}x{n}x{k} a{a[1]}b{a[4]}", 2.0 * m * n * k
    if name == "fv_linear_residual_bf16":
        m, n, k = a[11], a[12], a[13]
        return f"gemm[residual] {m}x{n}x{k}", 2.0 * m * n * k
    if name == "fv_wgrad_bf16":
        t, o, i = a[7], a[8], a[9]
        return f"wgrad {o}x{i}x{t} split{a[10]}", 2.0 * t * o * i
    if name == "fv_patch_embed_tf32":
        return "patch_embed", 2.0 * a[5] * (a[7] // 16) * (a[8] // 16) * a[6] * 256 * a[9]
    return name, 0.0


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--backbone", type=str, default="vit_base_patch16_224")
    args = ap.parse_args()
    import torch

    import bench
    import fedvit_b200  # noqa: F401
    from fedvit_b200 import _lib, fedavg, losses, model, optim, utils
    from fedvit_b200.arena import FlatArena

    dev = torch.device("cuda", 0)
    cfg = bench.model_config()
    cfg["model"]["backbone"] = args.backbone
    img = int(args.backbone.split("_")[-1])
    cfg["model"]["image_size"] = img
    utils.seed_everything(42)
    net = model.build_model(cfg).to(dev).train()
    arena = FlatArena(net)
    fedavg.broadcast_initial(arena, net)
    opt = optim.FusedAdamW(model.get_layerwise_lr_groups(net, 1e-4, 0.75, 1e-5), weight_decay=1e-5, arena=arena)
    crit = losses.build_loss(cfg)
    B = args.batch
    x = torch.randn(2 * B, 3, img, img, device=dev)
    y = torch.randint(0, bench.CLASSES, (2 * B,), device=dev)

    def step(i):
        j = (i % 2) * B
        opt.zero_grad(set_to_none=True)
        with torch.amp.autocast("cuda", dtype=torch.bfloat16):
            loss = crit(net(x[j:j + B])["logits"], y[j:j + B])
        loss.backward()
        utils.clip_grad_norm(net.parameters(), 1.0, optimizer=opt)
        opt.step()

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.TRACE = []
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    trace, _lib.TRACE = _lib.TRACE, None
    total = e0.elapsed_time(e1) / args.steps
    agg = collections.OrderedDict()
    for name, a, s, e in trace:
        t, fl = tag(name, a)
        d = agg.setdefault(t, [0.0, 0, 0.0])
        d[0] += s.elapsed_time(e)
        d[1] += 1
        d[2] += fl
    print(f"{args.backbone}: step {total:.3f} ms (with per-launch events), batch {B}; per step:")
    print(f"{'ms/step':>9} {'share':>6} {'n':>4} {'us/launch':>10} {'TF/s':>7}  kernel")
    acc = 0.0
    for t, (ms, n, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        ms_s = ms / args.steps
        acc += ms_s
        tf = f"{fl / ms / 1e9:7.0f}" if fl else "       "
        print(f"{ms_s:9.3f} {ms_s / total * 100:5.1f}% {n // args.steps:4d} {ms / n * 1e3:10.1f} {tf}  {t}")
    print(f"{acc:9.3f} {acc / total * 100:5.1f}%  sum of traced libfedvit launches; the rest is torch glue + gaps")


if __name__ == "__main__":
    main()
