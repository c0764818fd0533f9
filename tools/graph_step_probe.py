#!/usr/bin/env python
"""Eager training step against graphs.GraphedTrainStep (one CUDA-graph replay per step):
ViT-B/16 batch 256 (GPU-bound) and ViT-Tiny/16 batch 16 (BASELINE config 1, launch-bound).

    python tools/graph_step_probe.py > gpurun_out/graph_step_probe.txt
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    import torch

    import bench
    import fedvit_b200  # noqa: F401
    from fedvit_b200 import fedavg, graphs, losses, model, optim, utils
    from fedvit_b200.arena import FlatArena

    dev = torch.device("cuda", 0)
    for backbone, batch, steps in (("vit_base_patch16_224", 256, 10), ("vit_tiny_patch16_224", 16, 50)):
        cfg = bench.model_config()
        cfg["model"]["backbone"] = backbone
        utils.seed_everything(42)
        net = model.build_model(cfg).to(dev).train()
        arena = FlatArena(net)
        fedavg.broadcast_initial(arena, net)
        opt = optim.FusedAdamW(model.get_layerwise_lr_groups(net, 1e-4, 0.75, 1e-5), weight_decay=1e-5, arena=arena)
        crit = losses.build_loss(cfg)
        x = torch.randn(4 * batch, 3, 224, 224, device=dev)
        y = torch.randint(0, bench.CLASSES, (4 * batch,), device=dev)

        def eager(i):
            j = (i % 4) * batch
            opt.zero_grad(set_to_none=True)
            with torch.amp.autocast("cuda", dtype=torch.bfloat16):
                loss = crit(net(x[j:j + batch])["logits"], y[j:j + batch])
            loss.backward()
            utils.clip_grad_norm(net.parameters(), 1.0, optimizer=opt)
            opt.step()

        def timed(fn):
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

        t_eager = timed(eager)
        step = graphs.GraphedTrainStep(net, crit, opt, x[:batch], y[:batch], grad_clip=1.0)

        def graphed(i):
            j = (i % 4) * batch
            step(x[j:j + batch], y[j:j + batch])

        t_graph = timed(graphed)
        print(f"{backbone} batch {batch}: eager {t_eager:.3f} ms/step ({batch / t_eager * 1e3:.0f} images/s), "
              f"graph replay {t_graph:.3f} ms/step ({batch / t_graph * 1e3:.0f} images/s)")
        del step, net, opt, arena
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
