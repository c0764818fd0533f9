#!/usr/bin/env python
"""Where the end-to-end path (train_one_epoch from pinned host batches) loses time against the
device-resident step loop: same model / optimiser as bench.py, four loader / sync variants."""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    import torch

    import bench
    import fedvit_b200  # noqa: F401
    from fedvit_b200 import fedavg, losses, model, optim, train, utils
    from fedvit_b200.arena import FlatArena

    dev = torch.device("cuda", 0)
    cfg = bench.model_config()
    utils.seed_everything(42)
    net = model.build_model(cfg).to(dev).train()
    arena = FlatArena(net)
    fedavg.broadcast_initial(arena, net)
    opt = optim.FusedAdamW(model.get_layerwise_lr_groups(net, 1e-4, 0.75, 1e-5), weight_decay=1e-5, arena=arena)
    crit = losses.build_loss(cfg)
    B, pool, steps = bench.BATCH, 4, 10
    host_x = torch.randn(pool * B, 3, bench.IMG, bench.IMG).pin_memory()
    host_y = torch.randint(0, bench.CLASSES, (pool * B,)).pin_memory()
    dev_x, dev_y = host_x.to(dev), host_y.to(dev)

    class Loader:
        def __init__(self, n, on_device):
            self.n, self.dev = n, on_device

        def __len__(self):
            return self.n

        def __iter__(self):
            x, y = (dev_x, dev_y) if self.dev else (host_x, host_y)
            for i in range(self.n):
                j = (i % pool) * B
                yield {"image": x[j:j + B], "label": y[j:j + B]}

    def run(on_device, sync):
        c = bench.model_config()
        c["training"]["sync_loss_every_step"] = sync
        train.train_one_epoch(net, Loader(3, on_device), crit, opt, None, None, None, dev, c, 0, None)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        train.train_one_epoch(net, Loader(steps, on_device), crit, opt, None, None, None, dev, c, 1, None)
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / steps

    def plain():
        def step(i):
            j = (i % pool) * B
            opt.zero_grad(set_to_none=True)
            with torch.amp.autocast("cuda", dtype=torch.bfloat16):
                loss = crit(net(dev_x[j:j + B])["logits"], dev_y[j:j + B])
            loss.backward()
            utils.clip_grad_norm(net.parameters(), 1.0, optimizer=opt)
            opt.step()

        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        import time

        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        c0 = time.perf_counter()
        for i in range(steps):
            step(i)
        cpu_ms = (time.perf_counter() - c0) * 1e3 / steps  # host time to ENQUEUE a step (no sync inside)
        t1.record()
        torch.cuda.synchronize()
        print(f"  host enqueue time {cpu_ms:6.2f} ms/step (GPU step time below: the host has to stay under it)")
        return t0.elapsed_time(t1) / steps

    # pinned host -> device bandwidth of one batch (the per-step copy of the end-to-end path)
    dst = torch.empty_like(dev_x[:B])
    for _ in range(2):
        dst.copy_(host_x[:B], non_blocking=True)
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for i in range(8):
        dst.copy_(host_x[(i % pool) * B:(i % pool + 1) * B], non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    ms = h0.elapsed_time(h1) / 8
    print(f"pinned H2D of one batch ({dst.numel() * 4 / 1e6:.0f} MB): {ms:.2f} ms = {dst.numel() * 4 / ms / 1e6:.1f} GB/s")

    for rep in range(2):
        print(f"plain step loop (device batches)           {plain():7.2f} ms/step")
        print(f"train_one_epoch, device batches, no sync   {run(True, False):7.2f}")
        print(f"train_one_epoch, device batches, sync/step {run(True, True):7.2f}")
        print(f"train_one_epoch, host batches,   no sync   {run(False, False):7.2f}")
        print(f"train_one_epoch, host batches,   sync/step {run(False, True):7.2f}")


if __name__ == "__main__":
    main()
