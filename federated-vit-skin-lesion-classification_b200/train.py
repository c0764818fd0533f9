#!/usr/bin/env python3
"""Federated client training on B200 — the reference's ``train.py`` step loop plus the FedAvg
round loop the reference lacks.

Kept from the reference (same names / signatures / config keys):
    train_one_epoch(model, loader, criterion, optimizer, scheduler, scaler, ema, device, config,
                    epoch, logger) -> float                         (reference train.py:95-168)
    validate(model, loader, criterion, device, config) -> dict      (reference train.py:175-214)
``train_one_epoch`` is one client's local epoch. Differences, all below the seam:
  * mixed precision is bf16 (``training.amp_dtype``), so the GradScaler is an identity
    (reference: fp16 + GradScaler, train.py:144,270) — a deliberate divergence asked for by the
    baseline configuration;
  * the per-step ``loss.item()`` host sync (train.py:164) is gone: the running loss is
    accumulated on the device and read once per epoch;
  * clip + AdamW + EMA are one kernel sweep (optim.FusedAdamW).

Added (the reference has no federated code): ``run_federated`` — the round loop
shaped like the reference's epoch loop (train.py:281-319): every round each client restarts from
the global weights, trains ``local_epochs`` epochs on its shard, is folded into the weighted
accumulator; one NCCL allreduce; scheduler stepped once per round.

    python -m fedvit_b200.train --config config.yaml                       # 1 GPU
    torchrun --nproc-per-node 8 -m fedvit_b200.train --config config.yaml  # clients over 8 GPUs
"""
from __future__ import annotations

import argparse
import collections
import logging
import os
import sys
import time
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from .arena import FlatArena
from .data import SyntheticClientLoader, client_label_probs, client_sizes
from .fedavg import FedAvgAggregator, assign_clients, broadcast_initial, dist_info
from .losses import build_loss
from .model import build_model, count_parameters, get_layerwise_lr_groups
from .optim import FusedAdamW
from .utils import (EMA, MixupCutmix, WarmupCosineScheduler, clip_grad_norm, get_device, load_config,
                    mixup_criterion, seed_everything)


def setup_logging(log_dir: Optional[str] = None, tag: str = "fed") -> logging.Logger:
    logger = logging.getLogger(f"fedvit_{tag}")
    logger.setLevel(logging.INFO)
    if not logger.handlers:  # (the reference re-adds handlers on every call; once is enough)
        h = logging.StreamHandler(sys.stdout)
        h.setFormatter(logging.Formatter("%(asctime)s | %(message)s", datefmt="%H:%M:%S"))
        logger.addHandler(h)
        if log_dir:
            os.makedirs(log_dir, exist_ok=True)
            fh = logging.FileHandler(os.path.join(log_dir, f"train_{tag}.log"))
            fh.setFormatter(logging.Formatter("%(asctime)s | %(message)s"))
            logger.addHandler(fh)
    return logger


def _amp_settings(config: dict, device: torch.device):
    t = config.get("training", {})
    use_amp = bool(t.get("use_amp", True)) and device.type == "cuda"
    name = str(t.get("amp_dtype", "bf16")).lower()
    if name not in ("bf16", "bfloat16"):
        raise ValueError("training.amp_dtype: only bf16 is implemented on the B200 path")
    return use_amp, torch.bfloat16


_STAGING: Dict[tuple, tuple] = {}   # (device type, index) -> (copy stream, [slot 0, slot 1])


def _device_batches(loader, device: torch.device):
    """Yield the loader's batches on ``device`` with a one-batch look-ahead: the next batch's
    host->device copies run on a side stream while the current step computes (the reference issues
    them on the compute stream, train.py:132-136). Batches already on the device pass through.

    The copies land in two persistent staging slots (ping-pong) instead of freshly allocated
    tensors: a 154 MB allocation per step on the side stream plus ``record_stream`` made the
    caching allocator fall back to cudaMalloc / cudaFree — device-wide synchronisations — whenever
    the recorded uses had not retired (measured: 33 -> 50 ms steps on some runs)."""
    if device.type != "cuda":
        raise RuntimeError("this path trains on CUDA only")
    # the copy stream and the staging slots outlive the epoch: a fresh side stream per epoch has its
    # own (empty) allocator pool, so every epoch began with two 154 MB cudaMallocs and cold copies
    # (measured: the first timed epoch after a warm-up epoch still ran 39.8 instead of 32.6 ms/step)
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _STAGING:
        _STAGING[key] = (torch.cuda.Stream(device), [{}, {}])
    copy_stream, slots = _STAGING[key]   # slots: name -> device tensor
    slot_free = [None, None]    # event on the compute stream: the slot's last consumer was enqueued
    # whatever the previous epoch still had in flight from the slots was enqueued on the compute stream
    copy_stream.wait_stream(torch.cuda.current_stream(device))

    def stage(batch, k):
        moved = {}
        with torch.cuda.stream(copy_stream):
            if slot_free[k] is not None:
                copy_stream.wait_event(slot_free[k])
            for name, v in batch.items():
                if not torch.is_tensor(v) or v.device == device:
                    moved[name] = v
                    continue
                dst = slots[k].get(name)
                if dst is None or dst.shape != v.shape or dst.dtype != v.dtype:
                    dst = torch.empty(v.shape, dtype=v.dtype, device=device)
                    slots[k][name] = dst
                dst.copy_(v, non_blocking=True)
                moved[name] = dst
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        return moved, ev

    it = iter(loader)
    k = 0
    try:
        nxt = stage(next(it), k)
    except StopIteration:
        return
    while nxt is not None:
        cur, ev = nxt
        try:
            nxt = stage(next(it), k ^ 1)
        except StopIteration:
            nxt = None
        main = torch.cuda.current_stream(device)
        main.wait_event(ev)
        yield cur
        done = torch.cuda.Event()  # everything that reads slot k has been enqueued on `main`
        done.record(main)
        slot_free[k] = done
        k ^= 1


# ================================================================================================
# one client's local epoch
# ================================================================================================
def train_one_epoch(model: nn.Module, loader, criterion, optimizer, scheduler, scaler, ema: Optional[EMA],
                    device: torch.device, config: dict, epoch: int, logger: Optional[logging.Logger]) -> float:
    model.train()
    t = config.get("training", {})
    use_amp, amp_dtype = _amp_settings(config, device)
    grad_clip = t.get("grad_clip", 1.0)
    accum = max(1, int(t.get("gradient_accumulation_steps", 1)))
    use_meta = config.get("model", {}).get("metadata", {}).get("enabled", True)
    aug = config.get("augmentation", {})
    mixer = None  # MixUp / CutMix exactly as the reference wires them (train.py:115-124,139-150)
    mixup_a = aug.get("mixup", {}).get("alpha", 0.0)
    cutmix_p = aug.get("cutmix", {}).get("prob", 0.0)
    if mixup_a > 0 or cutmix_p > 0:
        mixer = MixupCutmix(mixup_alpha=mixup_a, cutmix_alpha=aug.get("cutmix", {}).get("alpha", 1.0),
                            cutmix_prob=cutmix_p)

    # training.cuda_graph (additive key, default off): the whole optimisation step — zero_grad, forward, loss,
    # backward, clip, fused AdamW (+ EMA) — replays as ONE CUDA graph (graphs.GraphedTrainStep). Worth it when
    # the step is launch-bound (ViT-Tiny, batch 16: 13.9 -> 2.2 ms/step); at ViT-B/16 batch 256 the GPU is the
    # limit and it changes nothing. Falls back to the eager step for batches of another shape (a ragged tail).
    graph_ok = (bool(t.get("cuda_graph", False)) and mixer is None and accum == 1 and use_amp
                and (scaler is None or not scaler.is_enabled()) and hasattr(optimizer, "enable_graph_mode"))

    loss_sum = torch.zeros((), device=device, dtype=torch.float32)
    seen = 0
    n_steps = len(loader)
    optimizer.zero_grad(set_to_none=True)
    # training.sync_loss_every_step: read every step's loss back to the host like the reference does
    # (train.py:164) — but `training.loss_read_lag` steps late (default 2), from pinned scalars, so
    # the device->host read of step i overlaps the kernels of the following steps instead of draining
    # the GPU every iteration, and a host hiccup of up to one step time never starves the device.
    sync_every_step = bool(t.get("sync_loss_every_step", False))
    lag = max(1, int(t.get("loss_read_lag", 2)))
    host_loss = 0.0
    pending = collections.deque()  # (event, pinned scalar, weight) of the last `lag` steps
    pinned = [torch.empty((), dtype=torch.float32, pin_memory=True) for _ in range(lag + 1)] if sync_every_step else None
    for step, batch in enumerate(_device_batches(loader, device)):
        images, labels, meta = batch["image"], batch["label"], batch.get("metadata")
        bs = images.size(0)

        graphed = None
        if graph_ok:
            key = (tuple(images.shape), tuple(labels.shape), id(optimizer), id(criterion), grad_clip,
                   meta is not None and use_meta)
            cached = getattr(model, "_fv_graph_step", None)
            if step == 0 and (cached is None or cached[0] != key):  # captured for the first batch's shape
                from .graphs import GraphedTrainStep

                cached = (key, GraphedTrainStep(model, criterion, optimizer, images, labels,
                                                metadata=meta if use_meta else None, grad_clip=grad_clip,
                                                amp_dtype=amp_dtype))
                model._fv_graph_step = cached
            if cached is not None and cached[0] == key:
                graphed = cached[1]
        if graphed is not None:
            loss = graphed(images, labels, meta if use_meta else None) / accum
            if ema is not None:
                ema.update()
            if sync_every_step:
                buf = pinned[step % (lag + 1)]
                buf.copy_(loss, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending.append((ev, buf, accum * bs))
                if len(pending) > lag:
                    old_ev, old_buf, w = pending.popleft()
                    old_ev.synchronize()
                    host_loss += float(old_buf) * w
            else:
                loss_sum += loss * (accum * bs)
            seen += bs
            continue

        if mixer is not None:
            images, labels_a, labels_b, lam = mixer(images, labels)
        with torch.amp.autocast(device_type=device.type, enabled=use_amp, dtype=amp_dtype):
            logits = model(images, metadata=meta if use_meta else None)["logits"]
            if mixer is not None:
                loss = mixup_criterion(criterion, logits, labels_a, labels_b, lam) / accum
            else:
                loss = criterion(logits, labels) / accum

        if scaler is not None:
            scaler.scale(loss).backward()
        else:
            loss.backward()

        if (step + 1) % accum == 0 or (step + 1) == n_steps:
            if scaler is not None and scaler.is_enabled():
                scaler.unscale_(optimizer)
                clip_grad_norm(model.parameters(), grad_clip)
                scaler.step(optimizer)
                scaler.update()
            else:
                clip_grad_norm(model.parameters(), grad_clip, optimizer=optimizer)
                optimizer.step()
            optimizer.zero_grad(set_to_none=True)
            if ema is not None:
                ema.update()

        if sync_every_step:
            buf = pinned[step % (lag + 1)]
            buf.copy_(loss.detach(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((ev, buf, accum * bs))
            if len(pending) > lag:
                old_ev, old_buf, w = pending.popleft()
                old_ev.synchronize()
                host_loss += float(old_buf) * w
        else:
            loss_sum += loss.detach() * (accum * bs)
        seen += bs
    if sync_every_step:
        for old_ev, old_buf, w in pending:
            old_ev.synchronize()
            host_loss += float(old_buf) * w
        return host_loss / max(seen, 1)
    return float(loss_sum.item()) / max(seen, 1)


# ================================================================================================
# validation
# ================================================================================================
def classification_metrics(labels: np.ndarray, preds: np.ndarray, num_classes: int) -> Dict[str, float]:
    """accuracy / balanced accuracy / macro-F1 with sklearn's conventions (mean recall over the
    classes present in ``labels``; F1 averaged over classes present in labels or predictions,
    zero_division=0) — what reference train.py:211-213 reports."""
    cm = np.zeros((num_classes, num_classes), dtype=np.int64)
    np.add.at(cm, (labels, preds), 1)
    tp = np.diag(cm).astype(np.float64)
    support, predicted = cm.sum(1).astype(np.float64), cm.sum(0).astype(np.float64)
    present = support > 0
    recall = np.divide(tp, support, out=np.zeros_like(tp), where=present)
    precision = np.divide(tp, predicted, out=np.zeros_like(tp), where=predicted > 0)
    denom = precision + recall
    f1 = np.divide(2 * precision * recall, denom, out=np.zeros_like(tp), where=denom > 0)
    used = present | (predicted > 0)
    return {
        "accuracy": float(tp.sum() / max(cm.sum(), 1)),
        "balanced_accuracy": float(recall[present].mean()) if present.any() else 0.0,
        "macro_f1": float(f1[used].mean()) if used.any() else 0.0,
    }


@torch.no_grad()
def validate(model: nn.Module, loader, criterion, device: torch.device, config: dict) -> dict:
    model.eval()
    use_amp, amp_dtype = _amp_settings(config, device)
    use_meta = config.get("model", {}).get("metadata", {}).get("enabled", True)
    loss_sum = torch.zeros((), device=device, dtype=torch.float32)
    seen = 0
    preds: List[torch.Tensor] = []
    gold: List[torch.Tensor] = []
    for batch in _device_batches(loader, device):
        images, labels, meta = batch["image"], batch["label"], batch.get("metadata")
        with torch.amp.autocast(device_type=device.type, enabled=use_amp, dtype=amp_dtype):
            logits = model(images, metadata=meta if use_meta else None)["logits"]
            loss = criterion(logits, labels)
        bs = images.size(0)
        loss_sum += loss * bs
        seen += bs
        preds.append(logits.argmax(1))
        gold.append(labels.clone())  # the staging slot behind `labels` is reused two batches later
    p = torch.cat(preds).cpu().numpy()
    y = torch.cat(gold).cpu().numpy()
    out = {"loss": float(loss_sum.item()) / max(seen, 1)}
    out.update(classification_metrics(y, p, int(config.get("model", {}).get("num_classes", 8))))
    return out


# ================================================================================================
# FedAvg round loop
# ================================================================================================
def build_client_loaders(config: dict, client_ids: List[int], device_resident: bool, device) -> Dict[int, SyntheticClientLoader]:
    fed = config.get("federated", {})
    m = config.get("model", {})
    t = config.get("training", {})
    sizes = client_sizes(config)
    k = len(sizes)
    probs = client_label_probs(k, int(m.get("num_classes", 8)), fed.get("partition", "iid"),
                               float(fed.get("dirichlet_alpha", 0.5)), int(config.get("seed", 42)))
    meta_on = m.get("metadata", {}).get("enabled", True)
    ch = 4 if config.get("data", {}).get("use_segmentation_mask", False) else 3
    return {
        c: SyntheticClientLoader(
            c, sizes[c], int(t.get("batch_size", 16)), int(m.get("image_size", 384)), ch,
            int(m.get("num_classes", 8)), None if fed.get("partition", "iid") == "iid" else probs[c],
            int(m.get("metadata", {}).get("input_dim", 13)) if meta_on else 0,
            pool=fed.get("synthetic_pool"), device=device if device_resident else None)
        for c in client_ids
    }


def run_federated(config: dict, device: Optional[torch.device] = None, logger: Optional[logging.Logger] = None,
                  loaders: Optional[Dict[int, SyntheticClientLoader]] = None, val_loader=None,
                  device_resident: bool = False) -> dict:
    """FedAvg over ``federated.num_clients`` clients for ``federated.rounds`` rounds. Returns the
    per-round records (wall time from CUDA events, images/s, mean client loss, allreduce time, every
    rank's busy time).

    Policies the reference cannot supply (it has no federated loop):
      * client optimiser state (AdamW moments, step count) is reset for every client, every round;
      * the lr schedule is stepped once per round (train.py:297 steps it once per epoch) — with
        ``scheduler.warmup_epochs > 0`` the reference's formula starts at lr = 0 (utils.py:179-185), so the
        first ROUND trains at lr 0, as the reference's first epoch does;
      * ``training.ema`` is a SERVER-side average here: one shadow of the GLOBAL weights, updated once per
        round from w^{r+1} after the aggregate (``decay`` applies per round) and used for validation.
        A client-side shadow carried across clients and rounds while ``load_global`` resets the weights
        under it would blend unrelated trajectories and differ from rank to rank;
      * clients are placed by ``fedavg.assign_clients`` (round-robin for equal shards, longest-processing-
        time-first on n_k otherwise)."""
    fed = config.get("federated", {})
    t = config.get("training", {})
    rank, world = dist_info()
    device = device or get_device("auto")
    logger = logger or setup_logging(tag=f"r{rank}")
    seed_everything(int(config.get("seed", 42)))

    model = build_model(config).to(device)
    arena = FlatArena(model)
    broadcast_initial(arena, model)
    opt_cfg, llrd = t.get("optimizer", {}), t.get("llrd", {})
    groups = get_layerwise_lr_groups(
        model, base_lr=opt_cfg.get("lr", 1e-4),
        decay_rate=llrd.get("decay_rate", 0.75) if llrd.get("enabled", True) else 1.0,
        weight_decay=opt_cfg.get("weight_decay", 1e-5))
    optimizer = FusedAdamW(groups, weight_decay=opt_cfg.get("weight_decay", 1e-5), arena=arena)
    rounds = int(fed.get("rounds", 1))
    sched_cfg = t.get("scheduler", {})
    scheduler = WarmupCosineScheduler(optimizer, warmup_epochs=sched_cfg.get("warmup_epochs", 0),
                                      total_epochs=max(rounds, 1), min_lr=sched_cfg.get("min_lr", 1e-6))
    ema_cfg = t.get("ema", {})
    # server-side EMA of the global model (see the docstring): NOT attached to the client optimiser
    ema = EMA(model, decay=ema_cfg.get("decay", 0.9995)) if ema_cfg.get("enabled", False) else None
    scaler = torch.amp.GradScaler(device.type, enabled=False)  # bf16: no loss scaling
    criterion = build_loss(config).to(device)

    sizes = client_sizes(config)
    n_total = sum(sizes)
    placement = assign_clients(sizes, world, str(fed.get("placement", "auto")))
    mine = placement[rank]
    root = next(r for r, cs in enumerate(placement) if 0 in cs)  # integer buffers come from client 0
    if loaders is None:
        loaders = build_client_loaders(config, mine, device_resident, device)
    agg = FedAvgAggregator(model, arena)
    local_epochs = int(fed.get("local_epochs", 1))
    if rank == 0:
        logger.info(f"FedAvg: {len(sizes)} clients over {world} GPU(s), {rounds} rounds x {local_epochs} "
                    f"local epoch(s); params {count_parameters(model):,}; clients of rank 0: {mine}")

    records = []
    for rnd in range(1, rounds + 1):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        ev[0].record()
        agg.begin_round(len(mine))
        losses = []
        for c in mine:
            agg.load_global()
            optimizer.reset_state()  # canonical FedAvg: client optimiser state does not survive the round
            for e in range(local_epochs):
                losses.append(train_one_epoch(model, loaders[c], criterion, optimizer, scheduler, scaler,
                                              None, device, config, e + 1, logger))
            if c == mine[-1]:
                ev[1].record()  # the aggregate: this rank's last fold (in place) + allreduce + bf16 re-cast
            agg.fold(sizes[c], n_total, client_id=c, last=c == mine[-1])
        if not mine:
            ev[1].record()
        agg.finish(root=root)
        if ema is not None:
            ema.update()  # shadow <- decay * shadow + (1 - decay) * w^{r+1}: identical on every rank
        ev[2].record()
        torch.cuda.synchronize(device)
        scheduler.step()
        ms = torch.tensor([ev[0].elapsed_time(ev[2]), ev[1].elapsed_time(ev[2])], device=device)
        busy = torch.tensor([ev[0].elapsed_time(ev[1])], device=device)  # this rank's local training + folds
        agg_min = torch.tensor([ev[1].elapsed_time(ev[2])], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(agg_min, op=dist.ReduceOp.MIN)
            all_busy = [torch.zeros_like(busy) for _ in range(world)]
            dist.all_gather(all_busy, busy)
            busy = torch.cat(all_busy)
        images = sum((s // int(t.get("batch_size", 16))) * int(t.get("batch_size", 16)) for s in sizes) * local_epochs
        # sample-weighted mean of the clients' epoch losses over ALL ranks (logging only)
        lw = torch.tensor([sum(l * sizes[c] for l, c in zip(losses[local_epochs - 1::local_epochs], mine)),
                           float(sum(sizes[c] for c in mine))], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(lw)
        # aggregate_ms is the MAX over ranks of (own last fold -> w^{r+1} installed): on every rank but the last
        # to arrive it includes the wait for the slowest client; aggregate_ms_last_rank (the MIN) is what the
        # fold + allreduce + re-cast cost once everybody is there
        rec = {"round": rnd, "round_ms": float(ms[0]), "aggregate_ms": float(ms[1]),
               "aggregate_ms_last_rank": float(agg_min[0]),
               "images_per_s": images / (float(ms[0]) / 1e3),
               "mean_client_loss": float(lw[0] / lw[1]) if float(lw[1]) > 0 else None,
               "rank_busy_ms": [float(v) for v in busy.tolist()]}
        if val_loader is not None:
            if ema is not None:
                ema.apply_shadow()
            rec["val"] = validate(model, val_loader, criterion, device, config)
            if ema is not None:
                ema.restore()
        records.append(rec)
        if rank == 0:
            logger.info(f"  round {rnd:02d} | {rec['round_ms']:.1f} ms | {rec['images_per_s']:.0f} img/s | "
                        f"aggregate {rec['aggregate_ms']:.2f} ms | loss {rec['mean_client_loss']}")
    return {"rounds": records, "model": model, "arena": arena, "ema": ema, "placement": placement}


def main() -> None:
    ap = argparse.ArgumentParser(description="FedAvg ViT client training on B200 (synthetic shards)")
    ap.add_argument("--config", type=str, default=os.path.join(os.path.dirname(__file__), "config.yaml"))
    ap.add_argument("--log", type=str, default=None)
    ap.add_argument("--seed", type=int, default=None)
    args = ap.parse_args()
    config = load_config(args.config)
    if args.seed is not None:
        config["seed"] = args.seed
    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    rank, _ = dist_info()
    run_federated(config, logger=setup_logging(args.log, tag=f"r{rank}"))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
