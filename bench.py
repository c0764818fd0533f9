#!/usr/bin/env python
"""bench.py — the headline benchmark: ViT-B/16 client-training throughput per B200 and the FedAvg
round it feeds (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the CPU port of the reference path
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank (= one FedAvg client) per GPU

A "step" is one client optimisation step on one batch of 256 synthetic 224x224x3 images:
forward, asymmetric-focal loss, backward, global-norm clip + AdamW (LLRD groups) — everything the
reference does per iteration of train.py:131-166. Prints ONE JSON line (rank 0).
  value     images/s over all GPUs, batches resident in HBM when the timed region starts (no
            instrumentation inside the region)
  fedavg    3 REAL rounds through train.run_federated (first = warm-up): round_ms, aggregate_ms, the gap to
            16 x ms_per_step; parity_rel = NCCL aggregate vs the sequential fixed-order sum (gate 1e-6)
  extra_configs  short runs of BASELINE configs 3 / 4 / 5 (masked ViT-B, ViT-L/16 384 step and non-IID
            round with unequal shards, eval sweep)
  e2e       the same through the public API (train.train_one_epoch) from pinned HOST batches:
            per-step host->device copies and a device->host read of the loss are inside the timing
  roofline  tensor-core GEMM kernel: algorithmic FLOPs / CUDA-event time of its launches, live in
            the timed region, vs the measured dense-bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle (CPU port of the reference path) on this box's host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ViT-B/16 train images/s per B200; FedAvg round wall-time at 1/2/4/8 GPUs"
UNIT = "images/s"
BATCH = 256
IMG = 224
CLASSES = 7
STEPS_PER_ROUND = 16  # n_k = 4096 samples per client / batch 256
TRAIN_GFLOP_PER_IMG = 105.381  # 3 x 35.127 GFLOP forward (2*MACs), ViT-B/16 224


def model_config() -> dict:
    return {
        "seed": 42,
        "model": {"backbone": "vit_base_patch16_224", "num_classes": CLASSES, "image_size": IMG, "pretrained": False,
                  "drop_path_rate": 0.0, "metadata": {"enabled": False}, "classifier": {"hidden_dim": 512, "dropout": 0.0}},
        "data": {"use_segmentation_mask": False},
        "training": {"use_amp": True, "amp_dtype": "bf16", "grad_clip": 1.0, "gradient_accumulation_steps": 1,
                     "batch_size": BATCH, "optimizer": {"lr": 1e-4, "weight_decay": 1e-5},
                     "llrd": {"enabled": True, "decay_rate": 0.75}},
        "augmentation": {"mixup": {"alpha": 0.0}, "cutmix": {"prob": 0.0}},
        "loss": {"asymmetric": {"gamma_neg": 4, "gamma_pos": 1, "clip": 0.05}},
    }


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": "ViT-Base/16 224px bf16, 8 FedAvg clients (one per B200), 1 local epoch per round "
                    f"[configs[1]; {n_gpus} client(s) on {n_gpus} GPU(s) here]",
        "per_gpu_batch": BATCH, "global_batch": BATCH * n_gpus, "image": f"{IMG}x{IMG}x3", "classes": CLASSES,
        "steps_per_round": STEPS_PER_ROUND, "optimizer": "AdamW, LLRD 0.75, clip 1.0 (fused sweep)",
        "parallelism": f"fedavg-clients x{n_gpus} (independent replicas + 1 NCCL allreduce per round)",
        "l2_policy": "inputs larger than L2: 4 distinct 154 MB batches cycled, ~17 GB of activations per step",
    }


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling guide)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.rows = []
        self._stop = threading.Event()
        self._index = index
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self) -> None:
        if self._run_nvml():
            return
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self._index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def _run_nvml(self) -> bool:
        """The same readings straight from NVML, every 10 ms (a timed region of ten steps lasts a third of a
        second; an nvidia-smi process per sample gets one or two readings out of it). Rows keep the
        nvidia-smi column layout. False = NVML not usable here, fall back to nvidia-smi."""
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._index)
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            reasons_fn(h)
        except Exception:
            return False
        bits = (0x8, 0x40, 0x20, 0x4)  # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = int(reasons_fn(h))
                try:
                    power = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                except Exception:
                    power = 0.0
                self.rows.append([str(sm), str(mx), f"{power:.1f}"] +
                                 ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            self._stop.wait(0.01)
        return True

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_tflops": d.get("bf16_tflops"), "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------------------------
def cpu_port_images_per_s(steps: int, warmup: int, sample_batch: int, threads: int):
    """The oracle's train step (CPU port of model.py + losses.py + clip + LLRD AdamW) on host cores."""
    import torch

    from oracle import asl, isic

    torch.set_num_threads(threads)
    cfg = model_config()
    torch.manual_seed(0)
    m = isic.model_from_config(cfg).train()
    opt = torch.optim.AdamW(isic.llrd_groups(m, 1e-4, 0.75, 1e-5), weight_decay=1e-5)
    g = torch.Generator().manual_seed(1000)
    x = torch.randn(sample_batch, 3, IMG, IMG, generator=g)
    y = torch.randint(0, CLASSES, (sample_batch,), generator=g)

    def one():
        opt.zero_grad(set_to_none=True)
        loss = asl.asymmetric_focal_loss(m(x)["logits"], y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return steps * sample_batch / dt, dt / steps * 1e3


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 8
    ips, ms = cpu_port_images_per_s(args.steps, args.warmup, sample, threads)
    desc = (f"{args.steps} timed steps of {sample} images each (the GPU arm's step is {BATCH}); fp32 CPU port of the "
            "reference path (oracle/: timm-ViT restatement + model.py head + losses.py + clip + LLRD AdamW) — the "
            "reference itself cannot run (timm not installed, /root/reference absent on this box)")
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def run_gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist

    import fedvit_b200  # noqa: F401
    from fedvit_b200 import _lib, fedavg, losses, model, ops, optim, train, utils
    from fedvit_b200.arena import FlatArena

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback (use --impl reference for the CPU port)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL (and anything else native) may write its version / debug lines to file descriptor 1:
        # point fd 1 at stderr for the duration of the run and restore it for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    else:
        saved_stdout = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def make_stepper(cfg, batch, img, chans, seed_rank=True):
        """Model + arena + fused optimiser + a pool of 4 distinct device-resident batches (larger than L2
        together with the ~17 GB of activations a step writes) and the closure that runs one client step."""
        utils.seed_everything(42)
        net = model.build_model(cfg).to(dev).train()
        arena = FlatArena(net)
        fedavg.broadcast_initial(arena, net)
        opt = optim.FusedAdamW(model.get_layerwise_lr_groups(net, 1e-4, 0.75, 1e-5), weight_decay=1e-5, arena=arena)
        crit = losses.build_loss(cfg)
        pool = 4
        g = torch.Generator().manual_seed(1000 + (rank if seed_rank else 0))  # client id == rank
        host_x = torch.randn(pool * batch, chans, img, img, generator=g)
        if chans == 4:  # lesion-mask plane in {-1, +1} (data.py:153-154)
            host_x[:, 3] = (torch.bernoulli(torch.full((pool * batch, img, img), 0.3), generator=g) - 0.5) / 0.5
        host_x = host_x.pin_memory()
        host_y = torch.randint(0, CLASSES, (pool * batch,), generator=g).pin_memory()
        dev_x, dev_y = host_x.to(dev), host_y.to(dev)

        def step(i: int):
            j = (i % pool) * batch
            x, y = dev_x[j:j + batch], dev_y[j:j + batch]
            opt.zero_grad(set_to_none=True)
            with torch.amp.autocast("cuda", dtype=torch.bfloat16):
                loss = crit(net(x)["logits"], y)
            loss.backward()
            utils.clip_grad_norm(net.parameters(), 1.0, optimizer=opt)
            opt.step()
            return loss

        return net, arena, opt, crit, step, host_x, host_y, pool

    def time_steps(step, warmup, steps, sample_clocks=False):
        for i in range(warmup):
            step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler is not None:
            sampler.__enter__()
        barrier()
        e0.record()
        for i in range(steps):
            loss = step(warmup + i)
        e1.record()
        barrier()
        if sampler is not None:
            sampler.__exit__()
        return max_over_ranks(e0.elapsed_time(e1)), loss, sampler

    cfg = model_config()
    cfg["model"]["cls_only_last_block"] = bool(args.cls_only_last_block)
    net, arena, opt, crit, step, host_x, host_y, pool = make_stepper(cfg, BATCH, IMG, 3)

    # ---- headline: device-resident throughput, nothing but the steps inside the timed region --------
    ops.GEMM_TRACE = None
    n0 = _lib.launch_count()
    ms_total, loss, clocks = time_steps(step, args.warmup, args.steps, sample_clocks=True)
    launches = _lib.launch_count() - n0
    state_bytes = arena.numel * 4
    final_loss = float(loss.detach())
    ms_step = ms_total / args.steps
    value = args.gpus * BATCH * args.steps / (ms_total / 1e3)

    # ---- roofline leg: a SECOND pass with every tensor-core GEMM launch bracketed by CUDA events -------
    # (2 960 event pairs per 10 steps: kept out of the headline region — r1 timed both at once)
    trace_steps = max(2, min(5, args.steps))
    ops.GEMM_TRACE = []
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(trace_steps):
        step(args.warmup + args.steps + i)
    t1.record()
    barrier()
    trace_ms_total = t0.elapsed_time(t1)
    trace, ops.GEMM_TRACE = ops.GEMM_TRACE or [], None
    gemm_ms = sum(a.elapsed_time(b) for a, b, _ in trace)
    gemm_flops = sum(f for _, _, f in trace)

    # ---- end to end through the public API: pinned host batches, H2D per step, loss read per step ----
    class HostLoader:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __iter__(self):
            for i in range(self.n):
                j = (i % pool) * BATCH
                yield {"image": host_x[j:j + BATCH], "label": host_y[j:j + BATCH]}

    cfg_e2e = model_config()
    cfg_e2e["training"]["sync_loss_every_step"] = True
    train.train_one_epoch(net, HostLoader(max(args.warmup, 3)), crit, opt, None, None, None, dev, cfg_e2e, 0, None)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    train.train_one_epoch(net, HostLoader(args.steps), crit, opt, None, None, None, dev, cfg_e2e, 1, None)
    t1.record()
    barrier()
    e2e_ms = max_over_ranks(t0.elapsed_time(t1))
    e2e_value = args.gpus * BATCH * args.steps / (e2e_ms / 1e3)

    # ---- FedAvg aggregate alone (last fold in place + allreduce + bf16 re-cast), timed on the device ----
    agg = fedavg.FedAvgAggregator(net, arena)
    n_k = STEPS_PER_ROUND * BATCH
    agg_ms = []
    for _ in range(3):
        agg.begin_round(1)
        agg.load_global()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        agg.fold(n_k, n_k * args.gpus, client_id=rank, last=True)
        agg.finish()
        a1.record()
        barrier()
        agg_ms.append(max_over_ranks(a0.elapsed_time(a1)))
    aggregate_ms = min(agg_ms)

    # ---- NCCL aggregate vs the sequential fixed-order sum (north_star: 1e-6 relative) --------------------
    # every rank contributes a different pseudo-client (its trained weights + a rank-seeded perturbation)
    # with unequal n_k; the product path (pre-scaled fold + ncclAllReduce) against acc = fl(c_0 w_0),
    # acc = fl(acc + fl(c_k w_k)) in rank order on one GPU — the oracle's order of operations
    parity_rel = None
    sizes_p = [n_k + 256 * r for r in range(args.gpus)]
    with torch.no_grad():
        base = arena.params.clone()
        gp = torch.Generator(device=dev).manual_seed(77 + rank)
        arena.params.add_(torch.randn(arena.numel, device=dev, generator=gp) * 1e-2 * base.abs().mean())
        mine_w = arena.params.clone()
        agg.begin_round(1)
        agg.load_global()
        agg.fold(sizes_p[rank], sum(sizes_p), client_id=rank, last=True)
        agg.finish()
        if world > 1:
            gathered = [torch.empty_like(mine_w) for _ in range(world)]
            dist.all_gather(gathered, mine_w)
            seq = None
            for r in range(world):
                term = gathered[r] * torch.tensor(fedavg.client_weight(sizes_p[r], sum(sizes_p)), device=dev)
                seq = term if seq is None else seq + term
            parity_rel = float((arena.params.double() - seq.double()).norm() / seq.double().norm())
            del gathered, seq
        else:
            seq = mine_w * torch.tensor(fedavg.client_weight(sizes_p[0], sum(sizes_p)), device=dev)
            parity_rel = float((arena.params.double() - seq.double()).norm() / seq.double().norm())
        arena.params.copy_(base)
        arena.refresh_lp(force=True)
        del base, mine_w

    if world > 1:
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt)

    # free the headline model before the round / extra-config legs build their own
    del net, arena, opt, crit, step, agg, host_x, host_y
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    quiet = type("Q", (), {"info": staticmethod(lambda *a, **k: None)})()

    # ---- the metric's second half: REAL FedAvg rounds through train.run_federated ------------------------
    # configs[1]: one client per GPU, n_k = 4096 (16 steps of 256), 1 local epoch; 3 rounds, first = warm-up;
    # CUDA events around the whole round incl. its opening barrier (run_federated's own timing, max over ranks)
    fed_cfg = model_config()
    fed_cfg["federated"] = {"num_clients": args.gpus, "rounds": 3, "local_epochs": 1, "samples_per_client": n_k,
                            "partition": "iid", "synthetic_pool": 4 * BATCH}
    fed_out = train.run_federated(fed_cfg, device=dev, logger=quiet, device_resident=True)
    rounds = fed_out["rounds"]
    timed = rounds[1:]
    round_ms = min(r["round_ms"] for r in timed)
    round_info = {
        "round_ms": round_ms, "round_ms_all": [r["round_ms"] for r in rounds], "warmup_rounds": 1,
        "aggregate_ms_in_round": min(r["aggregate_ms"] for r in timed),
        "aggregate_ms_in_round_last_rank": min(r["aggregate_ms_last_rank"] for r in timed),
        "round_images_per_s": args.gpus * n_k / (round_ms / 1e3),
        "steps_x_ms_step": STEPS_PER_ROUND * ms_step,
        "gap_to_steps_ms": round_ms - STEPS_PER_ROUND * ms_step,
        "api": "fedvit_b200.train.run_federated, device-resident client shards (pool of 4 batches cycled), "
               "load_global + optimiser reset + 16 x train_one_epoch steps + in-place fold + allreduce + install",
    }
    del fed_out
    gc.collect()
    torch.cuda.empty_cache()

    # ---- configs 3 / 4 / 5: short driver-visible runs (not the headline; each a few seconds) --------------
    extra = {}
    if not args.no_extra:
        peaks_x = measured_peaks()
        peak_x = peaks_x["bf16_tflops_sustained"] or peaks_x["bf16_tflops"]
        # config 4 as a ROUND at every N: ViT-L/16 384, 2 clients per GPU, unequal non-IID shards
        k4 = 2 * args.gpus
        sizes4 = [64 * (2 + (5 * c) % 7) for c in range(k4)]  # 128 ... 512 samples, batch 64: 2 ... 8 steps
        cfg4 = model_config()
        cfg4["model"].update({"backbone": "vit_large_patch16_384", "image_size": 384})
        cfg4["training"]["batch_size"] = 64
        cfg4["federated"] = {"num_clients": k4, "rounds": 2, "local_epochs": 1, "samples_per_client": sizes4,
                             "partition": "dirichlet", "dirichlet_alpha": 0.5, "synthetic_pool": 128}
        out4 = train.run_federated(cfg4, device=dev, logger=quiet, device_resident=True)
        r4 = out4["rounds"][-1]
        place = out4["placement"]
        loads = [sum(sizes4[c] for c in cs) for cs in place]
        busy = r4["rank_busy_ms"]
        rate = sum(loads) / (sum(busy) / 1e3)  # images per GPU-second actually achieved while busy
        extra["config4_round_vit_large_384"] = {
            "workload": f"ViT-Large/16 384px bf16, {k4} non-IID (Dirichlet 0.5) clients over {args.gpus} GPU(s), "
                        "unequal n_k, sample-weighted FedAvg; round 2 of 2 (shards scaled down 4x from 512...2048)",
            "samples_per_client": sizes4, "placement": place, "samples_per_rank": loads,
            "round_ms": r4["round_ms"], "aggregate_ms": r4["aggregate_ms"],
            "aggregate_ms_last_rank": r4["aggregate_ms_last_rank"], "rank_busy_ms": busy,
            "images_per_s": r4["images_per_s"],
            "imbalance_max_over_mean_busy": max(busy) / (sum(busy) / len(busy)),
            "round_ms_over_balanced_ideal": r4["round_ms"] / (sum(loads) / args.gpus / rate * 1e3),
            "state_bytes": out4["arena"].numel * 4,
        }
        del out4
        gc.collect()
        torch.cuda.empty_cache()
        if world == 1:
            # config 3: masked path (4-channel patch GEMM, K = 1024)
            cfg3 = model_config()
            cfg3["data"]["use_segmentation_mask"] = True
            s3 = make_stepper(cfg3, BATCH, IMG, 4)
            ms3, _, _ = time_steps(s3[4], 3, 5)
            v3 = BATCH * 5 / (ms3 / 1e3)
            extra["config3_vit_base_masked"] = {
                "workload": "ViT-Base/16 224px, 4-channel input (RGB + lesion mask), bf16, batch 256, 5 timed steps",
                "images_per_s": v3, "ms_per_step": ms3 / 5, "train_gflop_per_img": 105.612,
                "step_frac_of_peak": v3 * 105.612 / 1e3 / peak_x}
            del s3
            gc.collect()
            torch.cuda.empty_cache()
            # config 4: ViT-L/16 384, batch 64, one GPU's step
            cfg4s = model_config()
            cfg4s["model"].update({"backbone": "vit_large_patch16_384", "image_size": 384})
            s4 = make_stepper(cfg4s, 64, 384, 3)
            ms4, _, _ = time_steps(s4[4], 3, 3)
            v4 = 64 * 3 / (ms4 / 1e3)
            extra["config4_vit_large_384_step"] = {
                "workload": "ViT-Large/16 384px bf16, batch 64, 3 timed client steps on one GPU",
                "images_per_s": v4, "ms_per_step": ms4 / 3, "train_gflop_per_img": 1146.395,
                "step_frac_of_peak": v4 * 1146.395 / 1e3 / peak_x}
            del s4
            gc.collect()
            torch.cuda.empty_cache()
            # config 5: forward-only eval sweep (eager; CUDA-graph replay where the forward is launch-bound)
            from fedvit_b200.graphs import GraphedForward

            utils.seed_everything(42)
            enet = model.build_model(model_config()).to(dev).eval()
            FlatArena(enet)
            sweep = []
            for b in (1, 32, 256, 1024):
                xb = torch.randn(b, 3, IMG, IMG, device=dev)
                row = {"batch": b}
                for mode in (("eager", "graph") if b <= 32 else ("eager",)):
                    if mode == "graph":
                        fwd = GraphedForward(enet, xb)
                        run = lambda: fwd(xb)
                    else:
                        def run():
                            with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16):
                                return enet(xb)["logits"]
                    for _ in range(5):
                        run()
                    torch.cuda.synchronize(dev)
                    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    iters = 20
                    q0.record()
                    for _ in range(iters):
                        run()
                    q1.record()
                    torch.cuda.synchronize(dev)
                    ms = q0.elapsed_time(q1) / iters
                    row[f"{mode}_ms"] = ms
                    row[f"{mode}_images_per_s"] = b / (ms / 1e3)
                best = max(v for k, v in row.items() if k.endswith("images_per_s"))
                row["fwd_frac_of_peak"] = best * 35.127 / 1e3 / peak_x
                sweep.append(row)
            extra["config5_eval_sweep_vit_base"] = {
                "workload": "ViT-Base/16 forward-only (validate's per-batch work), bf16, 20 timed iterations after 5 warm-up",
                "fwd_gflop_per_img": 35.127, "rows": sweep}
            del enet
            gc.collect()
            torch.cuda.empty_cache()

    if rank == 0:
        peaks = measured_peaks()
        peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        tpath = next((q for q in (ROOT / "profiles" / "r2_ncu_traffic.json", ROOT / "profiles" / "r1_ncu_traffic.json")
                      if q.exists()), None)
        traffic = json.loads(tpath.read_text()) if tpath else {}
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(args.gpus),
                           **({"cls_only_last_block": "opt-in: last block's proj / LN2 / MLP on the cls rows only "
                                                      "(exact; FLOP-based fractions below still count the dense model)"}
                              if args.cls_only_last_block else {})),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * 3 * IMG * IMG * 4 + BATCH * 8,
                    "d2h_bytes_per_step": 4,
                    "api": "fedvit_b200.train.train_one_epoch: pinned host batches copied per step on a side stream, "
                           "every step's loss read back to the host (two steps late, pinned scalars)"},
            "gpu_launches": launches,
            "roofline": {
                "bound": "tensor", "kernel": "gemm_tc2_kernel / gemm_tc_kernel (tcgen05 bf16 GEMM: CTA-pair and single-CTA variants, all epilogues)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if achieved else None,
                "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                "measured_in": f"a separate pass of {trace_steps} steps with per-launch CUDA events (the headline region "
                               "carries no instrumentation)",
                "gemm_launches": len(trace), "gemm_share_of_step": gemm_ms / trace_ms_total if trace_ms_total else None,
                "traffic": traffic.get("dram_bytes_per_launch"), "traffic_unit": "bytes per launch",
                "traffic_note": (f"ncu --set full, one launch of {traffic.get('kernel')} at {traffic.get('shape')}: "
                                 f"{traffic.get('dram_bytes_per_launch')} B DRAM vs {traffic.get('algorithmic_bytes_per_launch')} B "
                                 f"algorithmic, tensor pipe active {traffic.get('tensor_pipe_active_pct_of_elapsed')} % of elapsed "
                                 f"({traffic.get('source')}; captured at {traffic.get('commit', 'round 1')})") if traffic else None,
                "step_tflops": value / args.gpus * TRAIN_GFLOP_PER_IMG / 1e3,
                "step_frac_of_peak": value / args.gpus * TRAIN_GFLOP_PER_IMG / 1e3 / peak,
            },
            "fedavg": dict(round_info, aggregate_ms=aggregate_ms, steps_per_round=STEPS_PER_ROUND,
                           state_bytes=state_bytes,
                           parity_rel=parity_rel, parity_gate=1e-6,
                           parity_note="NCCL allreduce of pre-scaled client weights vs the sequential fixed-order "
                                       "fp32 sum in rank order (all-gathered, summed on one GPU)"),
            "extra_configs": extra,
            "loss": final_loss,
        }
        if args.gpus == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            ips, ms = cpu_port_images_per_s(2, 1, 16, threads)
            line["cpu_baseline"] = {
                "value": ips, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "2 timed train steps (after 1 warm-up) of 16 images, fp32, ViT-B/16 224 — oracle/ CPU port "
                          "of the reference path on this box's host cores"}
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        if saved_stdout is not None:
            os.dup2(2, 1)  # teardown chatter goes to stderr again
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="fedvit", choices=["fedvit", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs legs (configs 3 / 4 / 5)")
    ap.add_argument("--cls-only-last-block", action="store_true",
                    help="opt-in model.cls_only_last_block: the last block's token-wise tail on the cls rows only "
                         "(same logits / gradients, ~4 %% fewer FLOPs); NOT the default measurement")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "fedvit":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
