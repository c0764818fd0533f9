#!/usr/bin/env python
"""bench.py — the headline benchmark: ViT-B/16 client-training throughput per B200 and the FedAvg
round it feeds (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the CPU port of the reference path
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank (= one FedAvg client) per GPU

A "step" is one client optimisation step on one batch of 256 synthetic 224x224x3 images:
forward, asymmetric-focal loss, backward, global-norm clip + AdamW (LLRD groups) — everything the
reference does per iteration of train.py:131-166. Prints ONE JSON line (rank 0).
  value     images/s over all GPUs, batches resident in HBM when the timed region starts
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

    cfg = model_config()
    cfg["model"]["cls_only_last_block"] = bool(args.cls_only_last_block)
    utils.seed_everything(42)
    net = model.build_model(cfg).to(dev).train()
    arena = FlatArena(net)
    fedavg.broadcast_initial(arena, net)
    opt = optim.FusedAdamW(model.get_layerwise_lr_groups(net, 1e-4, 0.75, 1e-5), weight_decay=1e-5, arena=arena)
    crit = losses.build_loss(cfg)
    agg = fedavg.FedAvgAggregator(net, arena)

    pool = 4
    g = torch.Generator().manual_seed(1000 + rank)  # client id == rank
    host_x = torch.randn(pool * BATCH, 3, IMG, IMG, generator=g).pin_memory()
    host_y = torch.randint(0, CLASSES, (pool * BATCH,), generator=g).pin_memory()
    dev_x, dev_y = host_x.to(dev), host_y.to(dev)

    def step(i: int):
        j = (i % pool) * BATCH
        x, y = dev_x[j:j + BATCH], dev_y[j:j + BATCH]
        opt.zero_grad(set_to_none=True)
        with torch.amp.autocast("cuda", dtype=torch.bfloat16):
            loss = crit(net(x)["logits"], y)
        loss.backward()
        utils.clip_grad_norm(net.parameters(), 1.0, optimizer=opt)
        opt.step()
        return loss

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

    # ---- device-resident throughput ---------------------------------------------------------------
    for i in range(args.warmup):
        step(i)
    barrier()
    ops.GEMM_TRACE = None if os.environ.get("FEDVIT_BENCH_NOTRACE") else []  # A/B switch: cost of the per-GEMM events
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for i in range(args.steps):
            loss = step(args.warmup + i)
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - n0
    trace, ops.GEMM_TRACE = ops.GEMM_TRACE or [], None
    gemm_ms = sum(a.elapsed_time(b) for a, b, _ in trace)
    gemm_flops = sum(f for _, _, f in trace)
    final_loss = float(loss.detach())
    ms_step = ms_total / args.steps
    value = args.gpus * BATCH * args.steps / (ms_total / 1e3)

    # ---- FedAvg aggregate (fold + allreduce + install), timed on the device ---------------------------
    agg.begin_round()
    agg_ms = []
    for _ in range(3):
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        agg._folded = 0
        agg.fold(STEPS_PER_ROUND * BATCH, STEPS_PER_ROUND * BATCH * args.gpus, client_id=rank)
        agg.finish()
        a1.record()
        barrier()
        agg_ms.append(max_over_ranks(a0.elapsed_time(a1)))
    aggregate_ms = min(agg_ms)

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

    if world > 1:
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt)

    if rank == 0:
        peaks = measured_peaks()
        peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        tpath = ROOT / "profiles" / "r1_ncu_traffic.json"
        traffic = json.loads(tpath.read_text()) if tpath.exists() else {}
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
                "gemm_launches": len(trace), "gemm_share_of_step": gemm_ms / ms_total if ms_total else None,
                "traffic": traffic.get("dram_bytes_per_launch"), "traffic_unit": "bytes per launch",
                "traffic_note": (f"ncu --set full, one launch of {traffic.get('kernel')} at {traffic.get('shape')}: "
                                 f"{traffic.get('dram_bytes_per_launch')} B DRAM vs {traffic.get('algorithmic_bytes_per_launch')} B "
                                 f"algorithmic, tensor pipe active {traffic.get('tensor_pipe_active_pct_of_elapsed')} % of elapsed "
                                 f"({traffic.get('source')})") if traffic else None,
                "step_tflops": value / args.gpus * TRAIN_GFLOP_PER_IMG / 1e3,
                "step_frac_of_peak": value / args.gpus * TRAIN_GFLOP_PER_IMG / 1e3 / peak,
            },
            "fedavg": {"aggregate_ms": aggregate_ms, "steps_per_round": STEPS_PER_ROUND,
                       "round_ms": STEPS_PER_ROUND * ms_step + aggregate_ms,
                       "state_bytes": arena.numel * 4},
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
