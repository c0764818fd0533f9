#!/usr/bin/env python
"""BASELINE configs[3] at full size: ViT-Large/16 384 px, bf16, 16 non-IID (Dirichlet 0.5) clients with unequal
shards (512 ... 2048 samples, batch 64) over the GPUs of one box, sample-weighted FedAvg — per-rank busy time,
round wall time and the imbalance tail, for the load-balanced placement (fedavg.assign_clients, longest
processing time first) and for plain round-robin (k mod G).

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4_round.py
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fedvit_b200  # noqa: F401,E402
from fedvit_b200 import train  # noqa: E402

SIZES = [512, 640, 768, 896, 1024, 1152, 1280, 1408, 1536, 1664, 1792, 1920, 2048, 512, 1024, 2048]


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    scale = int(os.environ.get("CONFIG4_SHARD_DIV", "1"))  # >1 shrinks every shard (quick checks)
    sizes = [max(64, s // scale // 64 * 64) for s in SIZES]
    quiet = type("Q", (), {"info": staticmethod(lambda *a, **k: None)})()
    out = {}
    for policy in ("lpt", "round_robin"):
        cfg = {
            "seed": 42,
            "model": {"backbone": "vit_large_patch16_384", "num_classes": 7, "image_size": 384, "pretrained": False,
                      "drop_path_rate": 0.0, "metadata": {"enabled": False}, "classifier": {"hidden_dim": 512, "dropout": 0.0}},
            "data": {"use_segmentation_mask": False},
            "training": {"use_amp": True, "amp_dtype": "bf16", "grad_clip": 1.0, "gradient_accumulation_steps": 1,
                         "batch_size": 64, "optimizer": {"lr": 1e-4, "weight_decay": 1e-5},
                         "llrd": {"enabled": True, "decay_rate": 0.75}},
            "augmentation": {"mixup": {"alpha": 0.0}, "cutmix": {"prob": 0.0}},
            "loss": {"asymmetric": {"gamma_neg": 4, "gamma_pos": 1, "clip": 0.05}},
            "federated": {"num_clients": len(sizes), "rounds": 2, "local_epochs": 1, "samples_per_client": sizes,
                          "partition": "dirichlet", "dirichlet_alpha": 0.5, "synthetic_pool": 128, "placement": policy},
        }
        res = train.run_federated(cfg, device=dev, logger=quiet, device_resident=True)
        r = res["rounds"][-1]
        place = res["placement"]
        loads = [sum(sizes[c] for c in cs) for cs in place]
        busy = r["rank_busy_ms"]
        rate = sum(loads) / (sum(busy) / 1e3)  # images per busy GPU-second
        out[policy] = {
            "placement": place, "samples_per_rank": loads, "rank_busy_ms": busy, "round_ms": r["round_ms"],
            "aggregate_ms": r["aggregate_ms"], "aggregate_ms_last_rank": r["aggregate_ms_last_rank"],
            "images_per_s": r["images_per_s"],
            "busy_max_over_mean": max(busy) / (sum(busy) / len(busy)),
            "round_ms_over_balanced_ideal": r["round_ms"] / (sum(loads) / world / rate * 1e3),
            "mean_client_loss": r["mean_client_loss"],
        }
        del res
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"config": "configs[3]: ViT-L/16 384 bf16, 16 non-IID clients, unequal n_k", "gpus": world,
                          "samples_per_client": sizes, "state_bytes": 1216876800, **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
