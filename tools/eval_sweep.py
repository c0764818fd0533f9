#!/usr/bin/env python
"""BASELINE config 5: ViT-Base/16 forward-only eval sweep, batch 1..1024, 1 GPU, bf16.

Times `model(images)["logits"].argmax(1)` (what `validate` does per batch, reference train.py:199-205)
with CUDA events after 5 warm-up iterations, >= 20 timed iterations, and prints one JSON line per
batch size: images/s, ms/batch, fraction of the measured dense-bf16 roofline (35.127 GFLOP/img fwd).
    python tools/eval_sweep.py [--max-batch 1024] > gpurun_out/eval_sweep.jsonl
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch

    import fedvit_b200  # noqa: F401
    from bench import measured_peaks, model_config
    from fedvit_b200 import model
    from fedvit_b200.arena import FlatArena

    ap = argparse.ArgumentParser()
    ap.add_argument("--max-batch", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--graph", action="store_true", help="replay the forward as one CUDA graph per batch size")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = model.build_model(model_config()).to(dev).eval()
    FlatArena(net)
    peak = measured_peaks()["bf16_tflops_sustained"]
    b = 1
    while b <= args.max_batch:
        x = torch.randn(b, 3, 224, 224, device=dev)
        if args.graph:
            from fedvit_b200.graphs import GraphedForward

            fwd = GraphedForward(net, x)
            run = lambda: fwd(x).argmax(1)
        else:
            run = lambda: net(x)["logits"].argmax(1)
        with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16):
            for _ in range(5):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                pred = run()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        ips = b / ms * 1e3
        print(json.dumps({"batch": b, "cuda_graph": bool(args.graph), "ms_per_batch": ms, "images_per_s": ips,
                          "fwd_tflops": ips * 35.127 / 1e3, "frac_of_sustained_bf16_peak": ips * 35.127 / 1e3 / peak}),
              flush=True)
        b *= 2


if __name__ == "__main__":
    main()
