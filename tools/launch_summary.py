#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python tools/launch_summary.py gpurun_out/launches.csv [first_kernel_regex]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    # whole steps only: the optimiser sweep (adamw_kernel) closes every training step, so the slice
    # between the first and the last of its launches covers an integer number of steps
    marks = [i for i, r in enumerate(rows) if "adamw_kernel" in r["Kernel Name"]]
    steps = 0
    if len(marks) >= 2:
        rows = rows[marks[0] + 1:marks[-1] + 1]
        steps = len(marks) - 1
    agg = collections.OrderedDict()
    tot = 0.0
    for row in rows:
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", row["Kernel Name"]))
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        t = t / 1e3 if unit == "ns" else t * 1e3 if unit == "ms" else t
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    whole = f" = {steps} whole training step(s)" if steps else ""
    print(f"{len(rows)} launches{whole}, {tot / 1e3:.2f} ms total device time (serialised, cold-cache: compare shares)")
    print(f"{'us':>10} {'share':>6} {'n':>5} {'avg us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} {100 * v[1] / tot:5.1f}% {v[0]:5d} {v[1] / v[0]:9.1f}  {k[:100]}")


if __name__ == "__main__":
    main()
