#!/usr/bin/env python
"""Transpose one launch of an `ncu --page raw --csv` export into the three-column (metric, unit, value) summaries kept
under profiles/, keeping the metrics the design notes refer to.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/x.csv && python tools/ncu_summary.py /tmp/x.csv > profiles/r2_ncu_full_x.csv
"""
import csv
import re
import sys

KEEP = re.compile(
    r"Kernel Name|gpu__time_duration|dram__bytes_(read|write)\.sum$|gpu__dram_throughput|lts__throughput|"
    r"sm__throughput|sm__pipe_tensor.*cycles_active|sm__inst_executed_pipe_(xu|alu|fma|fmaheavy|lsu|uniform|tensor)|"
    r"sm__pipe_(xu|alu|fma|fmaheavy)_cycles_active|l1tex__data_pipe_(lsu|tc)_wavefronts|l1tex__throughput|"
    r"smsp__issue_active|smsp__inst_executed\.sum$|sm__warps_active|launch__(block_size|grid_size|registers_per_thread|"
    r"shared_mem_per_block_dynamic|cluster|occupancy_limit)|sm__cycles_elapsed\.(avg|max)$|smsp__cycles_active\.avg$|"
    r"sm__tmem|tmem|smsp__warp_issue_stalled.*_per_warp_active|smsp__average_warp.*_per_issue_active")


def main():
    with open(sys.argv[1]) as f:
        rows = [r for r in csv.reader(l for l in f if not l.startswith("=="))]
    names, units, vals = rows[0], rows[1], rows[2]
    out = csv.writer(sys.stdout, quoting=csv.QUOTE_ALL)
    sys.stdout.write("metric,unit,value\n")
    for n, u, v in zip(names, units, vals):
        if KEEP.search(n):
            out.writerow([n, u, v])


if __name__ == "__main__":
    main()
