#!/usr/bin/env python
"""Per-kernel counts of the SASS instructions that show which hardware paths libfedvit.so uses
(tcgen05 MMA / TMEM traffic / TMA loads, stores, reductions / bulk copies / mbarrier), from
`cuobjdump -sass` of the built objects. No GPU needed.

    python tools/sass_summary.py > profiles/r1_sass_mnemonics.txt
"""
from __future__ import annotations

import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "federated-vit-skin-lesion-classification_b200" / "build"
PATTERNS = collections.OrderedDict([
    ("tcgen05.mma", re.compile(r"\bUTC[A-Z]*MMA")),
    ("tcgen05.ld/st", re.compile(r"\b(LDTM|STTM)")),
    ("tcgen05.commit/alloc", re.compile(r"\bUTC(BAR|ATOMSWS)")),
    ("TMA load", re.compile(r"\bUTMALDG")),
    ("TMA store", re.compile(r"\bUTMASTG")),
    ("TMA reduce", re.compile(r"\bUTMAREDG")),
    ("TMA prefetch", re.compile(r"\bUTMAPF")),
    ("bulk copy", re.compile(r"\bUBLKCP")),
    ("mbarrier", re.compile(r"\bSYNCS")),
    ("FFMA2", re.compile(r"\b(FFMA2|FMUL2|FADD2)")),
    ("MUFU", re.compile(r"\bMUFU")),
    ("HMMA (legacy)", re.compile(r"\bHMMA")),
])


def main() -> None:
    objs = sorted(BUILD.glob("*.o"))
    if not objs:
        raise SystemExit("build the library first: python -c 'import __graft_entry__ as g; g.build()'")
    print(f"{'kernel':58} " + " ".join(f"{k:>12}" for k in PATTERNS))
    for obj in objs:
        out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
        name, counts = None, None
        rows = []
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if name:
                    rows.append((name, counts))
                name, counts = m.group(1), collections.Counter()
                continue
            if name:
                for k, pat in PATTERNS.items():
                    if pat.search(line):
                        counts[k] += 1
        if name:
            rows.append((name, counts))
        for name, counts in rows:
            if not any(counts[k] for k in PATTERNS if k not in ("MUFU", "FFMA2")):
                continue
            short = subprocess.run(["c++filt", "-p", name], capture_output=True, text=True).stdout.strip() or name
            short = re.sub(r"\(.*", "", short).replace("fv::", "")
            print(f"{obj.stem + ':' + short:58.58} " + " ".join(f"{counts[k]:12d}" for k in PATTERNS))


if __name__ == "__main__":
    main()
