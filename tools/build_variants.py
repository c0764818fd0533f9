#!/usr/bin/env python
"""Build A/B variants of libfedvit.so: one source file recompiled with extra -D flags, linked against the
objects of the regular build. Variants land in <package>/build/variants/ (git-ignored, shipped to the GPU
box); run a tool against one with FEDVIT_LIB=<path>.

    python tools/build_variants.py attention_bwd2.cu name1:-DX=0 name2:-DX=1,-DY=2
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "federated-vit-skin-lesion-classification_b200"
sys.path.insert(0, str(PKG))
import build as fv_build  # noqa: E402


def main():
    src = PKG / "csrc" / sys.argv[1]
    fv_build.build()
    out_dir = PKG / "build" / "variants"
    out_dir.mkdir(parents=True, exist_ok=True)
    others = [str(PKG / "build" / (s.stem + ".o")) for s in sorted((PKG / "csrc").glob("*.cu")) if s != src]
    for spec in sys.argv[2:]:
        name, flags = spec.split(":", 1)
        obj = out_dir / f"{src.stem}_{name}.o"
        cmd = [fv_build.NVCC, *fv_build.FLAGS, *[f for f in flags.split(",") if f], "-c", str(src), "-o", str(obj)]
        subprocess.run(cmd, check=True)
        lib = out_dir / f"libfedvit_{name}.so"
        subprocess.run([fv_build.NVCC, "-shared", "-o", str(lib), str(obj), *others, "-lcudart"], check=True)
        print(lib)


if __name__ == "__main__":
    main()
