"""ctypes binding of libfedvit.so — the C-ABI boundary of the hot path.

The prototypes are read from ``include/fedvit.h`` itself, so the Python side can never drift from
the header: every ``fv_*`` function declared there gets ``argtypes`` / ``restype`` set here, and
``tests/test_abi.py`` checks that the shared object exports each one.

There is deliberately no fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path
from typing import Dict, List, Tuple

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
HEADER = REPO_ROOT / "include" / "fedvit.h"
LIB_PATH = PKG_DIR / "libfedvit.so"

_CTYPE = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "size_t": ctypes.c_size_t,
    "float": ctypes.c_float,
}


def parse_header(path: Path = HEADER) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """Return {name: (return_type, [(ctype_string, arg_name), ...])} for every fv_* prototype."""
    text = path.read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"^\s*#[^\n]*", " ", text, flags=re.M)
    protos: Dict[str, Tuple[str, List[Tuple[str, str]]]] = {}
    for m in re.finditer(r"((?:const\s+)?(?:int64_t|int|char|void|float)\s*\*?)\s*\b(fv_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        parsed: List[Tuple[str, str]] = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                parsed.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, parsed)
    return protos


def _to_ctype(t: str):
    t = t.replace("const ", "").strip()
    if t.endswith("*"):
        return ctypes.c_char_p if t == "char*" else ctypes.c_void_p
    return _CTYPE[t]


class FedVitError(RuntimeError):
    pass


class _Lib:
    def __init__(self) -> None:
        self._dll = None
        self._fns = {}

    def load(self):
        if self._dll is not None:
            return self._dll
        path = Path(os.environ.get("FEDVIT_LIB", LIB_PATH))
        if not path.exists():
            raise FedVitError(
                f"{path} not found — build it with `python {PKG_DIR.name}/build.py` "
                "(nvcc, sm_100a). There is no CPU or eager fallback for this path."
            )
        dll = ctypes.CDLL(str(path))
        for name, (ret, args) in parse_header().items():
            fn = getattr(dll, name)  # AttributeError here == header/library mismatch
            fn.argtypes = [_to_ctype(t) for t, _ in args]
            fn.restype = _to_ctype(ret)
            self._fns[name] = fn
        self._dll = dll
        return dll

    def call(self, name: str, *args):
        if self._dll is None:
            self.load()
        trace = TRACE
        if trace is not None:  # tools/step_breakdown.py: live per-launch device times
            import torch

            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = self._fns[name](*args)
            e1.record()
            trace.append((name, args, e0, e1))
        else:
            rc = self._fns[name](*args)
        if rc != 0:
            msg = self._fns["fv_last_error"]()
            raise FedVitError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")

    def raw(self, name: str):
        if self._dll is None:
            self.load()
        return self._fns[name]


# when set to a list, every C-ABI call is bracketed by CUDA events on the current stream and
# (name, args, start, end) is appended (measurement tooling only)
TRACE = None

LIB = _Lib()


def launch_count() -> int:
    return int(LIB.raw("fv_launch_count")())


# kernel families of fv_kernel_launches (include/fedvit.h: FV_KERNEL_*)
KERNEL_GEMM_TC, KERNEL_GEMM_TC_PAIR, KERNEL_ATTN_FWD, KERNEL_ATTN_BWD = 0, 1, 2, 3
KERNEL_ATTN_FWD_LONG, KERNEL_ATTN_BWD_LONG, KERNEL_ATTN_LEGACY = 4, 5, 6


def kernel_launches(family: int) -> int:
    """How many kernels of one family the library has launched since load: lets a test assert which
    kernel served a call (e.g. that a ViT-B step ran on the CTA-pair GEMM)."""
    return int(LIB.raw("fv_kernel_launches")(family))
