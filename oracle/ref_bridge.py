"""Import the UNMODIFIED reference modules (model.py, losses.py, utils.py) from /root/reference,
with oracle/timm answering the one third-party import they need (model.py:17 ``import timm``).

TEST INFRASTRUCTURE (see oracle/__init__.py). Works only where /root/reference is mounted — the
build container. Used by tests/golden/make_golden.py to produce the committed fixtures and by
tests/test_oracle.py to pin the restatements in isic.py / asl.py / step.py against the real thing;
everything that runs on the GPU box uses the fixtures instead.
"""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path
from types import ModuleType

REFERENCE_DIR = Path("/root/reference")
_ORACLE_DIR = Path(__file__).resolve().parent


def available() -> bool:
    return (REFERENCE_DIR / "model.py").exists()


def _load(name: str) -> ModuleType:
    alias = f"_reference_{name}"
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, REFERENCE_DIR / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns (model, losses, utils) modules of the reference."""
    if not available():
        raise RuntimeError("/root/reference is not mounted here")
    if "timm" not in sys.modules:
        if str(_ORACLE_DIR) not in sys.path:
            sys.path.insert(0, str(_ORACLE_DIR))
        import timm  # noqa: F401  (oracle/timm)
    return _load("model"), _load("losses"), _load("utils")
