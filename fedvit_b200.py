"""Import shim: ``import fedvit_b200`` loads the package that lives in the directory
``federated-vit-skin-lesion-classification_b200/`` (a name Python cannot import directly)."""
import importlib.util
import sys
from pathlib import Path

_dir = Path(__file__).resolve().parent / "federated-vit-skin-lesion-classification_b200"
_spec = importlib.util.spec_from_file_location(
    "fedvit_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fedvit_b200"] = _mod
_spec.loader.exec_module(_mod)
