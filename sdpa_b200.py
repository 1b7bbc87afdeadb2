"""Import shim: ``import sdpa_b200`` loads the package whose directory name
(``mpi-parallelized-scaled-dot-product-attention-with-avx-512-optimization_b200``) is not a
valid Python identifier."""
import importlib.util
import sys
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent / "mpi-parallelized-scaled-dot-product-attention-with-avx-512-optimization_b200"
_spec = importlib.util.spec_from_file_location(
    "sdpa_b200", _PKG_DIR / "__init__.py", submodule_search_locations=[str(_PKG_DIR)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sdpa_b200"] = _mod
_spec.loader.exec_module(_mod)
