import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()  # builds oracle/_build/liboracle.so on first use
    return o


@pytest.fixture(scope="session")
def golden():
    meta = json.loads((GOLDEN / "cases.json").read_text())
    data = np.load(GOLDEN / "reference_outputs.npz")
    return meta, data


@pytest.fixture(scope="session")
def sdpa():
    """The host mirror; builds the CUDA library in-tree if a source is newer than the .so."""
    import sdpa_b200
    if sdpa_b200.build_mod.sources_newer_than_lib():
        sdpa_b200.build_mod.build()
    sdpa_b200.lib()
    return sdpa_b200


def has_gpu() -> bool:
    try:
        import sdpa_b200
        return sdpa_b200.device_count() > 0
    except Exception:
        return False
