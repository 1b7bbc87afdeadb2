"""Build the CUDA library in-tree (``libsdpa_b200.so`` next to this file).

``nvcc -gencode arch=compute_100a,code=sm_100a`` cross-compiles without a GPU; the
built ``.so`` is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libsdpa_b200.so"
HARNESS_PATH = PKG_DIR / "attention_b200"


def sources_newer_than_lib() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    srcs = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.c"))
    srcs.append(PKG_DIR.parent / "include" / "sdpa_b200.h")
    return any(p.stat().st_mtime > t for p in srcs if p.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Run ``make`` in csrc/ (no-op when the library is up to date)."""
    if force:
        subprocess.run(["make", "-C", str(CSRC), "clean"], check=True, stdout=subprocess.DEVNULL)
    env = dict(os.environ)
    env.setdefault("PATH", "")
    if "/usr/local/cuda/bin" not in env["PATH"]:
        env["PATH"] = "/usr/local/cuda/bin:" + env["PATH"]
    res = subprocess.run(["make", "-C", str(CSRC), "-j8", "all"], env=env, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libsdpa_b200.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose=True))
