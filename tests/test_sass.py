"""The built library really contains what DESIGN.md says it does: tcgen05 MMAs on TMEM accumulators fed by TMA in the fused
kernels, bulk copies in the background cast.  Reads the SASS of the in-tree build with cuobjdump (no GPU needed)."""
import re
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def sass_counts():
    if shutil.which("cuobjdump") is None or shutil.which("cu++filt") is None:
        pytest.skip("cuobjdump / cu++filt not on PATH")
    if not list(ROOT.glob("*_b200/libsdpa_b200.so")):
        pytest.skip("library not built")
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "sass_evidence.py")], capture_output=True, text=True, check=True).stdout
    counts = {}
    lines = out.splitlines()
    for name, body in zip(lines, lines[1:]):
        if body.startswith("    instructions"):
            counts[name.strip()] = {k: int(v) for k, v in re.findall(r"([A-Z0-9_.]+) (\d+)", body.split(":", 1)[1])}
    return counts


def _kernel(counts, fragment):
    hits = {k: v for k, v in counts.items() if fragment in k}
    assert hits, f"no kernel named *{fragment}* in the library"
    return hits


def test_fused_kernels_are_tcgen05_tma_tmem(sass_counts):
    for frag in ("attn_umma_kernel_v8", "attn_umma_kernel_v7", "attn_umma_general_kernel"):
        for name, ops in _kernel(sass_counts, frag).items():
            assert ops.get("UTCHMMA.2CTA", 0) > 0, name      # tcgen05.mma.cta_group::2
            assert ops.get("UTMALDG", 0) > 0, name           # TMA tensor loads
            assert ops.get("LDTM", 0) > 0 and ops.get("STTM", 0) > 0, name   # tcgen05.ld / .st on TMEM
            assert ops.get("UTCBAR", 0) > 0, name            # tcgen05.commit


def test_background_cast_uses_bulk_copies_and_few_registers(sass_counts):
    for name, ops in _kernel(sass_counts, "cvt_in_batch_bg_kernel").items():
        assert ops.get("UBLKCP", 0) > 0, name                # cp.async.bulk into the shared-memory ring
    lib = next(ROOT.glob("*_b200/libsdpa_b200.so"))
    usage = subprocess.run(["cuobjdump", "--dump-resource-usage", str(lib)], capture_output=True, text=True, check=True).stdout
    regs = [int(r) for fn, r in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", usage) if "cvt_in_batch_bg" in fn or "merge_inbox_bg" in fn]
    assert regs and max(regs) <= 32, regs                    # 128 threads x 32 registers fit beside the fused kernel's 640 x 96
