"""The barrier protocol of the experimental persistent kernel (attn_umma_kernel_v8), checked on the CPU by the
discrete-event model in tools/v8_protocol_sim.py: no deadlock, no wait passing on a stale phase, every buffer
hand-over intact -- and the model does notice when a hand-over is removed."""
import importlib.util
import random
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _load():
    spec = importlib.util.spec_from_file_location("v8_protocol_sim", ROOT / "tools" / "v8_protocol_sim.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_protocol_model_passes_on_random_ranges():
    sim = _load()
    rng = random.Random(3)
    ran = 0
    for T, RB, C in [(1, 5, 2), (2, 3, 2), (5, 4, 3), (13, 2, 5), (40, 3, 2), (8, 6, 5)]:
        W = RB * T
        for c in range(C):
            begin, end = c * W // C, (c + 1) * W // C
            if end > begin:
                sim.Sim(begin, end, T, rng).run()
                ran += 1
    assert ran >= 15


def test_protocol_model_detects_a_missing_s_free_wait(monkeypatch):
    sim = _load()
    src = (ROOT / "tools" / "v8_protocol_sim.py").read_text()
    needle = '                yield from self.wait(f"s_free{g & 1}", 0, (g >> 1) & 1, (g >> 1) + 1)'
    assert needle in src
    broken = {}
    exec(compile(src.replace(needle, "                pass", 1), "v8_protocol_sim_broken", "exec"), broken)
    with pytest.raises((AssertionError, RuntimeError)):
        for seed in range(5):
            broken["Sim"](0, 12, 40, random.Random(seed)).run()
