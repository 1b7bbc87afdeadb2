"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys
the driver reads (it runs the reference's own CPU program -- oracle/_ref -- or the oracle port)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.timeout(600)
def test_reference_arm_json_line(oracle):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=580, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "attention_tflops" and d["unit"] == "TFLOP/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "rows" in cb["sample"]
    assert d["config"]["m"] == 8192 and d["config"]["n"] == 65536


def test_our_arm_refuses_without_gpu(sdpa):
    if sdpa.device_count() > 0:
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_clock_sampler_summary_and_graceful_start():
    """bench.py's clock sampler: idle samples are dropped, the median / reasons / power come from the loaded ones;
    without NVML or nvidia-smi (this container) start() returns an inert sampler instead of failing."""
    sys.path.insert(0, str(ROOT))
    import bench
    s = bench.ClockSampler(0, None, allow_subprocess=False)
    s.rows = [(1965.0, 1965.0, 120.0, []), (1800.0, 1965.0, 900.0, ["sw_power_cap"]), (1780.0, 1965.0, 950.0, ["sw_power_cap"]),
              (1770.0, 1965.0, 910.0, [])]
    out = s.summary()
    assert out["sm_mhz"] == 1780.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"] and out["power_w_max"] == 950.0
    live = bench.ClockSampler(0, None, allow_subprocess=False).start()
    with live:
        pass
    live.close()
    assert live.summary()["samples"] >= 0
