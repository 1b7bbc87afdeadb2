#!/bin/bash
# Round-2 visit 21 (one GPU): final sanity of the default build -- bench line (c3 + c2, alone leg, CPU baseline) and smoke().
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 170 python bench.py --steps 20 --warmup 5 > gpurun_out/v21_bench_c3.json 2> gpurun_out/v21_bench.err; echo "bench rc=$?"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/v21_bench_c3.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "fused", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["parity_check"]["ok"], "cpu", d.get("cpu_baseline",{}).get("value"), d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
print("alone", d["roofline"].get("alone"))
for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), v["kernel"], v["parity_check"]["max_abs_err"])
PY
