#!/bin/bash
# Round-2 visit 16 (one GPU): timeline of the background cast's CTAs against the fused kernel (tools/cast_trace.py), host issue
# time per queued pass, bench with cast-ahead on / off.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
timeout 300 python tools/cast_trace.py $OUT/cast_trace.txt 2>&1 | tail -5
SDPA_CAST_AHEAD=0 timeout 300 python tools/cast_trace.py $OUT/cast_trace_off.txt 2>&1 | grep "queued passes"
for a in 1 0; do
SDPA_CAST_AHEAD=$a timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extra none > $OUT/v16_ahead$a.json 2>> $OUT/v16_bench.err
SDPA_CAST_AHEAD=$a timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --extra none > $OUT/v16_ahead${a}_k100.json 2>> $OUT/v16_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v16_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "host issue us", round(d["host_issue_us_per_step"],1), "fused", round(d["roofline"]["achieved"],1), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["parity_check"]["ok"], d["clocks"].get("sm_mhz"))
    except Exception as e:
        print(f, "unreadable", e)
PY
