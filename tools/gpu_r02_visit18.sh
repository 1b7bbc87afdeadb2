#!/bin/bash
# Round-2 visit 18 (one GPU): cast-ahead with the per-pass fence as default -- suite, bench (with the `alone` leg), fence variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v18.log; rm -f $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/v18_pytest.log 2>&1
echo "pytest rc=$?" >> $S
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/v18_bench_c3.json 2> $OUT/v18_bench.err
echo "bench rc=$?" >> $S
SDPA_PASS_FENCE=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extra none --no-alone > $OUT/v18_bench_fence2.json 2>> $OUT/v18_bench.err
SDPA_PASS_FENCE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extra none --no-alone > $OUT/v18_bench_fence0.json 2>> $OUT/v18_bench.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --extra none --no-alone --stage-timing-every 10 > $OUT/v18_bench_k100.json 2>> $OUT/v18_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/v18_smoke.log 2>&1
echo "smoke rc=$?" >> $S
cat $S; grep -E "passed|failed" $OUT/v18_pytest.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/v18_pytest.log | head -20; tail -1 $OUT/v18_smoke.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v18_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "issue", round(d["host_issue_us_per_step"],1), "fused", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], d["parity_check"]["ok"], d["clocks"].get("sm_mhz"))
        if d["roofline"].get("alone"): print("     alone", d["roofline"]["alone"])
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), v["stage_ms_per_step"], v["kernel"], v["parity_check"]["max_abs_err"])
    except Exception as e:
        print(f, "unreadable", e)
PY
