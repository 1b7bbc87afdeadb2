#!/bin/bash
# Round-2 visit 10 (one GPU): suite + bench after pruning the kernel variants, timing the guard twin outside the fused stage and
# preloading the o vectors in the merge kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v10.log; rm -f $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/v10_pytest.log 2>&1
echo "pytest rc=$?" >> $S
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/v10_bench_c3.json 2> $OUT/v10_bench.err
echo "bench rc=$?" >> $S
SDPA_DEFER_TWIN=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extra none > $OUT/v10_bench_c3_allmarks.json 2>> $OUT/v10_bench.err
echo "bench allmarks rc=$?" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/v10_smoke.log 2>&1
echo "smoke rc=$?" >> $S
cat $S; grep -E "passed|failed" $OUT/v10_pytest.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/v10_pytest.log | head -20; tail -1 $OUT/v10_smoke.log
python - <<'PY'
import json
for f in ("v10_bench_c3","v10_bench_c3_allmarks"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "fused", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], d["parity_check"]["ok"], "launches", d["roofline"]["launches"])
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), v["stage_ms_per_step"], v["kernel"], v["parity_check"]["max_abs_err"])
    except Exception as e:
        print(f, "unreadable", e)
PY
