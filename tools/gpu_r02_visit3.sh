#!/bin/bash
# Round-2 visit 3 (one GPU): v5 removed (exact twin = general kernel), v8 default + its variants (V producer warp, progressive P
# stores, polynomial exponentials), full suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v3.log; rm -f $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=5 > $OUT/v3_pytest.log 2>&1
echo "pytest rc=$?" >> $S
for V in "0 0" "1 0" "2 0" "3 0" "3 2" "3 4" "3 6" "0 4"; do
  set -- $V
  SDPA_V8_OPT=$1 SDPA_UMMA_POLY=$2 timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --extra none > $OUT/v3_bench_opt$1_poly$2.json 2>> $OUT/v3_bench.err
  echo "bench opt=$1 poly=$2 rc=$?" >> $S
done
SDPA_UMMA_V8=0 timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --extra none > $OUT/v3_bench_v7.json 2>> $OUT/v3_bench.err
echo "bench v7 rc=$?" >> $S
cat $S; grep -E "passed|failed" $OUT/v3_pytest.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/v3_pytest.log | head -20
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v3_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "fused", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], "ok", d["parity_check"]["ok"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "unreadable", e)
PY
