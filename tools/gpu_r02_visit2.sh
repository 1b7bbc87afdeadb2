#!/bin/bash
# Round-2 visit 2 (one GPU): the general / split-precision kernel (attn_umma_general.cu) for the first time on hardware.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v2.log; rm -f $S
timeout 900 python -m pytest tests/test_gpu_general.py -q -m gpu -p no:cacheprovider --durations=5 > $OUT/v2_pytest_general.log 2>&1
echo "pytest general rc=$?" >> $S
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_general.py > $OUT/v2_pytest_rest.log 2>&1
echo "pytest rest rc=$?" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/v2_smoke.log 2>&1
echo "smoke rc=$?" >> $S
timeout 300 python bench.py --config c2 --steps 20 --warmup 3 --no-cpu-baseline --extra none > $OUT/v2_bench_c2_x3.json 2> $OUT/v2_bench_c2.err
echo "bench c2 x3 rc=$?" >> $S
timeout 300 python bench.py --config c3 --precision bf16x3 --steps 10 --warmup 3 --no-cpu-baseline --extra none > $OUT/v2_bench_c3_x3.json 2>> $OUT/v2_bench_c2.err
echo "bench c3 x3 rc=$?" >> $S
SDPA_UMMA_GENERAL=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --extra none > $OUT/v2_bench_c3_general.json 2>> $OUT/v2_bench_c2.err
echo "bench c3 general rc=$?" >> $S
cat $S; grep -E "passed|failed|error" $OUT/v2_pytest_general.log | tail -3; grep -E "^FAILED|^ERROR" $OUT/v2_pytest_general.log | head -40
grep -E "passed|failed" $OUT/v2_pytest_rest.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/v2_pytest_rest.log | head; tail -2 $OUT/v2_smoke.log
grep -h "sdpa_b200: mbarrier timeout" $OUT/v2_pytest_general.log | sort | uniq -c | head
python - <<'PY'
import json
for f in ("v2_bench_c2_x3","v2_bench_c3_x3","v2_bench_c3_general"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],2), "ms", round(d["ms_per_step"],4), "fused", round(d["roofline"]["achieved"],1), d["stage_ms_per_step"], "err", d["parity_check"]["max_abs_err"], d["impl_detail"]["kernel"])
    except Exception as e:
        print(f, "unreadable", e)
PY
