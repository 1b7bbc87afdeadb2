#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary10.log $OUT/sweep7.txt $OUT/sweep7.err
export SDPA_UMMA_V6=1
SDPA_UMMA_PARTS=4 timeout 600 python tools/umma_diag.py --out $OUT/umma_diag_v6p4.txt > $OUT/umma_diag_v6p4.log 2>&1
echo "umma_diag v6 parts4 rc=$?" >> $OUT/summary10.log
SDPA_UMMA_PARTS=4 timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -p no:cacheprovider > $OUT/pytest_bf16_v6p4.log 2>&1
echo "pytest_bf16 v6 parts4 rc=$?" >> $OUT/summary10.log
SDPA_UMMA_PARTS=2 timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -p no:cacheprovider > $OUT/pytest_bf16_v6p2.log 2>&1
echo "pytest_bf16 v6 parts2 rc=$?" >> $OUT/summary10.log
run() { label=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_$label.json 2>> $OUT/sweep7.err
  python - "$label" <<'PY' >> gpurun_out/sweep7.txt
import json,sys
try:
    d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json')); print(sys.argv[1],'value',round(d['value'],1),'fused TF',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],4),'fused_ms',round(d['stage_ms_per_step']['fused'],4),'step_ms',round(d['ms_per_step'],4), d['self_check'], d['clocks'].get('sm_mhz'), d['clocks'].get('samples'))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
run v6p2_poly4 SDPA_UMMA_PARTS=2 SDPA_UMMA_POLY=4
run v6p2_poly8 SDPA_UMMA_PARTS=2 SDPA_UMMA_POLY=8
run v6p4_poly0 SDPA_UMMA_PARTS=4 SDPA_UMMA_POLY=0
run v6p4_poly4 SDPA_UMMA_PARTS=4 SDPA_UMMA_POLY=4
run v6p4_poly8 SDPA_UMMA_PARTS=4 SDPA_UMMA_POLY=8
run v5_poly4 SDPA_UMMA_V6=0 SDPA_UMMA_POLY=4
SDPA_UMMA_PARTS=4 SDPA_UMMA_TRACE=$OUT/trace_v6p4.txt timeout 300 python tools/profile_target.py --steps 1 > $OUT/trace_run.log 2>&1
SDPA_UMMA_PARTS=2 SDPA_UMMA_TRACE=$OUT/trace_v6p2.txt timeout 300 python tools/profile_target.py --steps 1 >> $OUT/trace_run.log 2>&1
cat $OUT/summary10.log; grep -c RESULT $OUT/umma_diag_v6p4.log; grep -v RESULT $OUT/umma_diag_v6p4.log | tail -4; tail -3 $OUT/pytest_bf16_v6p4.log; tail -2 $OUT/pytest_bf16_v6p2.log; cat $OUT/sweep7.txt; tail -3 $OUT/sweep7.err
