#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary13.log $OUT/sweep10.txt $OUT/sweep10.err
export SDPA_UMMA_V7=1
timeout 300 python tools/umma_diag.py --out $OUT/umma_diag_v7g2.txt > $OUT/umma_diag_v7g2.log 2>&1
echo "umma_diag v7 groups2 rc=$?" >> $OUT/summary13.log
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -p no:cacheprovider > $OUT/pytest_bf16_v7g2.log 2>&1
echo "pytest_bf16 v7 groups2 rc=$?" >> $OUT/summary13.log
run() { label=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > $OUT/bench_$label.json 2>> $OUT/sweep10.err
  python - "$label" <<'PY' >> gpurun_out/sweep10.txt
import json,sys
try:
    d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json')); print(sys.argv[1],'value',round(d['value'],1),'fused TF',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],4),'fused_ms',round(d['stage_ms_per_step']['fused'],4),'step_ms',round(d['ms_per_step'],4), d['self_check'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
run v7g2_poly4 SDPA_UMMA_POLY=4
run v7g2_poly0 SDPA_UMMA_POLY=0
run v7g2_poly8 SDPA_UMMA_POLY=8
run v7g1_poly4 SDPA_UMMA_GROUPS=1 SDPA_UMMA_POLY=4
for S in 9 16 23; do EXTRA="--kv-splits $S" run v7g2_splits$S SDPA_UMMA_POLY=4; done
EXTRA=""
run v5 SDPA_UMMA_V7=0 SDPA_UMMA_POLY=4
SDPA_UMMA_TRACE=$OUT/trace_v7g2.txt timeout 300 python tools/profile_target.py --steps 1 > $OUT/trace_run.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attn_umma_kernel_v7 -s 1 -c 1 -f -o $OUT/prof_umma_v7g2 \
    python tools/profile_target.py --steps 2 > $OUT/ncu_full_v7g2.log 2>&1
echo "full capture rc=$?" >> $OUT/summary13.log
cat $OUT/summary13.log; grep -c '"nan": 0' $OUT/umma_diag_v7g2.log; grep -v RESULT $OUT/umma_diag_v7g2.log | tail -3; tail -3 $OUT/pytest_bf16_v7g2.log; cat $OUT/sweep10.txt; tail -2 $OUT/sweep10.err
