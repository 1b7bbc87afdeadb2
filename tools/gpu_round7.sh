#!/bin/bash
# v6 bring-up: diag ladder, bf16 tests, bench variants, trace, ncu -- all with SDPA_UMMA_V6=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary9.log $OUT/sweep6.txt
export SDPA_UMMA_V6=1
timeout 600 python tools/umma_diag.py --out $OUT/umma_diag_v6.txt > $OUT/umma_diag_v6.log 2>&1
echo "umma_diag v6 rc=$?" >> $OUT/summary9.log
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -p no:cacheprovider > $OUT/pytest_bf16_v6.log 2>&1
echo "pytest_bf16 v6 rc=$?" >> $OUT/summary9.log
run() { label=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > $OUT/bench_$label.json 2>> $OUT/sweep6.err
  python - "$label" <<'PY' >> gpurun_out/sweep6.txt
import json,sys
try:
    d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json')); print(sys.argv[1],'value',round(d['value'],1),'fused TF',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],4),'fused_ms',round(d['stage_ms_per_step']['fused'],4),'merge_ms',round(d['stage_ms_per_step']['merge'],4),'step_ms',round(d['ms_per_step'],4), d['self_check'])
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
run v6_poly4 SDPA_UMMA_POLY=4
run v6_poly0 SDPA_UMMA_POLY=0
run v6_poly8 SDPA_UMMA_POLY=8
run v5_poly4 SDPA_UMMA_V6=0 SDPA_UMMA_POLY=4
for S in 5 7 9 14; do EXTRA="--kv-splits $S" run v6_splits$S SDPA_UMMA_POLY=4; done
EXTRA=""
SDPA_UMMA_TRACE=$OUT/trace_v6.txt timeout 300 python tools/profile_target.py --steps 1 > $OUT/trace_run.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attn_umma_kernel_v6 -s 1 -c 1 -f -o $OUT/prof_umma_v6 \
    python tools/profile_target.py --steps 2 > $OUT/ncu_full_v6.log 2>&1
echo "full capture rc=$?" >> $OUT/summary9.log
cat $OUT/summary9.log; grep -v RESULT $OUT/umma_diag_v6.log | tail -6; python - <<'PY'
import json
for l in open('gpurun_out/umma_diag_v6.log'):
    if l.startswith('RESULT'):
        d=json.loads(l[7:]); print(d['m'],d['n'],d['splits'],d['q_batch'],'err',round(d['max_err'],5),'err_b',round(d['max_err_vs_bf16_inputs'],5),'nan',d['nan'])
PY
tail -5 $OUT/pytest_bf16_v6.log; cat $OUT/sweep6.txt; tail -3 $OUT/sweep6.err
