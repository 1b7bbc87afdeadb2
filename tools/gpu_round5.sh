#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary6.log
timeout 600 python tools/umma_diag.py --out $OUT/umma_diag.txt > $OUT/umma_diag.log 2>&1
echo "umma_diag rc=$?" >> $OUT/summary6.log
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -p no:cacheprovider > $OUT/pytest_bf16.log 2>&1
echo "pytest_bf16 rc=$?" >> $OUT/summary6.log
for P in 0 4 8; do
  SDPA_UMMA_POLY=$P timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_c3_poly$P.json 2> $OUT/bench_c3.err
  echo "bench poly=$P rc=$?" >> $OUT/summary6.log
done
SDPA_UMMA_SAFE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_c3_safe.json 2>> $OUT/bench_c3.err
SDPA_UMMA_TRACE=$OUT/trace_c3.txt timeout 300 python tools/profile_target.py --steps 1 > $OUT/trace_run.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attn_umma -s 2 -c 1 -f -o $OUT/prof_umma \
    python tools/profile_target.py --steps 2 > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?" >> $OUT/summary6.log
cat $OUT/summary6.log; grep -v RESULT $OUT/umma_diag.log | tail -5; python - <<'PY'
import json
for l in open('gpurun_out/umma_diag.log'):
    if l.startswith('RESULT'):
        d=json.loads(l[7:]); print(d['m'],d['n'],d['splits'],d['q_batch'],'err',round(d['max_err'],5),'err_b',round(d['max_err_vs_bf16_inputs'],5),'nan',d['nan'])
for name in ('poly0','poly4','poly8','safe'):
    try:
        d=json.load(open(f'gpurun_out/bench_c3_{name}.json')); print(name,'value',round(d['value'],1),'roofline',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],4),'fused_ms',round(d['stage_ms_per_step']['fused'],4))
    except Exception as e: print(name,'ERR',e)
PY
tail -6 $OUT/pytest_bf16.log; tail -3 $OUT/bench_c3.err
