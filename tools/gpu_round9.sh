#!/bin/bash
# v7 (2-CTA MMA) bring-up + host-side profile + ncu for casts / f32 kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary11.log $OUT/sweep8.txt $OUT/sweep8.err
SDPA_UMMA_V7=1 timeout 300 python tools/umma_diag.py --out $OUT/umma_diag_v7.txt > $OUT/umma_diag_v7.log 2>&1
echo "umma_diag v7 rc=$?" >> $OUT/summary11.log
run() { label=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_$label.json 2>> $OUT/sweep8.err
  python - "$label" <<'PY' >> gpurun_out/sweep8.txt
import json,sys
try:
    d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json')); print(sys.argv[1],'value',round(d['value'],1),'fused TF',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],4),'fused_ms',round(d['stage_ms_per_step']['fused'],4),'step_ms',round(d['ms_per_step'],4), d['self_check'], d['clocks'].get('sm_mhz'))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
if grep -q '"nan": 0' $OUT/umma_diag_v7.log; then
  run v7_poly4 SDPA_UMMA_V7=1 SDPA_UMMA_POLY=4
  run v7_poly8 SDPA_UMMA_V7=1 SDPA_UMMA_POLY=8
  run v7_parts4 SDPA_UMMA_V7=1 SDPA_UMMA_PARTS=4
  SDPA_UMMA_V7=1 SDPA_UMMA_TRACE=$OUT/trace_v7.txt timeout 300 python tools/profile_target.py --steps 1 > $OUT/trace_run.log 2>&1
fi
run v5 SDPA_UMMA_POLY=4
SDPA_HOST_PROFILE=1 timeout 300 python tools/profile_target.py --steps 6 > $OUT/host_profile.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:cvt_d2bf16 -s 2 -c 2 -f -o $OUT/prof_cast python tools/profile_target.py --steps 2 > $OUT/ncu_cast.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:attn_f32 -s 1 -c 1 -f -o $OUT/prof_f32 python tools/profile_target.py --steps 2 --precision f32 --m 4096 --n 4096 > $OUT/ncu_f32.log 2>&1
echo "ncu rc=$?" >> $OUT/summary11.log
cat $OUT/summary11.log; python - <<'PY'
import json
for l in open('gpurun_out/umma_diag_v7.log'):
    if l.startswith('RESULT'):
        d=json.loads(l[7:]); print(d['m'],d['n'],d['splits'],d['q_batch'],'err',round(d['max_err'],5),'nan',d['nan'], d.get('block_err_32x32',[[]])[0][:4] if 'block_err_32x32' in d else '')
    else: print(l.strip()[:400])
PY
cat $OUT/sweep8.txt; grep "host profile" $OUT/host_profile.log | tail -3; tail -2 $OUT/sweep8.err
