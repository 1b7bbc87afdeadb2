#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary7.log $OUT/sweep.txt
run() { # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > $OUT/bench_$label.json 2>> $OUT/sweep.err
  python - "$label" <<'PY' >> gpurun_out/sweep.txt
import json,sys
try:
    d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json')); print(sys.argv[1],'value',round(d['value'],1),'fused TF',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],4),'fused_ms',round(d['stage_ms_per_step']['fused'],4),'merge_ms',round(d['stage_ms_per_step']['merge'],4),'step_ms',round(d['ms_per_step'],4))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
run chunk1_poly4 SDPA_UMMA_CHUNK=1 SDPA_UMMA_POLY=4
run chunk0_poly4 SDPA_UMMA_CHUNK=0 SDPA_UMMA_POLY=4
run chunk0_poly0 SDPA_UMMA_CHUNK=0 SDPA_UMMA_POLY=0
run chunk1_poly0 SDPA_UMMA_CHUNK=1 SDPA_UMMA_POLY=0
for S in 4 5 9 14 18; do EXTRA="--kv-splits $S" run splits$S SDPA_UMMA_CHUNK=1; done
EXTRA=""
timeout 600 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -p no:cacheprovider > $OUT/pytest_bf16.log 2>&1
echo "pytest_bf16 rc=$?" >> $OUT/summary7.log
SDPA_UMMA_CHUNK=0 timeout 600 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -p no:cacheprovider -k "vs_oracle or guard or golden" > $OUT/pytest_bf16_chunk0.log 2>&1
echo "pytest_bf16 chunk0 rc=$?" >> $OUT/summary7.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_bench_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_launches.log 2>&1
echo "bench launch list rc=$?" >> $OUT/summary7.log
cat $OUT/summary7.log $OUT/sweep.txt; tail -2 $OUT/pytest_bf16.log $OUT/pytest_bf16_chunk0.log
