#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary4.log
SDPA_UMMA_TRACE=$OUT/trace_c3.txt timeout 300 python tools/profile_target.py --steps 1 > $OUT/trace_run.log 2>&1
echo "trace rc=$?" >> $OUT/summary4.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/pytest_all.log 2>&1
echo "pytest_all rc=$?" >> $OUT/summary4.log
timeout 600 python bench.py --config c2 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_f32.json 2> $OUT/bench_c2.err
timeout 600 python bench.py --config c2 --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_bf16.json 2>> $OUT/bench_c2.err
echo "bench_c2 rc=$?" >> $OUT/summary4.log
cat $OUT/summary4.log; tail -3 $OUT/pytest_all.log; head -c 600 $OUT/bench_c2_f32.json; echo; head -c 600 $OUT/bench_c2_bf16.json; echo; wc -l $OUT/trace_c3.txt
