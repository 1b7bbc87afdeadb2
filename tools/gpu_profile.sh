#!/bin/bash
# ncu evidence for the fused kernel: launch list of a short bench run + one full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
timeout 600 python tools/profile_target.py --steps 3 > $OUT/target_plain.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c3.csv \
    python tools/profile_target.py --steps 3 > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?" >> $OUT/summary2.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attn_umma -s 1 -c 1 -f -o $OUT/prof_umma \
    python tools/profile_target.py --steps 2 > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?" >> $OUT/summary2.log
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -p no:cacheprovider > $OUT/pytest_bf16.log 2>&1
echo "pytest_bf16 rc=$?" >> $OUT/summary2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err
echo "bench_c3 rc=$?" >> $OUT/summary2.log
cat $OUT/summary2.log; cat $OUT/target_plain.log | tail -2; tail -3 $OUT/pytest_bf16.log; cat $OUT/bench_c3.json; ls -la $OUT
