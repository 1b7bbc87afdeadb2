#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary8.log
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/pytest_all.log 2>&1
echo "pytest_all rc=$?" >> $OUT/summary8.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/summary8.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/bench_ref_c3.json 2> $OUT/bench_ref.err
echo "bench ref rc=$?" >> $OUT/summary8.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err
echo "bench c3 rc=$?" >> $OUT/summary8.log
timeout 600 python bench.py --config c2 --steps 20 --warmup 3 > $OUT/bench_c2.json 2> $OUT/bench_c2.err
echo "bench c2 rc=$?" >> $OUT/summary8.log
cat $OUT/summary8.log; tail -3 $OUT/pytest_all.log; tail -1 $OUT/smoke.log; cat $OUT/bench_ref_c3.json | head -c 700; echo; cat $OUT/bench_c3.json; echo; head -c 900 $OUT/bench_c2.json
