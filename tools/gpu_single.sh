#!/bin/bash
# Single-GPU visit: full GPU suite, smoke, bench (queued passes) for c3/c2, launch list of the bench command.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary_single.log
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=6 > $OUT/pytest_all.log 2>&1
echo "pytest_all rc=$?" >> $OUT/summary_single.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/summary_single.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err
echo "bench c3 rc=$?" >> $OUT/summary_single.log
timeout 900 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > $OUT/bench_c3_200.json 2>> $OUT/bench_c3.err
timeout 600 python bench.py --config c2 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c2.json 2> $OUT/bench_c2.err
SDPA_HOST_PROFILE=1 timeout 300 python tools/profile_target.py --steps 3 > $OUT/host_profile.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_bench_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_launches.log 2>&1
echo "bench launch list rc=$?" >> $OUT/summary_single.log
cat $OUT/summary_single.log; tail -12 $OUT/pytest_all.log; tail -1 $OUT/smoke.log; cat $OUT/bench_c3.json; echo; python -c "
import json
for f in ('bench_c3_200','bench_c2'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f,'value',round(d['value'],1),'fused',round(d['roofline']['achieved'],1),d['roofline']['frac'],'step_ms',d['ms_per_step'],d['stage_ms_per_step'],d['clocks'])"
tail -4 $OUT/host_profile.log
