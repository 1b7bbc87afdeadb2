#!/bin/bash
# Round-2 visit 7 (one GPU): suite after the merge/pool/mark changes, bench with sampled vs full stage marks, harness modes after the
# runtime prewarm, ncu captures of the split-precision kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v7.log; rm -f $S $OUT/harness_staging.txt
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/v7_pytest.log 2>&1
echo "pytest rc=$?" >> $S
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/v7_bench_c3.json 2> $OUT/v7_bench.err
echo "bench rc=$?" >> $S
timeout 300 python bench.py --steps 20 --warmup 5 --stage-timing-every 1 --no-cpu-baseline --extra none > $OUT/v7_bench_c3_allmarks.json 2>> $OUT/v7_bench.err
echo "bench allmarks rc=$?" >> $S
M=8192 N=65536 timeout 400 bash tools/gpu_harness_staging.sh > $OUT/v7_harness.log 2>&1
echo "harness rc=$?" >> $S
cap() { local name=$1 skip=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_umma_general_kernel -s $skip -c 1 -f -o $OUT/r02_prof_$name \
      python tools/profile_target.py "$@" > $OUT/r02_ncu_$name.log 2>&1; echo "capture $name rc=$?" >> $S; }
cap x3_c3 2 --steps 2 --precision bf16x3
cap x3_c2 2 --steps 2 --precision bf16x3 --m 4096 --n 4096
cat $S; grep -E "passed|failed" $OUT/v7_pytest.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/v7_pytest.log | head -20
python - <<'PY'
import json
for f in ("v7_bench_c3","v7_bench_c3_allmarks"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "fused", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], d["parity_check"]["ok"], "launches", d["roofline"]["launches"])
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), v["stage_ms_per_step"], v["kernel"], v["parity_check"]["max_abs_err"])
    except Exception as e:
        print(f, "unreadable", e)
PY
cat $OUT/harness_staging.txt
