#!/bin/bash
# Round-2 visit 14 (one GPU): cast-ahead -- the casts of queued pass i+1 as a small-footprint side-stream kernel beside the
# persistent fused kernel of pass i (SDPA_CAST_AHEAD=1, default) against the in-stream casts (=0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v14.log; rm -f $S
timeout 600 python -m pytest tests/test_gpu_bf16.py -q -m gpu -p no:cacheprovider -x -k "cast_ahead or persistent" > $OUT/v14_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $S
if [ $rc -ne 0 ]; then cat $S; tail -40 $OUT/v14_pytest.log; exit 1; fi
for a in 1 0 1 0; do
SDPA_CAST_AHEAD=$a timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extra none > $OUT/v14_ahead${a}_$RANDOM.json 2>> $OUT/v14_bench.err
echo "bench ahead=$a rc=$?" >> $S
done
SDPA_CAST_AHEAD=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extra none --stage-timing-every 1 > $OUT/v14_ahead1_allmarks.json 2>> $OUT/v14_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/v14_bench_full.json 2>> $OUT/v14_bench.err
echo "bench full rc=$?" >> $S
cat $S; tail -3 $OUT/v14_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v14_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "fused", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], d["parity_check"]["ok"], d["parity_check"]["max_abs_err"], d["clocks"].get("sm_mhz"))
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), v["stage_ms_per_step"], v["kernel"], v["parity_check"]["max_abs_err"])
    except Exception as e:
        print(f, "unreadable", e)
PY
