#!/bin/bash
# Round-2 visit 5 (one GPU): full suite after the flat merge kernels, bench, then the ncu evidence of tools/gpu_r02_profile.sh.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v5.log; rm -f $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/v5_pytest.log 2>&1
echo "pytest rc=$?" >> $S
timeout 300 python bench.py --steps 30 --warmup 5 > $OUT/v5_bench_c3.json 2> $OUT/v5_bench.err
echo "bench rc=$?" >> $S
SDPA_MERGE_FLAT_OFF=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --extra none > $OUT/v5_bench_c3_noflat.json 2>> $OUT/v5_bench.err
echo "bench noflat rc=$?" >> $S
bash tools/gpu_r02_profile.sh > $OUT/v5_profile.log 2>&1
echo "profile rc=$?" >> $S
cat $S; grep -E "passed|failed" $OUT/v5_pytest.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/v5_pytest.log | head -20; cat $OUT/v5_profile.log | tail -12
python - <<'PY'
import json
for f in ("v5_bench_c3","v5_bench_c3_noflat"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "fused", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], d["parity_check"]["ok"])
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), v["stage_ms_per_step"], v["kernel"], v["parity_check"]["max_abs_err"])
        if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("seconds"))
    except Exception as e:
        print(f, "unreadable", e)
PY
