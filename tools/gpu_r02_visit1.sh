#!/bin/bash
# Round-2 visit 1 (one GPU): the full GPU suite incl. the gated experimental tests, pipe micro-benchmarks, the new bench.py
# (parity_check + c2 extra + CPU arm), the persistent kernel (v8) against the default, harness staging modes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v1.log; rm -f $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/v1_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> $OUT/v1_gpu.txt
SDPA_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x --durations=5 > $OUT/v1_pytest.log 2>&1
echo "pytest(experimental on) rc=$?" >> $S
timeout 120 tools/ubench/pipes > $OUT/ubench_pipes.txt 2>&1
echo "ubench rc=$?" >> $S
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/v1_bench_c3.json 2> $OUT/v1_bench_c3.err
echo "bench c3 rc=$?" >> $S
SDPA_UMMA_V8=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --extra none > $OUT/v1_bench_c3_v8.json 2> $OUT/v1_bench_c3_v8.err
echo "bench c3 v8 rc=$?" >> $S
for P in 0 4 8; do
  SDPA_UMMA_POLY=$P timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --extra none > $OUT/v1_bench_c3_poly$P.json 2>> $OUT/v1_bench_c3.err
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/v1_bench_ref.json 2> $OUT/v1_bench_ref.err
echo "bench ref rc=$?" >> $S
M=8192 N=65536 timeout 400 bash tools/gpu_harness_staging.sh > $OUT/v1_harness.log 2>&1
echo "harness rc=$?" >> $S
cat $S; tail -5 $OUT/v1_pytest.log; cat $OUT/ubench_pipes.txt | head -70
python - <<'PY'
import json
for f in ("v1_bench_c3","v1_bench_c3_v8","v1_bench_c3_poly0","v1_bench_c3_poly4","v1_bench_c3_poly8","v1_bench_ref"):
    try:
        d=json.loads(open("gpurun_out/"+f+".json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],2), "ms", round(d["ms_per_step"],4), "fused", d.get("roofline",{}).get("achieved"), d.get("stage_ms_per_step"), d.get("parity_check"), d.get("impl_detail",{}).get("kernel"))
        if "configs" in d: print("   configs", json.dumps(d["configs"])[:1200])
        if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -12 $OUT/harness_staging.txt
