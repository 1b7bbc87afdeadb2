#!/bin/bash
# Round-2 visit 17b (one GPU): cast-ahead without stage marks -- ring depth of the background cast (shared memory beside the fused
# kernel) x a timing-event record behind every pass.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
for st in 5 3 2 1; do
 for f in 0 1; do
  echo "stages=$st fence=$f every=1000: $(env SDPA_CAST_AHEAD=1 SDPA_BG_STAGES=$st SDPA_PASS_FENCE=$f PASSES=200 SDPA_STAGE_TIMING=1 SDPA_STAGE_TIMING_EVERY=1000 NO_TRACE=1 timeout 120 python tools/cast_trace.py $OUT/ct_tmp.txt 2>&1 | grep 'queued passes')"
 done
done
for st in 5 2; do
echo "stages=$st trace, every=1000"; env SDPA_CAST_AHEAD=1 SDPA_BG_STAGES=$st PASSES=60 SDPA_STAGE_TIMING=1 SDPA_STAGE_TIMING_EVERY=1000 NO_TRACE=0 timeout 120 python tools/cast_trace.py $OUT/ct_tmp.txt 2>&1 | tail -3
done
