#!/bin/bash
# ncu evidence for the fused kernel (one GPU): launch list of the bench command, one --set full capture of the
# default kernel with source, and a clock64 timeline of CTA (0,0).  Digest here with tools/ncu_digest.py / trace_digest.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_bench_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_launches.log 2>&1
echo "bench launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attn_umma_kernel_v7 -s 1 -c 1 -f -o $OUT/prof_umma \
    python tools/profile_target.py --steps 2 > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?"
SDPA_UMMA_TRACE=$OUT/trace.txt timeout 300 python tools/profile_target.py --steps 1 > $OUT/trace_run.log 2>&1
echo "timeline rc=$?"
