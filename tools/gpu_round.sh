#!/bin/bash
# One GPU-box visit: probes, parity tests, diagnostics, bench.  Everything is time-boxed and
# logs under gpurun_out/ so a failing stage still leaves evidence.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
{
  echo "== host"; nproc; lscpu | grep -E "Model name|Socket|Core|Thread|avx512f" | head; grep -c avx512f /proc/cpuinfo
  echo "== gpu"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv
  ls oracle/_ref oracle/_build 2>&1
} > $OUT/probe.log 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_bf16.py -p no:cacheprovider > $OUT/pytest_f32.log 2>&1
echo "pytest_f32 rc=$?" >> $OUT/summary.log
timeout 600 python tools/umma_diag.py --out $OUT/umma_diag.txt > $OUT/umma_diag.log 2>&1
echo "umma_diag rc=$?" >> $OUT/summary.log
timeout 900 python -m pytest tests/test_gpu_bf16.py -q -m gpu -x -p no:cacheprovider > $OUT/pytest_bf16.log 2>&1
echo "pytest_bf16 rc=$?" >> $OUT/summary.log
timeout 600 python bench.py --config c2 --steps 10 --warmup 3 > $OUT/bench_c2.json 2> $OUT/bench_c2.err
echo "bench_c2 rc=$?" >> $OUT/summary.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err
echo "bench_c3 rc=$?" >> $OUT/summary.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke rc=$?" >> $OUT/summary.log
cat $OUT/summary.log
tail -5 $OUT/pytest_f32.log; tail -15 $OUT/umma_diag.log; tail -5 $OUT/pytest_bf16.log; cat $OUT/bench_c2.json; tail -3 $OUT/bench_c2.err; cat $OUT/bench_c3.json; tail -3 $OUT/bench_c3.err
