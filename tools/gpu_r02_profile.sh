#!/bin/bash
# Round-2 ncu evidence (one GPU): launch list of the bench command, --set full captures of the persistent bf16 kernel (v8),
# the plain-grid kernel (v7), the split-precision general kernel on c3 and c2, and the cast kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/r02_launches_bench_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() { # name, kernel regex, env..., -- args
  local name=$1 rx=$2; shift 2
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  local skip=1; for e in "${envs[@]}"; do case $e in SKIP=*) skip=${e#SKIP=};; esac; done   # general kernel: launches alternate fast / exact twin
  env "${envs[@]}" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $OUT/r02_prof_$name \
      python tools/profile_target.py "$@" > $OUT/r02_ncu_$name.log 2>&1
  echo "capture $name rc=$?"
}
cap v8 attn_umma_kernel_v8 X=1 -- --steps 2
cap v7 attn_umma_kernel_v7 SDPA_UMMA_V8=0 -- --steps 2
cap x3_c3 attn_umma_general_kernel SKIP=2 -- --steps 2 --precision bf16x3
cap x3_c2 attn_umma_general_kernel SKIP=2 -- --steps 2 --precision bf16x3 --m 4096 --n 4096
cap cast cvt_in_batch_kernel X=1 -- --steps 2
ls -la $OUT/r02_prof_*.ncu-rep
