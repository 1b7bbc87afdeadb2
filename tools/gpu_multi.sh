#!/bin/bash
# N-GPU visit: multi-GPU parity, then the bench at N (and optionally smaller N) with the sliced and the root merge.
cd "$(dirname "$0")/.."
N=${1:-2}
LIST=${2:-$N}
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary_multi.log
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -p no:cacheprovider > $OUT/pytest_multi.log 2>&1
echo "pytest_multi rc=$?" >> $OUT/summary_multi.log
for G in $LIST; do
  for MODE in sliced root; do
    if [ $G -eq 1 ]; then
      [ $MODE = root ] && continue
      timeout 240 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/scale_g${G}_$MODE.json 2> $OUT/scale_g${G}_$MODE.err
    else
      SDPA_IPC_MERGE=$MODE NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $G --steps 20 --warmup 3 > $OUT/scale_g${G}_$MODE.json 2> $OUT/scale_g${G}_$MODE.err
    fi
    echo "bench gpus=$G mode=$MODE rc=$?" >> $OUT/summary_multi.log
  done
done
cat $OUT/summary_multi.log; tail -8 $OUT/pytest_multi.log; python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scale_g*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'gpus',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],4),'roofline',round(d['roofline']['frac'],3),d['stage_ms_per_step'],d['clocks'].get('per_rank_sm_mhz'),d['clocks']['samples'])
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-1500:])
PY
