#!/bin/bash
# N-GPU visit (gpurun --gpus N): multi-GPU parity + the scaling bench at 1..N.
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out
rm -f $OUT/summary_multi.log
nvidia-smi -L > $OUT/multi_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -p no:cacheprovider > $OUT/pytest_multi.log 2>&1
echo "pytest_multi rc=$?" >> $OUT/summary_multi.log
for G in 1 2 4 8; do
  if [ $G -le $N ]; then
    if [ $G -eq 1 ]; then
      timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/scale_g$G.json 2> $OUT/scale_g$G.err
    else
      NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $G --steps 10 --warmup 3 > $OUT/scale_g$G.json 2> $OUT/scale_g$G.err
    fi
    echo "bench gpus=$G rc=$?" >> $OUT/summary_multi.log
  fi
done
cat $OUT/summary_multi.log; tail -15 $OUT/pytest_multi.log; python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scale_g*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'gpus',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],4),'roofline',round(d['roofline']['frac'],3),d['stage_ms_per_step'])
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-800:])
PY
