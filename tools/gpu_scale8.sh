#!/bin/bash
# 8-GPU visit: the bench at N=8 with the sliced (default) and the root merge of the device-side exchange.
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
OUT=gpurun_out
for MODE in sliced root; do
  SDPA_IPC_MERGE=$MODE NCCL_DEBUG=WARN timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 > $OUT/scale_g${N}_$MODE.json 2> $OUT/scale_g${N}_$MODE.err
  echo "bench gpus=$N mode=$MODE rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/scale_g${N}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'gpus',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms',round(d['ms_per_step'],4),'roofline',round(d['roofline']['frac'],3),d['stage_ms_per_step'],d['clocks'].get('per_rank_sm_mhz'),d['clocks']['samples'],d['clocks']['reasons'])
    except Exception as e: print(f,'ERR',e, open(f.replace('.json','.err')).read()[-1500:])
PY
