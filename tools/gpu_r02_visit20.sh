#!/bin/bash
# Round-2 visit 20 (TWO GPUs): the default root form "auto" (pushsync for single-batch passes, comm-stream merge for multi-batch):
# the two multi-process tests that mix the forms, then the bench with c4 (two Q batches per pass) as an extra.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v20.log; rm -f $S $OUT/v20_*
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -x -k "one_process_per_gpu or overlapped" > $OUT/v20_multi.log 2>&1
rc=$?; echo "multi rc=$rc" >> $S
if [ $rc -ne 0 ]; then cat $S; tail -40 $OUT/v20_multi.log; exit 1; fi
NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 30 --warmup 5 --extra c4 > $OUT/v20_g2_auto.json 2> $OUT/v20_g2_auto.err
echo "bench rc=$?" >> $S
cat $S; tail -2 $OUT/v20_multi.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v20_g*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], "parity", d["parity_check"]["ok"], d["parity_check"]["max_abs_err"], d["clocks"].get("per_rank_sm_mhz"))
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), v["stage_ms_per_step"], "parity", v["parity_check"]["ok"], v["parity_check"]["max_abs_err"], "batches", v["q_batches_per_step"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
