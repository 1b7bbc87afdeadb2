#!/bin/bash
# Round-2 visit 11 (FOUR GPUs): the scaling bench at 4 GPUs with the oracle parity check and c4 at its BASELINE shape
# (m=16384 > q_batch: ping-pong batches with an exchange per batch), exchange timeline, root vs sliced exchange.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v11.log; rm -f $S $OUT/xtrace_*
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_bf16.py -q -m gpu -p no:cacheprovider -x -k "persistent or variants or overflow" > $OUT/v11_sanity.log 2>&1
rc=$?; echo "sanity rc=$rc" >> $S
if [ $rc -ne 0 ]; then cat $S; tail -20 $OUT/v11_sanity.log; exit 1; fi
run() { local name=$1 g=$2; shift 2
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $g "$@" > $OUT/v11_$name.json 2> $OUT/v11_$name.err
  echo "bench $name rc=$?" >> $S
}
run g4_root 4 SDPA_EXCHANGE_TRACE=$OUT/xtrace_g4 -- --steps 20 --warmup 5
run g4_sliced 4 SDPA_IPC_MERGE=sliced -- --steps 20 --warmup 5 --extra none
run g4_nooverlap 4 SDPA_OVERLAP_PASSES=0 -- --steps 20 --warmup 5 --extra none
cat $S
python tools/exchange_digest.py $OUT/xtrace_g4
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v11_g*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], "parity", d["parity_check"]["ok"], d["parity_check"]["max_abs_err"], d["clocks"].get("per_rank_sm_mhz"))
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), v["stage_ms_per_step"], "parity", v["parity_check"]["ok"], v["parity_check"]["max_abs_err"], "batches", v["q_batches_per_step"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
