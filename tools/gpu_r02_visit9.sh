#!/bin/bash
# Round-2 visit 9 (EIGHT GPUs; box time is charged 8-fold, keep it short): the scaling bench at 8 and 4 GPUs with the oracle
# parity check and c5 / c4 at their BASELINE shapes, the exchange timeline at 8 GPUs, root vs sliced exchange.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v9.log; rm -f $S $OUT/xtrace_*
# ten-second sanity of the fused kernels on one GPU before eight GPUs are kept busy
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_bf16.py -q -m gpu -p no:cacheprovider -x -k "persistent or variants or overflow" > $OUT/v9_sanity.log 2>&1
rc=$?; echo "sanity rc=$rc" >> $S
if [ $rc -ne 0 ]; then cat $S; tail -20 $OUT/v9_sanity.log; exit 1; fi
run() { # name, gpus, env..., -- args
  local name=$1 g=$2; shift 2
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $g "$@" > $OUT/v9_$name.json 2> $OUT/v9_$name.err
  echo "bench $name rc=$?" >> $S
}
run g8_root 8 SDPA_EXCHANGE_TRACE=$OUT/xtrace_g8 -- --steps 20 --warmup 5
run g8_sliced 8 SDPA_IPC_MERGE=sliced -- --steps 20 --warmup 5 --extra none
run g4_root 4 X=1 -- --steps 20 --warmup 5
run g8_nooverlap 8 SDPA_OVERLAP_PASSES=0 -- --steps 20 --warmup 5 --extra none
cat $S
python tools/exchange_digest.py $OUT/xtrace_g8
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v9_g*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], "parity", d["parity_check"]["ok"], d["parity_check"]["max_abs_err"], d["clocks"].get("per_rank_sm_mhz"))
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), v["stage_ms_per_step"], "parity", v["parity_check"]["ok"], v["parity_check"]["max_abs_err"], "batches", v["q_batches_per_step"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
