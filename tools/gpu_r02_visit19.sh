#!/bin/bash
# Round-2 visit 19 (FOUR GPUs): the new defaults (in-stream root merge + cast-ahead) at 4 GPUs with c4 at its BASELINE shape, and the
# pushsync root form beside it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
G=${1:-4}
S=$OUT/summary_v19.log; rm -f $S $OUT/xtrace_* $OUT/v19_*
run() { local name=$1 g=$2; shift 2
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $g "$@" > $OUT/v19_$name.json 2> $OUT/v19_$name.err
  echo "bench $name rc=$?" >> $S
}
run g${G}_default $G SDPA_EXCHANGE_TRACE=$OUT/xtrace_g${G}_default -- --steps 30 --warmup 5
run g${G}_pushsync $G SDPA_ROOT_MERGE=pushsync SDPA_EXCHANGE_TRACE=$OUT/xtrace_g${G}_pushsync -- --steps 30 --warmup 5 --extra none
cat $S
for t in default pushsync; do echo "== $t"; python tools/exchange_digest.py $OUT/xtrace_g${G}_$t; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v19_g*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], "parity", d["parity_check"]["ok"], d["parity_check"]["max_abs_err"], d["clocks"].get("per_rank_sm_mhz"))
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), v["stage_ms_per_step"], "parity", v["parity_check"]["ok"], v["parity_check"]["max_abs_err"], "batches", v["q_batches_per_step"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
