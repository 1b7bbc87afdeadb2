#!/bin/bash
# Round-2 visit 15 (TWO or more GPUs): root forms of the device-side exchange (FORMS, default "pushsync instream") with cast-ahead.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
G=${1:-2}
S=$OUT/summary_v15.log; rm -f $S $OUT/xtrace_* $OUT/v15_*
if [ "$G" = "2" ] && [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -x > $OUT/v15_multi.log 2>&1
rc=$?; echo "multi rc=$rc" >> $S
if [ $rc -ne 0 ]; then cat $S; tail -40 $OUT/v15_multi.log; exit 1; fi
fi
run() { local name=$1 g=$2; shift 2
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $g "$@" > $OUT/v15_$name.json 2> $OUT/v15_$name.err
  echo "bench $name rc=$?" >> $S
}
for form in ${FORMS:-pushsync instream}; do
  run g${G}_${form}_ahead1 $G SDPA_ROOT_MERGE=$form SDPA_CAST_AHEAD=1 SDPA_EXCHANGE_TRACE=$OUT/xtrace_g${G}_${form}_ahead1 -- --steps 30 --warmup 5 --extra none
  run g${G}_${form}_ahead1_b $G SDPA_ROOT_MERGE=$form SDPA_CAST_AHEAD=1 -- --steps 30 --warmup 5 --extra ${EXTRA:-none}
done
cat $S
for form in ${FORMS:-pushsync instream}; do echo "== $form"; python tools/exchange_digest.py $OUT/xtrace_g${G}_${form}_ahead1; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v15_g*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], "parity", d["parity_check"]["ok"], d["parity_check"]["max_abs_err"], d["clocks"].get("per_rank_sm_mhz"))
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), v["stage_ms_per_step"], "parity", v["parity_check"]["ok"], v["parity_check"]["max_abs_err"], "batches", v["q_batches_per_step"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
