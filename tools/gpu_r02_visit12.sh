#!/bin/bash
# Round-2 visit 12 (TWO GPUs): every cross-GPU test on the final exchange code (fused wait/publish, pool buffers with peer access,
# sliced exchange fed by the persistent kernel's pieces), bench root vs sliced.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
S=$OUT/summary_v12.log; rm -f $S $OUT/xtrace_*
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > $OUT/v12_pytest_multi.log 2>&1
echo "pytest multi rc=$?" >> $S
run() {
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 "$@" > $OUT/v12_$name.json 2> $OUT/v12_$name.err
  echo "bench $name rc=$?" >> $S
}
run g2_root SDPA_EXCHANGE_TRACE=$OUT/xtrace_g2 -- --steps 20 --warmup 5 --extra c4
run g2_defer SDPA_DEFER_TWIN=2 -- --steps 20 --warmup 5 --extra c4
cat $S; grep -E "passed|failed" $OUT/v12_pytest_multi.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/v12_pytest_multi.log | head
python tools/exchange_digest.py $OUT/xtrace_g2
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v12_g*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k:round(v,4) for k,v in d["stage_ms_per_step"].items()}, d["impl_detail"]["kernel"], "parity", d["parity_check"]["ok"], d["parity_check"]["max_abs_err"])
        for k,v in d.get("configs",{}).items(): print("    ", k, "value", round(v["value"],1), "ms", round(v["ms_per_step"],4), "e2e", round(v["e2e"]["value"],1), v["stage_ms_per_step"], "parity", v["parity_check"]["ok"], v["parity_check"]["max_abs_err"], "batches", v["q_batches_per_step"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
