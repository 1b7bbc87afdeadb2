#!/bin/bash
# One GPU: the drop-in harness on a c3-sized data file (m=8192, n=65536, d=128; 142 MB of fp64 input) in three host-memory
# modes -- pinned (sdpa_host_alloc, default), malloc + the library's pinned ring, malloc + the driver's own pageable copy.
# "Elapsed time" covers exactly the attention() call (mpi.c:519-522).  M / N can be overridden for a quick run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
M=${M:-8192}; N=${N:-65536}
DATA=/tmp/harness_c3.bin
python - <<PY
import sys, time
sys.path.insert(0, ".")
from oracle import oracle as o
t = time.time()
Q, K, V = o.make_inputs($M, $N, 128, 128, seed=2)
ref = o.attention_f64(Q, K, V)        # OpenMP fp64 restatement of attention.c
o.write_data_file("$DATA", Q, K, V, ref)
print("data file written in %.1f s" % (time.time() - t))
PY
EXE="mpi-parallelized-scaled-dot-product-attention-with-avx-512-optimization_b200/attention_b200"
for MODE in pinned staged driver; do
  for REP in 1 2 3; do
    case $MODE in
      pinned) R=$(SDPA_PRECISION=bf16 $EXE $DATA) ;;
      staged) R=$(SDPA_PRECISION=bf16 HARNESS_PAGEABLE=1 $EXE $DATA) ;;
      driver) R=$(SDPA_PRECISION=bf16 HARNESS_PAGEABLE=1 SDPA_HOST_STAGING=0 $EXE $DATA) ;;
    esac
    echo "$MODE rep$REP: $(echo $R | tr '\n' ' ')" | tee -a $OUT/harness_staging.txt
  done
done
rm -f $DATA
