#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the bench workloads under torch.distributed.run, one rank per GPU, no collective on the
# data path.  Usage: bash scripts/gpu_scale.sh <N> <tag> [workloads...]
N=$1; TAG=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt; nproc > $OUT/nproc.txt
PORT=29511
for w in "$@"; do
  steps=10; [ "$w" = c3 ] && steps=20
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 --workload $w --steps $steps --warmup 3 --no-cpu-baseline --no-profile > $OUT/bench_${w}_n$N.json 2> $OUT/bench_${w}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --workload $w --steps $steps --warmup 3 --no-profile > $OUT/bench_${w}_n$N.json 2> $OUT/bench_${w}_n$N.err
  fi
  echo "$w n=$N rc=$?"; PORT=$((PORT+1))
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT/bench_${w}_n$N.json") if l.startswith("{")][-1]
    print("  value %.0f e2e %.0f  ms %.3f / %.3f  n_gpus %d scaling %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["n_gpus"], d["scaling"]))
except Exception as e:
    print("  no line:", e)
PY
done
