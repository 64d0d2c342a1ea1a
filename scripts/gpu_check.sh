#!/bin/bash
# Quick GPU visit: parity tests + smoke + bench (+ optional launch list).  Usage: bash scripts/gpu_check.sh <tag> [launches]
TAG=${1:-chk}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
cat $OUT/bench.json
if [ -n "$2" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
fi
