#!/bin/bash
# Round-end GPU visit without ncu: parity tests, smoke, default bench, reference arm, C4 and C1 bench lines.
# Usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
SECONDS=0; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? wall=${SECONDS}s"; tail -2 $OUT/bench.err
cut -c1-1200 $OUT/bench.json
timeout 600 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 $OUT/bench_ref.json
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "bench c4 rc=$?"; tail -2 $OUT/bench_c4.err; cut -c1-700 $OUT/bench_c4.json
timeout 300 python bench.py --workload c1 --steps 50 --warmup 3 > $OUT/bench_c1.json 2> $OUT/bench_c1.err; echo "bench c1 rc=$?"; tail -2 $OUT/bench_c1.err; cut -c1-500 $OUT/bench_c1.json
