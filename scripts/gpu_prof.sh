#!/bin/bash
# Profile-only GPU visit: full ncu capture of selected kernels with per-instruction (SASS + CUDA line) stall samples,
# exported as CSV on the box (the .ncu-rep itself is too large to travel: gpurun_out <= 64 MiB).
# Usage: bash scripts/gpu_prof.sh <tag> <kernel-regex> <skip> <count> [bench args...]
TAG=${1:-prof}; REGEX=$2; SKIP=${3:-0}; COUNT=${4:-4}; shift 4
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" \
    -s $SKIP -c $COUNT -o /tmp/prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile "$@" > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2> $OUT/prof.err
ncu -i /tmp/prof.ncu-rep --page details --csv > $OUT/prof_details.csv 2>> $OUT/prof.err
ncu -i /tmp/prof.ncu-rep --page source --print-source cuda,sass --csv > $OUT/prof_source.csv 2>> $OUT/prof.err
gzip -f $OUT/prof_raw.csv $OUT/prof_details.csv $OUT/prof_source.csv
ls -la $OUT; du -sh gpurun_out
