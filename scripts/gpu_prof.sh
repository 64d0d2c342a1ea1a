#!/bin/bash
# Profile-only GPU visit: launch list + full ncu capture of selected kernels (kept small: gpurun_out <= 64 MiB).
# Usage: bash scripts/gpu_prof.sh <tag> [kernel-regex] [skip] [count]
TAG=${1:-prof}
REGEX=${2:-'k_splat_staged|k_splat_scan|k_products|k_mf_point_l2|k_map_point_unary|k_embed|k_csr_fill|k_blur_fused'}
SKIP=${3:-160}
COUNT=${4:-80}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" \
    -s $SKIP -c $COUNT -o /tmp/prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
# the report itself is too large to travel: export the pages here
ncu -i /tmp/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2> $OUT/prof_raw.err
ncu -i /tmp/prof.ncu-rep --page details --csv > $OUT/prof_details.csv 2>> $OUT/prof_raw.err
python scripts/ncu_source_summary.py /tmp/prof.ncu-rep $OUT 2>> $OUT/prof_raw.err
gzip -f $OUT/prof_raw.csv $OUT/prof_details.csv
ls -la $OUT; du -sh gpurun_out
