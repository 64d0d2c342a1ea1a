#!/bin/bash
# One GPU-box visit of round 2: bench lines of every workload, ncu launch list of the default bench command, one
# `ncu --set full` capture of the hot kernels of a C3 step (batch 32), exported as CSV (the .ncu-rep is too large to travel).
# Usage (from the repo root, under gpurun): bash scripts/gpu_round2.sh [tag] [full-capture launches]
TAG=${1:-r02}
NFULL=${2:-48}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia_smi.txt 2>&1
nproc > $OUT/nproc.txt
echo "== bench c3 (default)"
timeout 900 python bench.py > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_c3_reference_arm.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"
for w in c5 c4 c2 c1; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w rc=$?"
done
timeout 600 python bench.py --impl reference --workload c5 --steps 5 --warmup 2 > $OUT/bench_c5_reference_arm.json 2>> $OUT/bench_ref.err
echo "== ncu launch list (default bench command, first 700 launches after the set-up)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
gzip -f $OUT/launches.csv
echo "== ncu full (hot kernels of the first C3 step: $NFULL launches)"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_map_point_unary|k_splat_rows|k_scan_compose|k_scan_walk|k_mf_point_l2|k_embed|k_csr_fill|k_csr_count' \
    -c $NFULL -o /tmp/prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2> $OUT/prof.err; gzip -f $OUT/prof_raw.csv
ls -la $OUT
