#!/bin/bash
# One GPU-box visit: parity tests, bench, launch list, full ncu capture of the hot kernels.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia_smi.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" 
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log
echo "== bench"
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"
cat $OUT/bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_splat_tile|k_scan_sums|k_scan_compose|k_scan_walk|k_mf_point_l2|k_map_point_unary|k_embed|k_csr_fill|k_csr_count|k_blur_fused|k_neighbours' \
    -s 160 -c 80 -o /tmp/prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2> $OUT/prof.err; gzip -f $OUT/prof_raw.csv
ls -la $OUT
