#!/bin/bash
# GPU sweep of the long-row scan chunk (entries per (chunk, label) composite task): rebuilds liblccrf.so per variant on the
# box, checks the long-row parity tests and times the bench in ordered mode.  kScanChunk is the one constant everything
# else derives from (engine.cuh: kChunkRecBytes; filter.cu: kScanIT, crossing window).
# Usage (under gpurun, from the repo root): bash scripts/scan_sweep.sh
mkdir -p gpurun_out/scansweep
cp lc-crf-slam_b200/csrc/engine.cuh /tmp/engine.cuh.orig
for v in 2048 4096 8192; do
  sed -e "s/^constexpr int kScanChunk = [0-9]*;/constexpr int kScanChunk = $v;/" /tmp/engine.cuh.orig > lc-crf-slam_b200/csrc/engine.cuh
  make -C lc-crf-slam_b200/csrc -j8 > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "long_rows or full_size or slam_crf_parity or c3_small" > gpurun_out/scansweep/t_$v.log 2>&1; echo "kScanChunk=$v parity rc=$?"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --splat ordered > gpurun_out/scansweep/b_$v.json 2> gpurun_out/scansweep/b_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/scansweep/b_$v.json"))
k=d["kernel_avg_launch_ms"]
print("kScanChunk=$v: value %.0f e2e %.0f | sums %.4f compose %.4f walk %.4f rows %.4f unary %.4f" % (d["value"], d["e2e"]["value"], k.get("k_scan_sums",0), k.get("k_scan_compose",0), k.get("k_scan_walk",0), k.get("k_splat_rows",0), k.get("k_map_point_unary",0)))
PY
done
cp /tmp/engine.cuh.orig lc-crf-slam_b200/csrc/engine.cuh
make -C lc-crf-slam_b200/csrc -j8 > /dev/null 2>&1
