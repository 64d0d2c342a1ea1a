#!/bin/bash
# GPU sweep of the unary kernel's staging geometry (kUCap, kUObs, CTAs/SM): rebuilds liblccrf.so per variant on the box.
mkdir -p gpurun_out/sweep
cp lc-crf-slam_b200/csrc/unary.cu /tmp/unary.cu.orig
for v in "1024 8 2" "512 4 3" "512 8 2" "2048 8 1" "1024 4 2"; do
  set -- $v
  sed -e "s/^constexpr int kUCap = [0-9]*;/constexpr int kUCap = $1;/" -e "s/^constexpr int kUObs = [0-9]*;/constexpr int kUObs = $2;/" \
      -e "s/__launch_bounds__(kUWarps \* 32, [0-9])/__launch_bounds__(kUWarps * 32, $3)/" \
      -e "s/(per_sm > [0-9] ? [0-9] : per_sm)/(per_sm > $3 ? $3 : per_sm)/" /tmp/unary.cu.orig > lc-crf-slam_b200/csrc/unary.cu
  make -C lc-crf-slam_b200/csrc -j8 > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sweep/b_$1_$2_$3.json 2> gpurun_out/sweep/b_$1_$2_$3.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep/b_$1_$2_$3.json"))
print("kUCap=$1 kUObs=$2 ctas=$3: value %.0f  unary %.4f ms" % (d["value"], d["kernel_avg_launch_ms"]["k_map_point_unary"]))
PY
done
cp /tmp/unary.cu.orig lc-crf-slam_b200/csrc/unary.cu
