#!/bin/bash
# Times bench.py against prebuilt experimental variants of liblccrf.so (gpurun_in/liblccrf_<tag>.so); timing only.
cp lc-crf-slam_b200/liblccrf.so /tmp/liblccrf_base.so
for so in /tmp/liblccrf_base.so gpurun_in/liblccrf_*.so; do
  cp $so lc-crf-slam_b200/liblccrf.so
  echo "== $so"; timeout 300 python scripts/step_probe.py 2>&1 | tail -2
done
cp /tmp/liblccrf_base.so lc-crf-slam_b200/liblccrf.so
