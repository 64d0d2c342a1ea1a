#!/bin/bash
mkdir -p gpurun_out/r02w2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scan_compose_w' --launch-skip 2 -c 2 -o /tmp/prof1 -f python scripts/step_probe.py 32 1 > gpurun_out/r02w2/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof1.ncu-rep --page raw --csv > gpurun_out/r02w2/prof_raw.csv 2> gpurun_out/r02w2/prof.err
ncu -i /tmp/prof1.ncu-rep --page source --csv --print-source sass > gpurun_out/r02w2/prof_source.csv 2>> gpurun_out/r02w2/prof.err
gzip -f gpurun_out/r02w2/prof_source.csv gpurun_out/r02w2/prof_raw.csv
ls -la gpurun_out/r02w2
