#!/bin/bash
mkdir -p gpurun_out/bsweep
for b in 4 8 16 32; do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-profile --batch $b > gpurun_out/bsweep/b$b.json 2> gpurun_out/bsweep/b$b.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bsweep/b$b.json"))
print("batch $b: value %.0f  ms/step %.3f  e2e %.0f clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
PY
done
