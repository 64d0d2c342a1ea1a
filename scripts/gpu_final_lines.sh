#!/bin/bash
# Last GPU visit of a round: the full GPU suite, smoke(), and the bench lines of every workload plus the reference arms.
# Usage (from the repo root, under gpurun): bash scripts/gpu_final_lines.sh [tag]
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/bench_c3_reference_arm.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"
for w in c5 c4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w rc=$?"
done
python - <<PY
import json
for f in ("bench_c3","bench_c3_reference_arm","bench_c5","bench_c4"):
    try:
        d=json.loads(open("$OUT/"+f+".json").read().strip().splitlines()[-1])
        print(f, "value %.1f e2e %.1f ms %.3f" % (d["value"], d["e2e"]["value"], d.get("ms_per_step",0)))
    except Exception as e: print(f, "ERR", e)
PY
