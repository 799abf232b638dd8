#!/bin/bash
# Quick GPU check: parity tests + bench of the main configs (no profiler).  bash tools/gpu_quick.sh [tag]
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
for WL in killeroo4k C2 C3 C4 C1; do
  timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline 2>$OUT/bench_$WL.err > $OUT/bench_$WL.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$WL.json"))
    print("$WL", "value %.0f Mrays/s  %.3f ms  e2e %.0f  fp32frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline",{}).get("fp32",{}).get("frac",0)))
except Exception as e:
    print("$WL FAILED", e); print(open("$OUT/bench_$WL.err").read()[-2000:])
PY
done
