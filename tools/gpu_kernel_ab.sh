#!/bin/bash
# Parity subset + kernel timing of the main workloads (no profiler).  bash tools/gpu_kernel_ab.sh tag [workloads...]
TAG=${1:-ab}; shift
WLS=${@:-killeroo4k C4 C3 C2 C1}
OUT=gpurun_out/$TAG; mkdir -p $OUT; cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
for WL in $WLS; do
  timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline 2>$OUT/bench_$WL.err > $OUT/bench_$WL.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$WL.json"))
    print("$WL", "value %.0f Mrays/s  %.3f ms  e2e %.0f (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("$WL FAILED", e); print(open("$OUT/bench_$WL.err").read()[-2000:])
PY
done
