#!/bin/bash
# Sweep launch-configuration knobs (RTM_OCC_MODE x RTM_THREADS) on a few workloads.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/sweep
for WL in killeroo4k C2 C4; do
 for MODE in 0 1 2; do
  for T in 128 256 512 1024; do
    R=$(RTM_OCC_MODE=$MODE RTM_THREADS=$T timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('%.0f Mrays/s %.3f ms' % (d['value'], d['ms_per_step']))")
    echo "$WL mode=$MODE threads=$T : $R"
  done
 done
done | tee gpurun_out/sweep/sweep.txt
