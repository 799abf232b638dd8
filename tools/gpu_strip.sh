#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for WL in killeroo4k C3 C2; do
 for PX in 2 4 8 16 32; do
    R=$(RTM_STRIP_PIXELS=$PX timeout 300 python bench.py --workload $WL --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('%.0f Mrays/s %.3f ms' % (d['value'], d['ms_per_step']))")
    echo "$WL strip_pixels=$PX : $R"
 done
done
