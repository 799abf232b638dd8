#!/bin/bash
# A/B of scheduler knobs on one GPU.  Prints kernel Mrays/s for each workload x setting.
cd $GRAFT_REPO_ROOT
for WL in killeroo4k C2 C3 C4; do
  for CO in 0 1; do for G in "0,0" "3,1"; do
    R=$(RTM_COST_ORDER=$CO RTM_GSS=$G timeout 300 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('%.0f Mrays/s %.3f ms e2e %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']))")
    echo "$WL cost_order=$CO gss=$G : $R"
  done; done
done
