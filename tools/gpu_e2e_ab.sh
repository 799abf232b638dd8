#!/bin/bash
# A/B of the end-to-end path on one GPU: env settings x workloads.  bash tools/gpu_e2e_ab.sh "ENV1=.. ENV2=..|ENV=..|" WL...
cd $GRAFT_REPO_ROOT
IFS='|' read -ra SETS <<< "$1"; shift
for WL in "$@"; do for S in "${SETS[@]}" ""; do
  R=$(env $S timeout 300 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --no-side-configs 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); e=d['e2e']; m=e['rank0_call_ms']
print('kernel %.3f ms | e2e %.3f ms (kernel inside %.3f, launched %.3f, traced %.3f, copied %.3f) parity %s' % (d['ms_per_step'], e['ms_per_step'], e['kernel_ms_per_step'], m['launched'], m['traced'], m['copied'], d['parity']['image_md5_ok']))")
  echo "$WL [${S:-default}] $R"
done; done
