#!/bin/bash
# Multi-GPU session: tests + scaling bench at N = 1, 2 (and up to the GPUs present).  bash tools/gpu_multi.sh tag
TAG=${1:-multi}; OUT=gpurun_out/$TAG; mkdir -p $OUT; cd $GRAFT_REPO_ROOT
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
for WL in killeroo4k C4; do
for N in 1 2 4 8; do
  [ $N -gt $NG ] && continue
  [ "$WL" = C4 ] && [ $N -ne 1 ] && [ $N -ne $NG ] && continue   # C4: end points only
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${WL}_n$N.json 2>$OUT/bench_${WL}_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --workload $WL --steps 10 --warmup 3 > $OUT/bench_${WL}_n$N.json 2>$OUT/bench_${WL}_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_${WL}_n$N.json") if l.startswith("{")][-1])
    print("$WL N=$N", "value %.0f Mrays/s  %.3f ms  e2e %.0f (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("$WL N=$N FAILED", e); print(open("$OUT/bench_${WL}_n$N.err").read()[-3000:])
PY
done; done
