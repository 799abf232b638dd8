#!/bin/bash
# N-GPU session on a box with N GPUs: multi-GPU parity tests, then bench lines at this N.  bash tools/gpu_scale.sh tag [workloads...]
TAG=${1:-scale}; shift
WLS=${@:-killeroo4k C5 C4}
OUT=gpurun_out/$TAG; mkdir -p $OUT; cd $GRAFT_REPO_ROOT
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu_multi.txt
for WL in $WLS; do
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --no-side-configs > $OUT/bench_${WL}_n$N.json 2>$OUT/bench_${WL}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --workload $WL --steps 10 --warmup 3 --no-side-configs > $OUT/bench_${WL}_n$N.json 2>$OUT/bench_${WL}_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_${WL}_n$N.json") if l.startswith("{")][-1])
    e=d["e2e"]
    print("$WL N=$N", "kernel %.0f Mrays/s %.3f ms | e2e %.0f Mrays/s %.3f ms | parity %s | by rank %s | call %s" % (
        d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], d["parity"].get("image_md5_ok"), e["kernel_ms_by_rank"],
        {k: round(v, 3) for k, v in e["rank0_call_ms"].items()}))
except Exception as ex:
    print("$WL N=$N FAILED", ex); print(open("$OUT/bench_${WL}_n$N.err").read()[-3000:])
PY
done
