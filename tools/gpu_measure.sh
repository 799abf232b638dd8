#!/bin/bash
# One GPU-box session: tests, smoke, bench on every config, ncu launch list + one full capture.
# Usage (from the repo root on the box): bash tools/gpu_measure.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt
lscpu | head -20 > $OUT/cpu.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
for WL in killeroo4k C2 C3 C4 C1; do
  echo "== bench $WL"
  timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 $( [ $WL != killeroo4k ] && echo --no-cpu-baseline ) 2>$OUT/bench_$WL.err | tee $OUT/bench_$WL.json
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee $OUT/bench_reference.json
if [ -z "$NO_NCU" ]; then
CMD="python bench.py --workload killeroo4k --steps 2 --warmup 3 --no-cpu-baseline"
echo "== ncu launch list"
$CMD > $OUT/ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "== ncu full"
$CMD > $OUT/ncu_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:trace_tiles -s 3 -c 1 -o $OUT/prof_trace_tiles $CMD > $OUT/ncu_full.log 2>&1
ls -la $OUT
fi
