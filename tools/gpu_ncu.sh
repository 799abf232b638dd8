#!/bin/bash
# ncu: launch list + one full capture of trace_tiles on a workload.  bash tools/gpu_ncu.sh tag [workload]
TAG=${1:-ncu}; WL=${2:-killeroo4k}
OUT=gpurun_out/$TAG; mkdir -p $OUT; cd $GRAFT_REPO_ROOT
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$WL.csv $CMD > $OUT/ncu_launches.log 2>&1
$CMD > $OUT/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:trace_tiles -s 3 -c 1 -o $OUT/prof_$WL $CMD > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log; ls -la $OUT
