#!/bin/bash
# Config C5: kernel time per (grid_res, tuning environment) case (tools/c5_probe.py).
#   bash tools/gpu_c5_ab.sh tag "512:RTM_POOL=0" "512:" "640:" ...
TAG=${1:-c5ab}; shift
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/$TAG
python tools/c5_probe.py "$@" 2>&1 | tee gpurun_out/$TAG/c5_ab.txt
