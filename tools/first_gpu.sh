set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -40
