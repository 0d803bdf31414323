set -x
cd $GRAFT_REPO_ROOT
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_golden.py -m gpu -q -x -k "detection_step" 2>&1 | tail -30; done
