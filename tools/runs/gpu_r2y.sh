set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_search.py -m gpu -q -x -k pipeline 2>&1 | tail -30
