cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_learner.py -x -q 2>&1 | tail -2
GRAPH=1 python tools/time_train.py 65536 10 2>&1 | tail -1
