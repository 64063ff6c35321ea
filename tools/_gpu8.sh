cd $GRAFT_REPO_ROOT
timeout 300 python tools/test_conv_tc.py all 2>&1 | tail -10
python -m pytest tests/test_gpu_learner.py -q -m gpu -k "conv_stack" 2>&1 | grep -E "^E  |passed|failed" | head -10
GRAPH=1 python tools/time_train.py 65536 10 2>&1 | tail -1
