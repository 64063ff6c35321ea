cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_fused.py -q -m gpu -k "fused_learner_matches" 2>&1 | grep -E "^FAILED|passed|failed|AssertionError: \(" | cut -c1-300
T2D_CONV_IMPL=simt python -m pytest tests/test_gpu_fused.py -q -m gpu -k "fused_learner_matches" 2>&1 | grep -E "^FAILED|passed|failed" | cut -c1-300
