cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_fused.py tests/test_gpu_learner.py -q -m gpu > gpurun_out/t_fused.log 2>&1; grep -n "^E  \|passed\|failed\|^FAILED\|Error" gpurun_out/t_fused.log | head -60
