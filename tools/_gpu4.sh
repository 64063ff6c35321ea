cd $GRAFT_REPO_ROOT
timeout 300 python tools/test_conv_tc.py fwd 2>&1 | tail -12
T2D_CONV_IMPL=simt timeout 300 python tools/test_conv_tc.py fwd 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_learner.py -q -m gpu -x 2>&1 | tail -5
tools/learn_experiment.sh curve4096 4096 1000 50 > /dev/null 2>&1; cat gpurun_out/learn_curve4096/evals.txt | cut -c1-120
