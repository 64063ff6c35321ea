cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -2
python tools/prof_train.py 65536 1 2>&1 | grep "gemm_tf32x3" | cut -c1-95,150-215
GRAPH=1 python tools/time_train.py 65536 10 2>&1 | tail -1
