cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -x -k "planning_ahead or pipelined" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
timeout 600 python tools/bench_nav.py 2>&1 | tail -6
