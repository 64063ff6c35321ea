cd $GRAFT_REPO_ROOT
timeout 300 python tools/test_conv_tc.py all 2>&1 | tail -12
T2D_CONV_IMPL=simt timeout 300 python tools/test_conv_tc.py all 2>&1 | tail -3
