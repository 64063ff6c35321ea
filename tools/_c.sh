cd $GRAFT_REPO_ROOT
timeout 200 python tools/test_conv_tc.py all 65536 2>&1 | tail -10
