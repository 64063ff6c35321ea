cd $GRAFT_REPO_ROOT
timeout 300 python tools/test_conv_tc.py all 65536 2>&1 | tail -9
python tools/prof_train.py 65536 1 2>&1 | grep "conv_tc" | cut -c1-95,150-215
