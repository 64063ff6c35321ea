cd $GRAFT_REPO_ROOT
timeout 300 python tools/test_conv_tc.py fwd 2>&1 | tail -4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_fwd -s 8 -c 1 -o gpurun_out/conv_fwd_tc -f python tools/test_conv_tc.py fwd > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log
