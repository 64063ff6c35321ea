cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_bwd -s 8 -c 1 -o gpurun_out/conv_bwd_tc -f python tools/test_conv_tc.py all > gpurun_out/ncu_conv_bwd.log 2>&1; tail -2 gpurun_out/ncu_conv_bwd.log
python -m pytest tests -q -m gpu -x 2>&1 | tail -4
GRAPH=1 python tools/time_train.py 65536 10 2>&1 | tail -1
GRAPH=1 python tools/time_train.py 4096 20 2>&1 | tail -1
