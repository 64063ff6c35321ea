cd $GRAFT_REPO_ROOT
python tools/prof_train.py 65536 1 > gpurun_out/prof_tc_65536.txt 2>&1; head -34 gpurun_out/prof_tc_65536.txt | cut -c1-100,165-200
