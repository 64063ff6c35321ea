# gpurun payload: forward-conv microbench, the whole GPU test suite, the default bench line (gpurun -- "bash tools/gpu_suite.sh")
cd $GRAFT_REPO_ROOT
timeout 300 python tools/test_conv_tc.py fwd 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
