cd $GRAFT_REPO_ROOT
(time python -m pytest tests -q -m gpu -x) 2>&1 | grep -E "^E  |passed|failed|Error|real" | head -20
timeout 900 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_graph.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300; wc -l gpurun_out/launches_r2_graph.csv
