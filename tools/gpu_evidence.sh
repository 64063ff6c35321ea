# gpurun payload that regenerates the evidence under profiles/: bench lines of configs 2, 3, 5, the per-launch list under graph replay, one ncu --set full of the conv backward
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 300 gpurun_out/bench_r2_final.err
timeout 400 python bench.py --config 2 --no-cpu-baseline > gpurun_out/bench_r2_final_config2.json 2> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --config 3 --no-cpu-baseline > gpurun_out/bench_r2_final_config3.json 2> gpurun_out/bench_cfg3.err
timeout 400 python bench.py --config 5 --envs 8192 --no-cpu-baseline > gpurun_out/bench_r2_final_config5_8192.json 2> gpurun_out/bench_cfg5s.err
timeout 900 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches_r2_final.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_bwd -s 8 -c 1 -o gpurun_out/conv_bwd_tc_final -f python tools/test_conv_tc.py all > gpurun_out/ncu_conv_bwd.log 2>&1; tail -1 gpurun_out/ncu_conv_bwd.log
gzip -f gpurun_out/launches_r2_final.csv
