set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv
free -g | head -2; nproc
(time python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu) > gpurun_out/t_fullsize.log 2>&1; tail -15 gpurun_out/t_fullsize.log
(time python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_fullsize.py) > gpurun_out/t_gpu.log 2>&1; tail -5 gpurun_out/t_gpu.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:step_kernel --csv --log-file gpurun_out/step_traffic.csv python tools/measure_step_traffic.py run > gpurun_out/traffic.log 2>&1; tail -3 gpurun_out/traffic.log
python bench.py --steps 10 --warmup 3 --e2e-obs f32 --e2e-steps 5 > gpurun_out/bench_head_f32e2e.json 2> gpurun_out/bench_head.err; cat gpurun_out/bench_head_f32e2e.json
