cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
timeout 600 python tools/bench_nav.py 2>&1 | tail -4
python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_r2_cfg3_async.json 2>gpurun_out/bench_r2_cfg3_async.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2_cfg3_async.json'));print('cfg3', d['value'], d['ms_per_step'], d['env_only']['us_per_step'], d['e2e']['value'], d['gpu_launches'], d['config']['launch_mode'])"; tail -2 gpurun_out/bench_r2_cfg3_async.err
