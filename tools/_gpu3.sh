cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; grep -n "^E  \|passed\|failed\|^FAILED\|Error" gpurun_out/t_all.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_fused.json 2> gpurun_out/bench_r2_fused.err; cat gpurun_out/bench_r2_fused.json; tail -3 gpurun_out/bench_r2_fused.err
NO_TORCH_PROF= python tools/prof_train.py 65536 1 > gpurun_out/prof_fused_65536.txt 2>&1; head -60 gpurun_out/prof_fused_65536.txt | cut -c1-200
for c in 2 3 4; do python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_r2_cfg$c.json 2>gpurun_out/bench_r2_cfg$c.err; python -c "
import json;d=json.load(open('gpurun_out/bench_r2_cfg$c.json'));print($c, d['value'], d['ms_per_step'], d['env_only']['us_per_step'], d['e2e']['value'], d['gpu_launches'])"; done
