cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_tc.json 2> gpurun_out/bench_r2_tc.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_tc.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "e2e_u8", d["e2e_u8"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "launches", d["gpu_launches"], "roof", d["roofline"]["frac"])
PY
tail -3 gpurun_out/bench_r2_tc.err
