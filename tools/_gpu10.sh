cd $GRAFT_REPO_ROOT
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err; tail -c 1500 gpurun_out/bench_r2_2gpu.json; tail -3 gpurun_out/bench_r2_2gpu.err
T2D_NCCL_IN_GRAPH=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/bench_r2_2gpu_ncclgraph.json 2> gpurun_out/bench_r2_2gpu_ncclgraph.err; echo "rc=$?"; python - <<'PY'
import json
for f in ("gpurun_out/bench_r2_2gpu.json","gpurun_out/bench_r2_2gpu_ncclgraph.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["config"]["launch_mode"], d["replicas_identical"], d["e2e"]["value"])
    except Exception as ex: print(f, "ERR", ex)
PY
tail -5 gpurun_out/bench_r2_2gpu_ncclgraph.err
