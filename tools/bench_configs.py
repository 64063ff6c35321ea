"""Throughput of the five BASELINE.json configs (single GPU each; the multi-GPU ones at their per-GPU share).
python tools/bench_configs.py > profiles/configs_r1.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from active_tracking_rl_b200.envs import Track2DVecEnv
from active_tracking_rl_b200.train import Trainer, default_args


def env_only(env_id, E, steps=200, warmup=60):
    env = Track2DVecEnv(env_id, num_envs=E, seed=1, rng="philox", auto_reset=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = [torch.randint(0, 4, (E, 2), generator=g, device="cuda", dtype=torch.int32) for _ in range(8)]
    for i in range(warmup):
        env.step(acts[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        env.step(acts[i % 8])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = env.status()
    env.close()
    return E / ms * 1e3, ms, st


def train(env_id, E, iters=5, **kw):
    """rollout + update iterations replayed from one CUDA graph (what bench.py does on a single GPU)"""
    tr = Trainer(default_args(env=env_id, num_envs=E, **kw), "cuda:0")
    tr.capture(warmup=3)
    for _ in range(2):
        tr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        tr.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tr.env.close()
    return E * 20 / ms * 1e3, ms


if __name__ == "__main__":
    rows = [
        ("configs[1] BlockPartialPZR-v0, 4096 envs, tat-maze-lstm", "Track2D-BlockPartialPZR-v0", 4096, dict()),
        ("configs[2] BlockPartialNav-v0, 16384 envs, A*-Nav target, tracker-only training", "Track2D-BlockPartialNav-v0", 16384, dict(train_mode=0)),
        ("configs[3] MazePartialAdv-v0, 32768 envs = 8192/GPU x 4, maze-lstm naive dueling", "Track2D-MazePartialAdv-v0", 8192,
         dict(network="maze-lstm", aux="none", entropy_target=0.01)),
        ("configs[3'] same, all 32768 envs on one GPU", "Track2D-MazePartialAdv-v0", 32768, dict(network="maze-lstm", aux="none", entropy_target=0.01)),
        ("configs[4] BlockPartialPZR-v0, 65536 envs = 8192/GPU x 8, full AD-VAT", "Track2D-BlockPartialPZR-v0", 8192, dict()),
        ("configs[4'] same, all 65536 envs on one GPU (bench.py workload)", "Track2D-BlockPartialPZR-v0", 65536, dict()),
    ]
    print("%-84s %14s %10s %16s %10s" % ("config (1 x B200, per-GPU share)", "env-only st/s", "us/step", "train env-st/s", "ms/iter"))
    for name, env_id, E, kw in rows:
        nav = "Nav" in env_id
        eo, ms, st = env_only(env_id, E, steps=60 if nav else 200, warmup=20 if nav else 60)
        tv, tms = train(env_id, E, iters=3 if nav else 5, **kw)
        print("%-84s %14.3e %10.1f %16.3e %10.1f   status=%d" % (name, eo, ms * 1e3, tv, tms, st), flush=True)
