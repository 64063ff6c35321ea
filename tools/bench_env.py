"""Env-only throughput probe: step kernel (+ auto-reset) with resident random actions, CUDA-event timed."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from active_tracking_rl_b200.envs import Track2DVecEnv


def run(E, steps, warmup, env_id, obs_dtype, auto_reset, n_action_sets=16, plan_ahead=False):
    env = Track2DVecEnv(env_id, num_envs=E, seed=1, rng="philox", auto_reset=auto_reset, obs_dtype=obs_dtype, plan_ahead=plan_ahead)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = [torch.randint(0, 4, (E, 2), generator=g, device="cuda", dtype=torch.int32) for _ in range(n_action_sets)]
    for i in range(warmup):
        env.step(acts[i % n_action_sets])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        env.step(acts[i % n_action_sets])
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    env.join()
    eps, _ = env.counters()
    env.close()
    bytes_per = 1755 if obs_dtype == torch.float32 else 741
    sps = E / (ms * 1e-3)
    return dict(E=E, env=env_id, obs=str(obs_dtype).split(".")[-1], auto_reset=auto_reset, plan_ahead=plan_ahead, ms_per_step=round(ms, 5),
                env_steps_per_s=round(sps), GBps=round(sps * bytes_per / 1e9, 1), frac_of_6539=round(sps * bytes_per / 6539.2e9, 4),
                episodes=eps)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[4096, 16384, 65536, 262144])
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--env-id", default="Track2D-BlockPartialPZR-v0")
    ap.add_argument("--plan-ahead", action="store_true", help="only the auto-reset rows, with and without standby worlds (T2D_FLAG_PLAN_AHEAD)")
    a = ap.parse_args()
    for E in a.envs:
        if a.plan_ahead:
            for pa in (False, True):
                print(json.dumps(run(E, a.steps, a.warmup, a.env_id, torch.uint8, True, plan_ahead=pa)), flush=True)
            continue
        for dt in (torch.float32, torch.uint8):
            for ar in (False, True):
                print(json.dumps(run(E, a.steps, a.warmup, a.env_id, dt, ar)), flush=True)
