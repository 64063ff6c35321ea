"""The step-kernel roofline measurement of bench.py on its own (back-to-back launches, observation ring > L2)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from active_tracking_rl_b200.envs import Track2DVecEnv

ENV_ID = "Track2D-BlockPartialPZR-v0"
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = "cuda:0"
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
ring = 8
env = Track2DVecEnv(ENV_ID, num_envs=E, device=dev, seed=1, rng="philox", auto_reset=False)
env.reset()
g = torch.Generator(device=dev).manual_seed(0)
acts = [torch.randint(0, 4, (E, 2), generator=g, device=dev, dtype=torch.int32) for _ in range(ring)]
obs_ring = [torch.empty_like(env.obs) for _ in range(ring)]
rew, done = torch.empty_like(env.reward), torch.empty_like(env.done)
for rep in range(3):
    for i in range(20):
        env.step_into(acts[i % ring], obs_ring[i % ring], rew, done)
    torch.cuda.synchronize()
    n_l = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_l):
        env.step_into(acts[i % ring], obs_ring[i % ring], rew, done)
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / n_l
    achieved = 1755 * E / (k_ms * 1e-3) / 1e9
    print("E=%d  %.2f us per launch  %.1f GB/s algorithmic  frac %.4f of %.1f" % (E, k_ms * 1e3, achieved, achieved / peak, peak), flush=True)
assert env.status() == 0
