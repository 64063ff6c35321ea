"""Short env-only loop for ncu: python tools/prof_env.py E steps [u8]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from active_tracking_rl_b200.envs import Track2DVecEnv
E, steps = int(sys.argv[1]), int(sys.argv[2])
dt = torch.uint8 if len(sys.argv) > 3 and sys.argv[3] == "u8" else torch.float32
env = Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=E, seed=1, rng="philox", auto_reset=True, obs_dtype=dt)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
acts = [torch.randint(0, 4, (E, 2), generator=g, device="cuda", dtype=torch.int32) for _ in range(8)]
for i in range(steps):
    env.step(acts[i % 8])
torch.cuda.synchronize()
print("done", env.counters())
