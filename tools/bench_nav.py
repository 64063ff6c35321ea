"""env-only throughput of the Nav / RPF targets (A* replanning on device): python tools/bench_nav.py [E]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_configs import env_only

E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for env_id in ("Track2D-BlockPartialNav-v0", "Track2D-MazePartialNav-v0", "Track2D-BlockPartialRPF-v0"):
    v, ms, st = env_only(env_id, E, steps=60, warmup=20)
    print("%-30s E=%d  %.3e env-steps/s  %.2f ms/step  status=%d" % (env_id, E, v, ms, st), flush=True)
