"""Chrome trace of one end-to-end iteration (host-buffer env.step): python tools/e2e_trace.py [E] [u8|f32] -> gpurun_out/e2e_trace.json"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from active_tracking_rl_b200.train import Trainer, default_args

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dt = torch.uint8 if (len(sys.argv) > 2 and sys.argv[2] == "u8") else torch.float32
tr = Trainer(default_args(num_envs=E), "cuda:0")
host = tr.env.alloc_host_buffers(obs_dtype=dt)
for _ in range(2):
    tr.iteration(host=host)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.iteration(host=host)
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/e2e_trace.json")
