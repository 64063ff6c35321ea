"""Profile one rollout+update iteration: python tools/prof_train.py E [iters]  -> kernel table on stdout"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("EMU") == "1":
    from active_tracking_rl_b200 import blas
    blas.enable_fp32_emulation(os.environ.get("EMU_STRATEGY", "performant"))
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from active_tracking_rl_b200.train import Trainer, default_args
E = int(sys.argv[1]); iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tr = Trainer(default_args(num_envs=E), "cuda:0")
for _ in range(2):
    tr.iteration()
torch.cuda.synchronize()
if os.environ.get("NO_TORCH_PROF"):
    for _ in range(iters):
        tr.iteration()
    torch.cuda.synchronize()
else:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(iters):
            tr.iteration()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
