"""Time rollout+update iterations: python tools/time_train.py E iters [conv_impl] [cudnn_benchmark]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("EMU") == "1":
    from active_tracking_rl_b200 import blas
    print("fp32 emulation:", blas.enable_fp32_emulation(os.environ.get("EMU_STRATEGY", "performant")), blas.status())
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from active_tracking_rl_b200 import model as M
from active_tracking_rl_b200.train import Trainer, default_args
E = int(sys.argv[1]); iters = int(sys.argv[2])
if len(sys.argv) > 3: M.CONV_IMPL = sys.argv[3]
if len(sys.argv) > 4: torch.backends.cudnn.benchmark = bool(int(sys.argv[4]))
tr = Trainer(default_args(num_envs=E, tf32=os.environ.get("TF32") == "1"), "cuda:0")
use_graph = os.environ.get("GRAPH") == "1"
if use_graph:
    tr.capture()
    step = tr.replay
else:
    step = tr.iteration
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    out = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("tf32=%s graph=%s " % (os.environ.get("TF32") == "1", use_graph) + "E=%d conv=%s bench=%s  %.1f ms/iter  %.3e env-steps/s  mem=%.1f GB  vloss=%.4f" % (E, M.CONV_IMPL, torch.backends.cudnn.benchmark, ms, E * 20 / ms * 1e3,
      torch.cuda.max_memory_allocated() / 1e9, float(out[1].mean())))
