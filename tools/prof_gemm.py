"""A few launches of track2d_gemm_tf32x3 at the policy's shapes, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -c 3 -o gpurun_out/gemm python tools/prof_gemm.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from active_tracking_rl_b200 import gemm as G

E = 65536
dev = "cuda:0"
x = torch.randn(E, 1024, device=dev)
w = torch.randn(256, 1024, device=dev)
gy = torch.randn(E, 256, device=dev)
G.gemm(x, 0, 1024, w, 0, 1024, E, 256, 1024)          # fc TAT forward
G.gemm(gy, 0, 256, w, 1, 1024, E, 1024, 256)          # fc TAT dgrad
G.gemm(gy, 1, 256, x, 1, 1024, 256, 1024, E)          # fc TAT wgrad
torch.cuda.synchronize()
