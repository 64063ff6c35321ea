"""conv stack: tensor-core kernels (track2d_conv_tc.cu) vs float64 F.conv2d, and timing against the CUDA-core kernels.
usage: python tools/test_conv_tc.py [fwd|bwd|all] [N]   (T2D_CONV_IMPL=simt selects the CUDA-core kernels)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from active_tracking_rl_b200 import _lib

lib = _lib.load()
what = sys.argv[1] if len(sys.argv) > 1 else "all"
STRIDE = int(os.environ.get("X_STRIDE", "169"))
DEV = "cuda:0"
p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731


def make(N, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    obs = torch.tensor([0, 1, 2, 4], device=DEV, dtype=torch.uint8)[torch.randint(0, 4, (N, 169), generator=g, device=DEV)]
    w1, b1 = torch.rand(16, 1, 3, 3, generator=g, device=DEV) - 0.5, torch.rand(16, generator=g, device=DEV) - 0.5
    w2, b2 = (torch.rand(32, 16, 3, 3, generator=g, device=DEV) - 0.5) * 0.3, torch.rand(32, generator=g, device=DEV) - 0.5
    gy = torch.randn(N, 512, generator=g, device=DEV)
    return obs, w1, b1, w2, b2, gy


def fwd(obs, w1, b1, w2, b2):
    N = obs.shape[0]
    y = torch.empty(N, 512, device=DEV)
    if STRIDE != 169:  # the learner's layout: both agents' images interleaved, one agent's read with stride 338
        wide = getattr(fwd, "_wide", None)
        if wide is None or wide.shape[0] != N:
            wide = fwd._wide = torch.zeros(N, STRIDE, device=DEV, dtype=torch.uint8)
            wide[:, :169] = obs
        obs = wide
    _lib.check(lib.track2d_maze_conv_forward_ex(p(obs), 1, STRIDE, N, p(w1), p(b1), p(w2), p(b2), p(y), st()), lib)
    return y


def bwd(obs, y, gy, w1, b1, w2):
    N = obs.shape[0]
    gr = [torch.zeros_like(t) for t in (w1, b1, w2)] + [torch.zeros(32, device=DEV)]
    _lib.check(lib.track2d_maze_conv_backward_ex(p(obs), 1, 169, p(y), p(gy), N, p(w1), p(b1), p(w2), *[p(t) for t in gr], st()), lib)
    return gr


def ref(obs, w1, b1, w2, b2, gy):
    d = lambda t: t.detach().double().requires_grad_(True)  # noqa: E731
    W1, B1, W2, B2 = d(w1), d(b1), d(w2), d(b2)
    x = obs.double().view(-1, 1, 13, 13)
    y = F.relu(F.conv2d(F.relu(F.conv2d(x, W1, B1, stride=2, padding=1)), W2, B2, stride=2, padding=1)).reshape(-1, 512)
    g = torch.autograd.grad(y, [W1, B1, W2, B2], gy.double())
    return y.detach(), g


ok = True
for N in (1, 7, 8, 9, 1000, 5003):
    obs, w1, b1, w2, b2, gy = make(N, N)
    yr, gr = ref(obs, w1, b1, w2, b2, gy)
    y = fwd(obs, w1, b1, w2, b2)
    torch.cuda.synchronize()
    err = float((y.double() - yr).abs().max())
    line = "N=%5d  fwd max|err| %.2e (max|y| %.2f)" % (N, err, float(yr.abs().max()))
    ok &= err < 2e-5
    if what in ("bwd", "all"):
        g = bwd(obs, y, gy, w1, b1, w2)
        torch.cuda.synchronize()
        for a, b, name in zip(g, gr, ("dw1", "db1", "dw2", "db2")):
            e = float((a.double() - b).abs().max()) / (float(b.abs().max()) + 1e-9)
            line += "  %s %.1e" % (name, e)
            ok &= e < 2e-4
    print(line, flush=True)
print("PARITY", "OK" if ok else "FAILED", "(impl: %s)" % os.environ.get("T2D_CONV_IMPL", "tc"))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 196608
obs, w1, b1, w2, b2, gy = make(N, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("fwd", lambda: fwd(obs, w1, b1, w2, b2)),) + ((("bwd", lambda: bwd(obs, y, gy, w1, b1, w2)),) if what in ("bwd", "all") else ()):
    y = fwd(obs, w1, b1, w2, b2)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%s  N=%d  %.3f ms  (%.1f ns/image)" % (name, N, ms, ms * 1e6 / N), flush=True)
