"""Where the end-to-end iteration (host-buffer env.step, bench.py's `e2e`) spends its time: python tools/e2e_breakdown.py [E] [u8|f32]
Synchronises after every phase (the loop already synchronises once per env step), so the parts add up to slightly more than bench's figure."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from active_tracking_rl_b200.train import Trainer, default_args

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dt = torch.uint8 if (len(sys.argv) > 2 and sys.argv[2] == "u8") else torch.float32
tr = Trainer(default_args(num_envs=E), "cuda:0")
host = tr.env.alloc_host_buffers(obs_dtype=dt)
if os.environ.get("T2D_FWD_SLICES"):
    v = os.environ["T2D_FWD_SLICES"]
    host["forward_slices"] = int(v) if v.isdigit() else tuple(int(x) for x in v.split(","))
if os.environ.get("T2D_CHUNKS"):
    host["chunks"] = int(os.environ["T2D_CHUNKS"])
if os.environ.get("T2D_PREFETCH") == "0":
    host["prefetch"] = False
for _ in range(2):
    tr.iteration(host=host)
torch.cuda.synchronize()
p, T = tr.player, tr.args.num_steps
eng = p.engine
orig_fwd = eng.forward
acc = {"forward": 0.0, "env+pcie": 0.0, "optimize": 0.0, "total": 0.0}


def timed_forward(*a, **k):
    t0 = time.perf_counter()
    out = orig_fwd(*a, **k)
    torch.cuda.synchronize()
    acc["forward"] += time.perf_counter() - t0
    return out


eng.forward = timed_forward
n = 5
t_all = time.perf_counter()
for _ in range(n):
    p.update_rnn_hiden()
    for _ in range(T):
        f0 = acc["forward"]
        t0 = time.perf_counter()
        p.action_train(host=host)
        torch.cuda.synchronize()
        acc["env+pcie"] += time.perf_counter() - t0 - (acc["forward"] - f0)
    f0 = acc["forward"]
    t0 = time.perf_counter()
    p.optimize(None, tr.optimizer, tr.model, tr.args.train_mode, None, world_size=1, allreduce=tr.allreduce)
    torch.cuda.synchronize()
    acc["optimize"] += time.perf_counter() - t0 - (acc["forward"] - f0)
acc["total"] = time.perf_counter() - t_all
obs_bytes = E * 2 * 169 * (1 if dt == torch.uint8 else 4)
print("E=%d host obs %s: per iteration  forward (21x) %.1f ms | env step + PCIe (20x) %.1f ms (%.2f ms/step; obs %.1f MB each way -> %.1f GB/s if "
      "it were all copy time) | optimize %.1f ms | total %.1f ms" % (E, str(dt).split(".")[-1], 1e3 * acc["forward"] / n, 1e3 * acc["env+pcie"] / n,
      1e3 * acc["env+pcie"] / n / T, obs_bytes / 1e6, obs_bytes / (acc["env+pcie"] / n / T) / 1e9, 1e3 * acc["optimize"] / n, 1e3 * acc["total"] / n))
# raw copy rates of this box for the same buffer
x = torch.empty(obs_bytes, dtype=torch.uint8, device="cuda")
h = torch.empty(obs_bytes, dtype=torch.uint8).pin_memory()
for name, fn in (("D2H", lambda: h.copy_(x, non_blocking=True)), ("H2D", lambda: x.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    print("%s %.1f MB: %.1f GB/s" % (name, obs_bytes / 1e6, 10 * obs_bytes / (time.perf_counter() - t0) / 1e9))
s2 = torch.cuda.Stream()
t0 = time.perf_counter()
for _ in range(10):
    h.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2):
        x2 = getattr(sys.modules[__name__], "_x2", None)
        if x2 is None:
            x2 = sys.modules[__name__]._x2 = torch.empty_like(x)
            h2 = sys.modules[__name__]._h2 = torch.empty_like(h).pin_memory()
        x2.copy_(sys.modules[__name__]._h2, non_blocking=True)
torch.cuda.synchronize()
print("D2H + H2D concurrently: %.1f GB/s each way" % (10 * obs_bytes / (time.perf_counter() - t0) / 1e9))
