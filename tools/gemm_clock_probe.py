"""Times a long run of one GEMM shape in windows while nvidia-smi samples the SM clock (power-cap behaviour of the tensor pipe)."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_tracking_rl_b200 import gemm as G
dev = "cuda:0"
E = 65536
x = torch.randn(E, 1024, device=dev); w = torch.randn(256, 1024, device=dev); out = torch.empty(E, 256, device=dev)
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()
for _ in range(5):
    G.gemm(x, 0, 1024, w, 0, 1024, E, 256, 1024, out=out)
torch.cuda.synchronize()
time.sleep(0.5)
t_start = time.time()
for win in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(400):
        G.gemm(x, 0, 1024, w, 0, 1024, E, 256, 1024, out=out)
    e1.record(); torch.cuda.synchronize()
    print("window %2d  t=%.2fs  %.1f us per GEMM" % (win, time.time() - t_start, e0.elapsed_time(e1) / 400 * 1e3), flush=True)
proc.terminate()
print("clock samples during the run (MHz, W, power cap):", [r[1] for r in rows if r[0] >= t_start][::3][:20])
