"""PCIe copy rates of this box with pinned buffers: one direction at a time, both directions at once (two streams)."""
import sys
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 88_600_000
x, x2 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
h, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def rate(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


def d2h():
    with torch.cuda.stream(s1):
        h.copy_(x, non_blocking=True)


def h2d():
    with torch.cuda.stream(s2):
        x2.copy_(h2, non_blocking=True)


def both():
    d2h(); h2d()


print("D2H alone %.1f GB/s | H2D alone %.1f GB/s | both at once %.1f GB/s each way" % (rate(d2h), rate(h2d), rate(both)))
for chunks in (8, 32):
    c = n // chunks

    def piped():  # D2H chunk i, then its H2D as soon as it has landed (event), the next D2H running meanwhile
        evs = []
        with torch.cuda.stream(s1):
            for i in range(chunks):
                h[i * c:(i + 1) * c].copy_(x[i * c:(i + 1) * c], non_blocking=True)
                e = torch.cuda.Event(); e.record(s1); evs.append(e)
        with torch.cuda.stream(s2):
            for i in range(chunks):
                s2.wait_event(evs[i])
                x2[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
    r = rate(piped)
    print("round trip in %d chunks (H2D of chunk i behind its D2H): %.2f ms per %.1f MB round trip" % (chunks, n / r / 1e6, n / 1e6))
