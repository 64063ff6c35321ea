"""Bring-up / timing of track2d_gemm_tf32x3 against float64 and torch.mm (fp32 SIMT).  GPU only.

    python tools/test_gemm.py [--big] [--time]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from active_tracking_rl_b200 import gemm as G

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"


def check(name, M, N, K, a_mn, b_mn, bias=False, relu=False, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    A = torch.randn((K, M) if a_mn else (M, K), generator=g, device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), generator=g, device=dev)
    bv = torch.randn(N, generator=g, device=dev) if bias else None
    Am = A.t() if a_mn else A
    Bm = B.t() if b_mn else B
    ref = Am.double() @ Bm.double().t()
    if bias:
        ref = ref + bv.double()
    if relu:
        ref = ref.clamp_min(0)
    out = G.gemm(A, a_mn, A.stride(0), B, b_mn, B.stride(0), M, N, K, bias=bv, relu=relu)
    torch.cuda.synchronize()
    t32 = Am @ Bm.t()
    if bias:
        t32 = t32 + bv
    if relu:
        t32 = t32.clamp_min(0)
    scale = (Am.double().abs() @ Bm.double().abs().t()).clamp_min(1e-30)
    err = ((out.double() - ref).abs() / scale).max().item()
    err32 = ((t32.double() - ref).abs() / scale).max().item()
    bad = int(((out.double() - ref).abs() > 1e-4 * scale).sum().item())
    print("%-28s M=%-6d N=%-5d K=%-6d a_mn=%d b_mn=%d  max|err|/(|A||B|) = %.3e   (torch fp32: %.3e)  bad=%d  %s" % (
        name, M, N, K, a_mn, b_mn, err, err32, bad, "OK" if err < 2e-6 else "FAIL"), flush=True)
    if err >= 2e-6:
        d = (out.double() - ref).abs()
        idx = torch.nonzero(d > 1e-4 * scale)[:8].tolist()
        print("   first bad (m, n):", idx, " out", [out[i, j].item() for i, j in idx][:4], " ref", [ref[i, j].item() for i, j in idx][:4], flush=True)
    return err < 2e-6


def timeit(name, M, N, K, a_mn, b_mn, iters=20):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    Am = A.t() if a_mn else A
    Bm = B.t() if b_mn else B
    out = torch.empty((M, N), device=dev)
    for _ in range(3):
        G.gemm(A, a_mn, A.stride(0), B, b_mn, B.stride(0), M, N, K, out=out)
        torch.mm(Am, Bm.t())
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        G.gemm(A, a_mn, A.stride(0), B, b_mn, B.stride(0), M, N, K, out=out)
    e1.record()
    for _ in range(iters):
        torch.mm(Am, Bm.t())
    e2.record()
    torch.cuda.synchronize()
    t1, t2 = e0.elapsed_time(e1) / iters * 1e3, e1.elapsed_time(e2) / iters * 1e3
    fl = 2.0 * M * N * K
    print("%-28s M=%-6d N=%-5d K=%-6d  tf32x3 %8.1f us (%6.1f TFLOP/s)   torch.mm fp32 %8.1f us (%6.1f TFLOP/s)   x%.2f" % (
        name, M, N, K, t1, fl / t1 / 1e6, t2, fl / t2 / 1e6, t2 / t1), flush=True)


if __name__ == "__main__":
    ok = True
    ok &= check("fwd one tile", 128, 128, 32, 0, 0)
    ok &= check("fwd one tile K=128", 128, 128, 128, 0, 0)
    ok &= check("fwd 2x2 tiles", 256, 256, 512, 0, 0)
    ok &= check("fwd bias+relu", 256, 256, 512, 0, 0, bias=True, relu=True)
    ok &= check("fwd ragged", 300, 132, 100, 0, 0, bias=True)
    ok &= check("dgrad (B MN-major)", 256, 512, 256, 0, 1)
    ok &= check("dgrad ragged", 260, 132, 36, 0, 1)
    ok &= check("wgrad (A,B MN-major)", 256, 512, 1024, 1, 1)
    ok &= check("wgrad ragged split-K", 256, 1024, 4100, 1, 1)
    ok &= check("A MN-major only", 256, 128, 256, 1, 0)
    if "--big" in sys.argv or "--time" in sys.argv:
        ok &= check("fwd fc tracker", 65536, 256, 512, 0, 0, bias=True, relu=True)
        ok &= check("wgrad fc TAT", 256, 1024, 65536, 1, 1)
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    if "--time" in sys.argv:
        E = 65536
        timeit("fc tracker fwd", E, 256, 512, 0, 0)
        timeit("fc TAT fwd", E, 256, 1024, 0, 0)
        timeit("lstm ih fwd", E, 512, 256, 0, 0)
        timeit("lstm hh fwd", E, 512, 128, 0, 0)
        timeit("fc TAT dgrad", E, 1024, 256, 0, 1)
        timeit("lstm ih dgrad", E, 256, 512, 0, 1)
        timeit("lstm hh dgrad", E, 128, 512, 0, 1)
        timeit("fc tracker wgrad", 256, 512, E, 1, 1)
        timeit("fc TAT wgrad", 256, 1024, E, 1, 1)
        timeit("lstm ih wgrad", 512, 256, E, 1, 1)
        timeit("lstm hh wgrad", 512, 128, E, 1, 1)
    sys.exit(0 if ok else 1)
