"""float32 Linear layers on the tcgen05 tensor cores: thin host side of `track2d_gemm_tf32x3` (csrc/track2d_gemm.cu).

The reference's policy (model.py:116-127,175-182, perception.py:73) is nn.Linear / nn.LSTMCell in float32.  cuBLAS runs
fp32 GEMMs on B200 as SIMT kernels; the library kernel splits every fp32 operand into two TF32 values on the fly and
accumulates three tensor-core products in fp32 (3xTF32), which keeps fp32 accuracy.

    y  = linear(x, W, b, relu)        x (M, K), W (N, K)  ->  (M, N)        A = x K-major,  B = W K-major
    dx = dy W                                                                 A = dy K-major, B = W MN-major
    dW = dy^T x                                                               A = dy MN-major, B = x MN-major (K = batch)

No operand is ever transposed or copied.  There is no fallback: CUDA tensors whose shapes the kernel does not take
(extents not multiples of 4) must go through torch's own linear.
"""
import ctypes as C

import torch

from . import _lib

_WS = {}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _workspace(lib, M, N, K, device):
    n = int(lib.track2d_gemm_workspace_floats(M, N, K))
    if n == 0:
        return None, 0
    key = (device.index, n)
    ws = _WS.get(key)
    if ws is None:  # persistent per (device, size): addresses stay fixed under CUDA-graph replay
        ws = _WS[key] = torch.empty(n, dtype=torch.float32, device=device)
    return ws, n


def gemm(a, a_mn_major, lda, b, b_mn_major, ldb, M, N, K, bias=None, relu=False, out=None):
    """out[m, n] = act(sum_k A(m, k) B(n, k) + bias[n]); see include/track2d.h for the addressing of A and B."""
    lib = _lib.load()
    dev = a.device
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=dev)
    ws, n_ws = _workspace(lib, M, N, K, dev)
    _lib.check(lib.track2d_gemm_tf32x3(_p(a), int(a_mn_major), lda, _p(b), int(b_mn_major), ldb, _p(out), out.stride(0), M, N, K,
                                       _p(bias), int(relu), _p(ws), n_ws, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), lib)
    return out


def supported(x, weight):
    """shapes / layouts the tensor-core path takes"""
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 2
            and weight.shape[0] % 4 == 0 and weight.shape[1] % 4 == 0 and x.stride(1) == 1 and x.stride(0) % 4 == 0
            and weight.is_contiguous() and x.data_ptr() % 16 == 0 and weight.data_ptr() % 16 == 0)


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        M, K = x.shape
        N = weight.shape[0]
        y = gemm(x, False, x.stride(0), weight, False, K, M, N, K, bias=bias, relu=relu)
        ctx.relu = relu
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        M, K = x.shape
        N = weight.shape[0]
        if ctx.relu:
            gy = torch.ops.aten.threshold_backward(gy, y, 0.0)
        gy = gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm(gy, False, N, weight, True, K, M, K, N)                   # dx[m, k] = sum_n dy[m, n] W[n, k]
        if ctx.needs_input_grad[1]:
            gw = gemm(gy, True, N, x, True, x.stride(0), N, K, M)               # dW[n, k] = sum_m dy[m, n] x[m, k]
        if ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb, None


def linear(x, weight, bias=None, relu=False):
    """F.linear (+ optional ReLU) with fp32 accuracy on the tensor cores; falls through to torch for shapes the kernel does not
    take and for CPU tensors (host-side tests)."""
    if supported(x, weight):
        return _Linear.apply(x, weight, bias, relu)
    y = torch.nn.functional.linear(x, weight, bias)
    return torch.relu(y) if relu else y
