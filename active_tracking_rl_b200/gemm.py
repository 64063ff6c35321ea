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


_WS_FLOATS = {}


def _workspace(lib, M, N, K, device):
    n = _WS_FLOATS.get((M, N, K))
    if n is None:  # the launch plan is a pure function of the shape: ask the library once
        n = _WS_FLOATS[(M, N, K)] = int(lib.track2d_gemm_workspace_floats(M, N, K))
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
    with _lib.on_device(dev):
        _lib.check(lib.track2d_gemm_tf32x3(_p(a), int(a_mn_major), lda, _p(b), int(b_mn_major), ldb, _p(out), out.stride(0), M, N, K,
                                           _p(bias), int(relu), _p(ws), n_ws, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), lib)
    return out


def supported(x, weight):
    """shapes / layouts the tensor-core path takes"""
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 2
            and weight.shape[0] % 4 == 0 and weight.shape[1] % 4 == 0 and x.stride(1) == 1 and x.stride(0) % 4 == 0
            and weight.is_contiguous() and x.data_ptr() % 16 == 0 and weight.data_ptr() % 16 == 0)


def colsum(x):
    """x.sum(0) of a tall contiguous (M, N) matrix in a fixed summation order (csrc/track2d_lstm.cu); torch's own for widths the
    kernel does not take"""
    M, N = x.shape
    if not (x.is_cuda and x.dtype == torch.float32 and 4 <= N <= 1024 and (N & (N - 1)) == 0 and x.stride(1) == 1 and x.stride(0) % 4 == 0
            and x.data_ptr() % 16 == 0):
        return x.sum(0)
    lib = _lib.load()
    key = (x.device.index, "colsum", M, N)
    ws = _WS.get(key)
    if ws is None:
        ws = _WS[key] = torch.empty(int(lib.track2d_colsum_workspace_floats(M, N)), dtype=torch.float32, device=x.device)
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    with _lib.on_device(x.device):
        _lib.check(lib.track2d_colsum(_p(x), x.stride(0), M, N, _p(out), _p(ws), ws.numel(), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), lib)
    return out


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
            gb = colsum(gy)
        return gx, gw, gb, None


def linear(x, weight, bias=None, relu=False):
    """F.linear (+ optional ReLU) with fp32 accuracy on the tensor cores; falls through to torch for shapes the kernel does not
    take and for CPU tensors (host-side tests)."""
    if supported(x, weight):
        return _Linear.apply(x, weight, bias, relu)
    y = torch.nn.functional.linear(x, weight, bias)
    return torch.relu(y) if relu else y


class _LSTMCellPointwise(torch.autograd.Function):
    """hy, cy = lstm_pointwise(igates, hgates, cx, b_ih, b_hh) -- csrc/track2d_lstm.cu.  The backward also produces the bias
    gradient (column sums of dgates) in the same pass, in a fixed summation order."""

    @staticmethod
    def forward(ctx, igates, hgates, cx, b_ih, b_hh):
        lib = _lib.load()
        E, H4 = igates.shape
        H = H4 // 4
        dev = igates.device
        hy = torch.empty((E, H), dtype=torch.float32, device=dev)
        cy = torch.empty((E, H), dtype=torch.float32, device=dev)
        act = torch.empty((E, H4), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            _lib.check(lib.track2d_lstm_cell_forward(_p(igates), _p(hgates), _p(b_ih), _p(b_hh), _p(cx), cx.stride(0), _p(hy), _p(cy), _p(act), E, H,
                                                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), lib)
        ctx.save_for_backward(cx, cy, act)
        ctx.set_materialize_grads(False)
        return hy, cy

    @staticmethod
    def backward(ctx, dhy, dcy):
        lib = _lib.load()
        cx, cy, act = ctx.saved_tensors
        E, H4 = act.shape
        H = H4 // 4
        dev = act.device
        if dhy is not None and (dhy.stride(1) != 1 or dhy.stride(0) % 4 or dhy.data_ptr() % 16):
            dhy = dhy.contiguous()
        if dcy is not None and (dcy.stride(1) != 1 or dcy.stride(0) % 4 or dcy.data_ptr() % 16):
            dcy = dcy.contiguous()
        dgates = torch.empty((E, H4), dtype=torch.float32, device=dev)
        dcx = torch.empty((E, H), dtype=torch.float32, device=dev)
        db = torch.empty(H4, dtype=torch.float32, device=dev)
        n_ws = int(lib.track2d_lstm_bias_workspace_floats(E, H))
        key = (dev.index, "lstm", n_ws)
        ws = _WS.get(key)
        if ws is None:
            ws = _WS[key] = torch.empty(n_ws, dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            _lib.check(lib.track2d_lstm_cell_backward(_p(dhy), dhy.stride(0) if dhy is not None else H, _p(dcy), dcy.stride(0) if dcy is not None else H,
                                                      _p(cx), cx.stride(0), _p(cy), _p(act), _p(dgates), _p(dcx), _p(db), _p(ws), n_ws, E, H,
                                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), lib)
        return dgates, dgates, dcx, db, db


def lstm_pointwise_supported(igates, cx):
    return (igates.is_cuda and igates.dtype == torch.float32 and igates.shape[1] == 512 and igates.is_contiguous() and cx.stride(1) == 1
            and cx.stride(0) % 4 == 0 and cx.data_ptr() % 16 == 0)


def lstm_pointwise(igates, hgates, cx, b_ih, b_hh):
    return _LSTMCellPointwise.apply(igates, hgates, cx, b_ih, b_hh)
