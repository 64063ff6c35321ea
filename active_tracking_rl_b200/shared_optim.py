"""SharedAdam (shared_optim.py:90-175) for the synchronous on-device learner.

The reference keeps Adam state in POSIX shared memory so 16 Hogwild workers can update it without
locks (main.py:86,93).  Here one process per GPU owns a replica: parameters, gradients and the three
moment tensors live in FLAT fp32 buffers (every nn.Parameter is a view into one allocation), so
`clip_grad_norm_ -> ensure_shared_grads -> optimizer.step()` (player_util.py:157-160) is two kernel
launches of libtrack2d (track2d_sharedadam_step), and a multi-GPU run needs exactly one NCCL
all-reduce of one contiguous buffer per rollout.
"""
import ctypes as C

import torch

from . import _lib

_ALIGN = 64  # floats; keeps every tensor 256-byte aligned inside the flat buffers


class FlatParams(object):
    """Re-homes a list of parameters into one flat buffer (+ a flat gradient buffer)."""

    def __init__(self, params):
        self.params = [p for p in params]
        assert self.params, "no parameters"
        dev = self.params[0].device
        self.offsets = []
        n = 0
        for p in self.params:
            assert p.dtype == torch.float32 and p.device == dev
            self.offsets.append(n)
            n += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = n
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, self.offsets):
            view = self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[off:off + p.numel()].view_as(p)

    def zero_grad(self):
        self.grad.zero_()
        for p, off in zip(self.params, self.offsets):  # re-attach in case something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                p.grad = self.grad[off:off + p.numel()].view_as(p)


class SharedAdam(object):
    """AMSGrad with eps added after the square root and bias correction folded into the step size
    (shared_optim.py:155-173), fused with clip_grad_norm_ (player_util.py:157).

    One update counter for the whole flat buffer (the reference keeps state['step'] per parameter and skips parameters whose
    grad is None).  The two agree whenever every parameter of the optimizer gets a gradient in every update -- all README
    commands.  With --init-step > 0 (tracker-only phase while the optimizer holds both agents) the target's moments decay and
    its bias-correction step runs ahead during that phase; its weights do not move (zero gradient, zero first moment)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-3, weight_decay=0, amsgrad=True):
        if weight_decay != 0 or not amsgrad:
            raise NotImplementedError("the 2D path uses amsgrad=True, weight_decay=0 (main.py:29,92)")
        self.lib = _lib.load()
        self.fp = params if isinstance(params, FlatParams) else FlatParams(params)
        if not self.fp.flat.is_cuda:
            raise _lib.Track2DError("SharedAdam's fused step is a CUDA kernel; there is no CPU fallback")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        z = lambda: torch.zeros_like(self.fp.flat)  # noqa: E731
        self.exp_avg, self.exp_avg_sq, self.max_exp_avg_sq = z(), z(), z()
        self.step_count = 0  # host mirror of the device-resident update counter
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.fp.flat.device)
        self.norm_scratch = torch.zeros(2, dtype=torch.float32, device=self.fp.flat.device)

    def share_memory(self):  # single process per GPU: nothing to share
        return self

    def zero_grad(self):
        self.fp.zero_grad()

    def step(self, max_grad_norm=50.0, grad_scale=1.0, peer=None):
        """one update.  peer (peer.PeerAllReduce): the gradient is first summed over the ranks -- inside the SAME kernel when there is no
        clipping (track2d_peer_sharedadam_step), else by peer(grad) followed by the plain step; fp.grad holds the sum afterwards"""
        self.step_count += 1
        dev = self.fp.flat.device
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with _lib.on_device(dev):
            if peer is not None and not (max_grad_norm and max_grad_norm > 0):
                _lib.check(self.lib.track2d_peer_sharedadam_step(
                    peer.h, p(self.fp.flat), p(self.fp.grad), p(self.exp_avg), p(self.exp_avg_sq), p(self.max_exp_avg_sq), self.lr,
                    self.betas[0], self.betas[1], self.eps, float(grad_scale), p(self.norm_scratch), p(self.step_dev), st), self.lib)
                return
            if peer is not None:
                peer(self.fp.grad)
            _lib.check(self.lib.track2d_sharedadam_step(
                p(self.fp.flat), p(self.fp.grad), p(self.exp_avg), p(self.exp_avg_sq), p(self.max_exp_avg_sq), self.fp.numel,
                self.step_count, self.lr, self.betas[0], self.betas[1], self.eps, float(max_grad_norm or 0.0), float(grad_scale),
                p(self.norm_scratch), p(self.step_dev), st), self.lib)

    def advance_for_replay(self):
        """a CUDA-graph replay of step() advances the device counter by itself; keep the host mirror in sync"""
        self.step_count += 1

    def grad_norm(self):
        """total gradient norm seen by the last step (before clipping; only computed when clipping is on)"""
        return float(self.norm_scratch[0].sqrt().item())

    def state_dict(self):
        return dict(step=self.step_count, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, max_exp_avg_sq=self.max_exp_avg_sq)

    def load_state_dict(self, sd):
        self.step_count = int(sd['step'])
        self.step_dev.fill_(self.step_count)
        self.exp_avg.copy_(sd['exp_avg'])
        self.exp_avg_sq.copy_(sd['exp_avg_sq'])
        self.max_exp_avg_sq.copy_(sd['max_exp_avg_sq'])
