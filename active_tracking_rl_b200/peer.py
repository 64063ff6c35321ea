"""Gradient all-reduce over NVLink / NVSwitch peer memory (csrc/track2d_peer.cu): the sum over ranks of the flat fp32 gradient, in a
fixed order, as plain kernels on the current stream.  Replaces the NCCL all-reduce of the learner's one exchange between GPUs
(main.py:102-116 + utils.py:36-44: the workers' gradients meet in one shared model) so that a multi-GPU iteration is ONE CUDA graph.

    ar = PeerAllReduce(n_floats, device, rank, world)      # exchanges CUDA IPC handles over torch.distributed
    ar(flat_grad)                                          # every rank, once per update; in place
"""
import ctypes as C

import torch

from . import _lib


class PeerAllReduce(object):
    def __init__(self, n_floats, device, rank, world, connect=True):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.rank, self.world, self.n = int(rank), int(world), int(n_floats)
        self.h = C.c_void_p()
        _lib.check(self.lib.track2d_peer_create(self.rank, self.world, self.n, self.device.index or 0, C.byref(self.h)), self.lib)
        if connect and self.world > 1:
            self._connect_over_torch_distributed()

    def _connect_over_torch_distributed(self):
        import torch.distributed as dist
        mine = (C.c_uint8 * 64)()
        _lib.check(self.lib.track2d_peer_handle(self.h, mine), self.lib)
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device if dist.get_backend() == "nccl" else "cpu")
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        blob = bytes(torch.cat(out).cpu().tolist())
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _lib.check(self.lib.track2d_peer_connect(self.h, buf), self.lib)

    def segment(self):
        p = C.c_void_p()
        _lib.check(self.lib.track2d_peer_segment(self.h, C.byref(p)), self.lib)
        return p.value

    def connect_local(self, segments):
        """peers living in this process (tests): their raw segment pointers, rank-major"""
        arr = (C.c_void_p * self.world)(*segments)
        _lib.check(self.lib.track2d_peer_connect_local(self.h, arr), self.lib)

    def __call__(self, grad):
        if not (grad.is_cuda and grad.dtype == torch.float32 and grad.is_contiguous() and grad.numel() == self.n):
            raise ValueError("PeerAllReduce: a contiguous float32 CUDA tensor of %d elements is expected" % self.n)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.track2d_peer_allreduce(self.h, C.c_void_p(grad.data_ptr()), st), self.lib)
        return grad

    def status(self):
        """0, or 1 + the rank whose signal did not arrive in time (synchronises)"""
        v = C.c_uint64()
        _lib.check(self.lib.track2d_peer_status(self.h, C.byref(v)), self.lib)
        return int(v.value)

    def close(self):
        if self.h:
            self.lib.track2d_peer_destroy(self.h)
            self.h = C.c_void_p()
