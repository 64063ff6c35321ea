"""Gradient all-reduce over peer memory (csrc/track2d_peer.cu; reference: utils.py:36-44 ensure_shared_grads, main.py:102-116): the sum
over ranks of the flat gradient in a fixed order.  On one GPU the ranks are simulated by several handles in this process, one CUDA
stream each (the protocol -- publish, handshake, sum, release -- is the same; only the transport of the segment pointers differs); with
two or more GPUs the real thing runs under torch.multiprocessing with CUDA IPC handles exchanged over torch.distributed."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,n", [(2, 1000), (4, 803_469), (8, 13)])
def test_peer_allreduce_sums_in_rank_order(world, n):
    from active_tracking_rl_b200.peer import PeerAllReduce
    dev = torch.device("cuda:0")
    ranks = [PeerAllReduce(n, dev, r, world, connect=False) for r in range(world)]
    segs = [p.segment() for p in ranks]
    for p in ranks:
        p.connect_local(segs)
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    g = torch.Generator(device=dev).manual_seed(3)
    for it in range(6):  # several epochs: a segment is only overwritten after every peer has read it
        grads = [torch.randn(n, generator=g, device=dev) * (10.0 ** (r - 2)) for r in range(world)]
        want = torch.zeros(n, device=dev)
        for x in grads:  # float32, rank order: what every rank must hold afterwards, bit for bit
            want = want + x
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                ranks[r](grads[r])
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(grads[r], want), (it, r)
    assert all(p.status() == 0 for p in ranks)
    for p in ranks:
        p.close()


def test_peer_allreduce_is_graph_capturable():
    from active_tracking_rl_b200.peer import PeerAllReduce
    dev = torch.device("cuda:0")
    n, world = 4096, 2
    ranks = [PeerAllReduce(n, dev, r, world, connect=False) for r in range(world)]
    segs = [p.segment() for p in ranks]
    for p in ranks:
        p.connect_local(segs)
    src = [torch.arange(n, device=dev, dtype=torch.float32) * (r + 1) for r in range(world)]
    buf = [torch.zeros(n, device=dev) for _ in range(world)]
    graphs, streams = [], [torch.cuda.Stream(device=dev) for _ in range(world)]
    for r in range(world):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=streams[r]):
            buf[r].copy_(src[r])
            ranks[r](buf[r])
        graphs.append(gr)
    for it in range(3):
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                graphs[r].replay()
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(buf[r], src[0] + src[1]), (it, r)
    assert all(p.status() == 0 for p in ranks)
    for p in ranks:
        p.close()


@pytest.mark.parametrize("world", [2, 3])
def test_fused_exchange_and_update_equals_allreduce_then_step(world):
    """track2d_peer_sharedadam_step == track2d_peer_allreduce + track2d_sharedadam_step, bit for bit, over several updates (the
    device-resident counter, AMSGrad state and the summed gradient left in grad included)"""
    from active_tracking_rl_b200.peer import PeerAllReduce
    from active_tracking_rl_b200.shared_optim import SharedAdam
    dev = torch.device("cuda:0")
    n = 70_001
    g = torch.Generator(device=dev).manual_seed(5)
    init = torch.randn(n, generator=g, device=dev)

    def make_side():
        params = [torch.nn.Parameter(init.clone())]
        opts = [SharedAdam([torch.nn.Parameter(init.clone())], lr=1e-3, amsgrad=True) for _ in range(world)]
        peers = [PeerAllReduce(opts[0].fp.numel, dev, r, world, connect=False) for r in range(world)]
        segs = [p.segment() for p in peers]
        for p in peers:
            p.connect_local(segs)
        del params
        return opts, peers

    (oa, pa), (ob, pb) = make_side(), make_side()
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    for it in range(4):
        grads = [torch.randn(oa[0].fp.numel, generator=g, device=dev) for _ in range(world)]
        for r in range(world):
            oa[r].fp.grad.copy_(grads[r])
            ob[r].fp.grad.copy_(grads[r])
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                pa[r](oa[r].fp.grad)                                        # exchange, then the plain update
                oa[r].step(max_grad_norm=0.0, grad_scale=1.0 / world)
                ob[r].step(max_grad_norm=0.0, grad_scale=1.0 / world, peer=pb[r])  # both in one kernel
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(oa[r].fp.flat, ob[r].fp.flat) and torch.equal(oa[r].fp.grad, ob[r].fp.grad), (it, r)
            assert torch.equal(oa[r].max_exp_avg_sq, ob[r].max_exp_avg_sq) and torch.equal(oa[r].step_dev, ob[r].step_dev)
            assert torch.equal(ob[r].fp.flat, ob[0].fp.flat)  # replicas identical
    assert all(p.status() == 0 for p in pa + pb)
    for p in pa + pb:
        p.close()


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from active_tracking_rl_b200.train import Trainer, default_args
    tr = Trainer(default_args(num_envs=512, num_steps=5, seed=2), "cuda:%d" % rank, rank, world)
    assert tr.peer is not None
    tr.capture(warmup=1)
    assert tr._graph_apply is None  # ONE graph per iteration, all-reduce inside
    for _ in range(4):
        tr.replay()
    torch.cuda.synchronize()
    flat = tr.optimizer.fp.flat.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], x) for x in gathered)
    out[rank] = (bool(same), tr.peer.status(), tr.env.status())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_training_with_the_all_reduce_inside_one_graph():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29641, out), nprocs=world, join=True)
    assert all(out[r] == (True, 0, 0) for r in range(world)), dict(out)
