"""world_size-2 data-parallel logic on CPU (gloo): identical initial replicas, one sum all-reduce of the FLAT
gradient buffer per rollout, 1/world scaling == the mean of the per-shard gradients, and the per-rank env /
sampling seeds (seed + rank, train.py:20).  The fused optimizer kernel itself is CUDA-only and is covered by the
GPU tests; here the host-side plumbing that feeds it is checked."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import a3c_oracle
from active_tracking_rl_b200.model import build_model
from active_tracking_rl_b200.shared_optim import FlatParams
from active_tracking_rl_b200.spaces import Box, Discrete
from active_tracking_rl_b200.train import default_args

OBS = [Box(0, 6, (1, 13, 13)), Box(0, 6, (1, 13, 13))]
ACT = [Discrete(4), Discrete(4)]


def shard_loss(model, obs, forced):
    E = obs.shape[0]
    hx, cx = torch.zeros(E, 2, 128), torch.zeros(E, 2, 128)
    v, a, ent, lp, _, rp = model((obs, (hx, cx)), False, forced)
    per_env = (-(lp * 0.7) - 0.01 * ent + 0.25 * (1.0 - v) ** 2).sum(1) + (rp.squeeze(1) - 0.3).abs()
    return per_env.mean()


def make_data(E, seed):
    rs = np.random.RandomState(seed)
    obs = torch.from_numpy(rs.choice([0, 1, 2, 4], size=(E, 2, 1, 13, 13), p=[0.7, 0.2, 0.05, 0.05]).astype(np.float32))
    forced = torch.from_numpy(rs.randint(0, 4, size=(E, 2)))
    return obs, forced


def worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    args = default_args(seed=1)
    torch.manual_seed(args.seed)  # identical replicas (Trainer does the same before build_model)
    model = build_model(OBS, ACT, args, torch.device("cpu"))
    fp = FlatParams(model.parameters())
    w0 = fp.flat.clone()
    obs, forced = make_data(4, 100 + rank)  # each rank owns its own env shard
    fp.zero_grad()
    shard_loss(model, obs, forced).backward()
    local = fp.grad.clone()
    dist.all_reduce(fp.grad, op=dist.ReduceOp.SUM)  # the ONE collective per rollout
    g = fp.grad * (1.0 / world)                     # grad_scale handed to track2d_sharedadam_step
    gathered = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    ws = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(ws, w0)
    if rank == 0:
        out.put(dict(same_grad=bool(torch.equal(gathered[0], gathered[1])), same_w0=bool(torch.equal(ws[0], ws[1])),
                     g=g.numpy(), local_norm=float(local.norm()), numel=fp.numel))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch_mean():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["same_w0"], "replicas must start from identical weights"
    assert res["same_grad"], "after the all-reduce every rank holds the same gradient"
    # single process over the union of the two shards
    args = default_args(seed=1)
    torch.manual_seed(args.seed)
    model = build_model(OBS, ACT, args, torch.device("cpu"))
    fp = FlatParams(model.parameters())
    assert fp.numel == res["numel"] and fp.numel >= 801291 and fp.numel % 64 == 0
    o0, f0 = make_data(4, 100)
    o1, f1 = make_data(4, 101)
    fp.zero_grad()
    shard_loss(model, torch.cat([o0, o1]), torch.cat([f0, f1])).backward()
    assert np.allclose(fp.grad.numpy(), res["g"], rtol=1e-4, atol=1e-6)
    assert res["local_norm"] > 0


def test_flat_params_keep_state_dict_and_views():
    model = build_model(OBS, ACT, default_args(), torch.device("cpu"))
    sd = a3c_oracle.det_state_dict(tat=True)
    fp = FlatParams(model.parameters())
    model.load_state_dict(sd)  # in-place into the flat buffer
    for (name, p), off in zip(model.named_parameters(), fp.offsets):
        assert off % 64 == 0
        assert torch.equal(fp.flat[off:off + p.numel()].view_as(p), sd[name])
        assert p.grad.data_ptr() == fp.grad.data_ptr() + 4 * off
    obs, forced = make_data(3, 5)
    fp.zero_grad()
    shard_loss(model, obs, forced).backward()
    for p, off in zip(fp.params, fp.offsets):  # autograd accumulated INTO the flat buffer
        assert p.grad.data_ptr() == fp.grad.data_ptr() + 4 * off
    assert float(fp.grad.abs().sum()) > 0
