"""Checks against the LIVE reference tree (build container only; skipped where /root/reference is absent, e.g. on
the GPU box): fresh seeds beyond the committed fixtures, checkpoint interchange with the reference's own
A3C_Dueling, and per-env forward parity of the batched policy with the reference modules."""
import os

import numpy as np
import pytest
import torch

import oracle
import ref

pytestmark = pytest.mark.skipif(not ref.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("env_id,seed", [("Track2D-BlockPartialPZR-v0", 31), ("Track2D-MazePartialRam-v0", 32), ("Track2D-BlockPartialNav-v0", 33),
                                         ("Track2D-MazePartialNav-v0", 34), ("Track2D-EmptyPartialFar-v0", 35), ("Track2D-BlockFullRPF-v1", 36)])
def test_oracle_equals_live_reference_on_fresh_seeds(env_id, seed):
    gym = ref.load_reference()
    env = gym.make(env_id)
    o = oracle.OracleEnv(env_id)
    np.random.seed(seed)
    o.seed(seed)
    rs = np.random.RandomState(seed + 1)
    for ep in range(3):
        a, b = env.reset(), o.reset()
        assert (np.asarray(a) == b).all()
        for t in range(60):
            act = [int(rs.randint(4)), int(rs.randint(4))]
            if ep == 1:
                act[0] = 0
            oa, ra, da, _ = env.step(act)
            ob, rb, db, _ = o.step(act)
            assert (np.asarray(oa) == ob).all() and np.asarray(ra, np.float64).tobytes() == rb.tobytes() and bool(da) == db
            if da:
                break
    st = np.random.get_state()
    key, pos = o.rng_state()
    assert st[2] == pos and (st[1] == key).all()


@pytest.mark.parametrize("network", ["tat-maze-lstm", "maze-lstm"])
def test_checkpoints_interchange_with_reference_model(network, tmp_path):
    ref.load_reference()
    from environment import create_env
    from model import build_model as ref_build
    from active_tracking_rl_b200.model import build_model
    from active_tracking_rl_b200.spaces import Box, Discrete
    from active_tracking_rl_b200.train import default_args
    args = ref.RefArgs(network=network)
    env = create_env("Track2D-BlockPartialPZR-v0", args)
    torch.manual_seed(0)
    rmodel = ref_build(env.observation_space, env.action_space, args, torch.device("cpu"))
    mine = build_model([Box(0, 6, (1, 13, 13))] * 2, [Discrete(4)] * 2, default_args(network=network), torch.device("cpu"))
    # reference -> here (what test.py:124 saves)
    path = os.path.join(tmp_path, "all-new.dat")
    torch.save(rmodel.state_dict(), path)
    mine.load_state_dict(torch.load(path), strict=True)
    torch.save(rmodel.player0.state_dict(), os.path.join(tmp_path, "tracker-new.dat"))
    mine.player0.load_state_dict(torch.load(os.path.join(tmp_path, "tracker-new.dat")), strict=True)
    # here -> reference (main.py:81-85 --load-model-dir)
    torch.save(mine.state_dict(), path)
    rmodel.load_state_dict(torch.load(path), strict=True)
    # and the two compute the same thing per env
    np.random.seed(3)
    state = torch.from_numpy(np.float32(env.reset()))  # frame_stack -> (2, 1, 1, 13, 13)
    hx, cx = torch.randn(2, 128) * 0.2, torch.randn(2, 128) * 0.2
    orig = torch.Tensor.multinomial
    forced = [2, 1]
    it = iter(forced)
    torch.Tensor.multinomial = lambda self, n, *a, **k: torch.tensor([[next(it)]])
    try:
        rv, ra, rent, rlp, (rh, rc), rpred = rmodel((state, (hx, cx)))
    finally:
        torch.Tensor.multinomial = orig
    obs = state[:, 0].unsqueeze(0)  # (1, 2, 1, 13, 13)
    v, a, ent, lp, (h, c), pred = mine((obs, (hx.unsqueeze(0), cx.unsqueeze(0))), False, torch.tensor([forced]))
    tol = dict(rtol=1e-5, atol=1e-6)
    assert torch.allclose(v[0], rv.view(2), **tol) and torch.allclose(ent[0], rent.view(2), **tol) and torch.allclose(lp[0], rlp.view(2), **tol)
    assert torch.allclose(h[0], rh, **tol) and torch.allclose(c[0], rc, **tol)
    if "tat" in network:
        assert torch.allclose(pred[0], rpred.view(1), **tol)


def test_reference_init_statistics_match():
    """weights_init (utils.py:47-62) is the last init applied: per-tensor bounds must agree with the reference's."""
    ref.load_reference()
    from environment import create_env
    from model import build_model as ref_build
    from active_tracking_rl_b200.model import build_model
    from active_tracking_rl_b200.spaces import Box, Discrete
    from active_tracking_rl_b200.train import default_args
    args = ref.RefArgs()
    env = create_env("Track2D-BlockPartialPZR-v0", args)
    rmodel = ref_build(env.observation_space, env.action_space, args, torch.device("cpu"))
    mine = build_model([Box(0, 6, (1, 13, 13))] * 2, [Discrete(4)] * 2, default_args(), torch.device("cpu"))
    for (k, a), (k2, b) in zip(rmodel.state_dict().items(), mine.state_dict().items()):
        assert k == k2 and a.shape == b.shape
        if "lstm.weight" in k or ("weight" in k and a.numel() > 500):
            assert abs(float(a.abs().max()) - float(b.abs().max())) < 0.1 * float(a.abs().max()), k
            assert abs(float(a.std()) - float(b.std())) < 0.1 * float(a.std()), k
        if k.endswith("bias"):
            assert float(a.abs().max()) == 0.0 and float(b.abs().max()) == 0.0, k
