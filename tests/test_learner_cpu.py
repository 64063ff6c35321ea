"""CPU tests of the batched policy: reference-format state_dict, per-env forward parity with the learner
oracle (float32 CPU both sides; tolerance 1e-5), sampling semantics."""
import numpy as np
import pytest
import torch

import a3c_oracle
from active_tracking_rl_b200.model import build_model
from active_tracking_rl_b200.spaces import Box, Discrete
from active_tracking_rl_b200.train import default_args

OBS = [Box(0, 6, (1, 13, 13)), Box(0, 6, (1, 13, 13))]
ACT = [Discrete(4), Discrete(4)]


def rand_obs(E, seed=0):
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.choice([0, 1, 2, 4], size=(E, 2, 1, 13, 13), p=[0.7, 0.2, 0.05, 0.05]).astype(np.float32))


@pytest.mark.parametrize("network,tat", [("tat-maze-lstm", True), ("maze-lstm", False)])
def test_state_dict_is_reference_format(network, tat):
    model = build_model(OBS, ACT, default_args(network=network), torch.device("cpu"))
    sd = a3c_oracle.det_state_dict(tat=tat)
    assert list(model.state_dict().keys()) == list(sd.keys())  # same names, same ORDER as the reference registers them
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=True)
    n = sum(p.numel() for p in model.parameters())
    assert n == (801291 if tat else 668810)


def test_initialisation_follows_weights_init():
    model = build_model(OBS, ACT, default_args(), torch.device("cpu"))
    for name, p in model.named_parameters():
        if name.endswith("bias"):
            assert float(p.abs().max()) == 0.0, name  # utils.py:54,62 and model.py:117-118
    w = model.player1.encoder.fc.weight
    bound = np.sqrt(6.0 / (1024 + 256))
    assert float(w.abs().max()) <= bound and float(w.abs().max()) > 0.95 * bound


@pytest.mark.parametrize("network,tat", [("tat-maze-lstm", True), ("maze-lstm", False)])
def test_batched_forward_equals_oracle_per_env(network, tat):
    E = 5
    model = build_model(OBS, ACT, default_args(network=network), torch.device("cpu"))
    sd = a3c_oracle.det_state_dict(tat=tat)
    model.load_state_dict(sd)
    obs = rand_obs(E)
    g = torch.Generator().manual_seed(1)
    hx, cx = torch.randn(E, 2, 128, generator=g) * 0.3, torch.randn(E, 2, 128, generator=g) * 0.3
    forced = torch.randint(0, 4, (E, 2), generator=g)
    v, a, ent, lp, (h2, c2), rp = model((obs, (hx, cx)), False, forced)
    assert a.dtype == torch.int64 and torch.equal(a, forced)
    for e in range(E):
        ov, oa, oent, olp, (oh, oc), orp = a3c_oracle.forward(sd, obs[e].unsqueeze(1), hx[e], cx[e], tat, forced=forced[e].tolist())
        tol = dict(rtol=1e-5, atol=1e-6)
        assert torch.allclose(v[e], ov.view(2), **tol) and torch.allclose(ent[e], oent.view(2), **tol)
        assert torch.allclose(lp[e], olp.view(2), **tol)
        assert torch.allclose(h2[e], oh, **tol) and torch.allclose(c2[e], oc, **tol)
        if tat:
            assert torch.allclose(rp[e], orp.view(1), **tol)
        else:
            assert rp is None
    # greedy mode returns argmax actions and the full log-prob rows (model.py:46-47)
    v, a, ent, lp, _, _ = model((obs, (hx, cx)), True)
    assert lp.shape == (E, 2, 4) and torch.equal(a, lp.argmax(-1))


def test_sampling_follows_the_policy_distribution():
    torch.manual_seed(0)
    model = build_model(OBS, ACT, default_args(), torch.device("cpu"))
    E = 4000
    obs = rand_obs(1).expand(E, 2, 1, 13, 13).contiguous()
    hx, cx = torch.zeros(E, 2, 128), torch.zeros(E, 2, 128)
    with torch.no_grad():
        _, a, _, _, _, _ = model((obs, (hx, cx)))
        _, _, _, lp, _, _ = model((obs[:1], (hx[:1], cx[:1])), True)
    p0 = lp[0, 0].exp().numpy()
    freq = np.bincount(a[:, 0].numpy(), minlength=4) / E
    assert np.abs(freq - p0).max() < 0.03


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        build_model(OBS, ACT, default_args(network="cnn-lstm"), torch.device("cpu"))
    with pytest.raises(NotImplementedError):
        build_model(OBS, ACT, default_args(network="maze-lstm-continuous"), torch.device("cpu"))


def test_fused_heads_and_linear_fall_back_to_torch_on_cpu():
    """model._heads / model._linear / gemm.colsum take torch's own ops for CPU tensors (host-side tests never touch the CUDA
    library) and return exactly what the separate nn.Linear modules do"""
    import torch
    from active_tracking_rl_b200 import gemm, model as M
    torch.manual_seed(0)
    actor, critic, aux = torch.nn.Linear(128, 4), torch.nn.Linear(128, 1), torch.nn.Linear(128, 1)
    hx = torch.randn(6, 128)
    logit, value, r = M._heads(hx, actor, critic, aux)
    assert torch.equal(logit, actor(hx)) and torch.equal(value, critic(hx)) and torch.equal(r, aux(hx))
    fc = torch.nn.Linear(512, 256)
    x = torch.randn(6, 512)
    assert torch.equal(M._linear(x, fc, relu=True), torch.relu(fc(x)))
    assert torch.equal(gemm.colsum(x), x.sum(0))
    assert not gemm.supported(x, fc.weight)
