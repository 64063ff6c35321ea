"""Pins the learner oracle (oracle/a3c_oracle.py) to the reference: replays tests/golden/learner_*.npz --
recorded from the unmodified model.py / player_util.py / shared_optim.py by
oracle/refharness/make_golden_learner.py -- and checks per-step values / log-probs / entropies / reward
predictions, the loss tensors, per-tensor gradients and the weights after every SharedAdam step.
Floating point: float32 CPU vs float32 CPU of the same operations; tolerances stated per check."""
import glob
import os

import numpy as np
import pytest
import torch

import a3c_oracle
from conftest import GOLDEN

FILES = sorted(glob.glob(os.path.join(GOLDEN, "learner_*.npz")))


def replay(g, on_iteration=None):
    env_id, network, aux, train_mode, ent_t, env_seed, scale = [str(x) for x in g["meta"]]
    tat, train_mode = "tat" in network, int(train_mode)
    sd = a3c_oracle.det_state_dict(tat=tat, seed=1234, scale=float(scale))
    adam_state = {}
    hx, cx = torch.zeros(2, 128), torch.zeros(2, 128)
    step, oi = 0, 0
    opt_names = [n for n in sd if train_mode == -1 or n.startswith("player%d." % train_mode)]
    for it in range(len(g["iter_len"])):
        local = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        if g["reset_before"][it]:
            hx, cx = torch.zeros(2, 128), torch.zeros(2, 128)
        hx, cx = hx.detach(), cx.detach()
        values, log_probs, entropies, preds, rewards = [], [], [], [], []
        n = int(g["iter_len"][it])
        state = torch.from_numpy(g["obs"][oi].astype(np.float32)).view(2, 1, 1, 13, 13)
        oi += 1
        for i in range(n):
            v, acts, ent, lp, (hx, cx), rp = a3c_oracle.forward(local, state, hx, cx, tat, forced=g["actions"][step])
            # forward pass: float32 vs float32 of the same ops
            np.testing.assert_allclose(v.detach().numpy().reshape(2), g["values"][step], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(lp.detach().numpy().reshape(2), g["log_probs"][step], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(ent.detach().numpy().reshape(2), g["entropies"][step], rtol=1e-5, atol=1e-6)
            if tat:
                np.testing.assert_allclose(float(rp.item()), g["preds"][step], rtol=1e-5, atol=1e-6)
            values.append(v); log_probs.append(lp); entropies.append(ent); preds.append(rp)
            rewards.append(torch.from_numpy(g["rewards"][step]).float().unsqueeze(1))
            state = torch.from_numpy(g["obs"][oi].astype(np.float32)).view(2, 1, 1, 13, 13)
            oi += 1
            step += 1
        done = bool(g["iter_done"][it])
        R = torch.zeros(2, 1)
        if not done:
            with torch.no_grad():
                R = a3c_oracle.forward(local, state, hx, cx, tat, forced=g["boot_actions"][it])[0]
        loss, pl, vl, prl = a3c_oracle.segment_loss(values, log_probs, entropies, preds, rewards, R, 0.9, 1.0, 0.01, float(ent_t),
                                                    aux == "reward" and tat, train_mode)
        np.testing.assert_allclose(pl.detach().numpy().reshape(2), g["policy_loss"][it], rtol=2e-5, atol=1e-5)
        np.testing.assert_allclose(vl.detach().numpy().reshape(2), g["value_loss"][it], rtol=2e-5, atol=1e-5)
        np.testing.assert_allclose(float(prl.sum()), g["pred_loss"][it], rtol=2e-5, atol=1e-5)
        names = list(local)
        grads = torch.autograd.grad(loss, [local[k] for k in names], allow_unused=True)
        gd = {k: (gr if gr is not None else torch.zeros_like(local[k])) for k, gr in zip(names, grads)}
        gnames = [str(x) for x in g["grad_names"]]
        for j, k in enumerate(gnames):
            np.testing.assert_allclose(float(gd[k].norm()), g["grad_norm"][it][j], rtol=2e-4, atol=1e-6, err_msg="grad norm %s it %d" % (k, it))
            np.testing.assert_allclose(float(gd[k].double().sum()), g["grad_sum"][it][j], rtol=2e-3, atol=2e-4, err_msg="grad sum %s it %d" % (k, it))
        # the reference's clip_grad_norm_ is inert (exhausted generator): update with the raw gradient
        a3c_oracle.shared_adam_step({k: sd[k] for k in opt_names}, gd, adam_state)
        pnames = [str(x) for x in g["param_names"]]
        for j, k in enumerate(pnames):
            np.testing.assert_allclose(float(sd[k].double().sum()), g["param_sum"][it][j], rtol=1e-5, atol=2e-4, err_msg="param sum %s it %d" % (k, it))
            np.testing.assert_allclose(float(sd[k].double().norm()), g["param_norm"][it][j], rtol=1e-6, atol=1e-6, err_msg="param norm %s it %d" % (k, it))
        if on_iteration:
            on_iteration(it, gd)
    return sd


def test_learner_fixtures_present():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[8:-4] for p in FILES])
def test_learner_oracle_replays_reference(path):
    replay(np.load(path))


def test_reference_gradient_clip_is_inert():
    """target_only_bigw has total gradient norms of 1587 and 314 (>> 50) and the recorded weights still follow the
    UNCLIPPED SharedAdam update: clip_grad_norm_(params, 50) never clips in the reference (see a3c_oracle.Worker)."""
    g = np.load(os.path.join(GOLDEN, "learner_target_only_bigw.npz"))
    assert g["total_grad_norm"][0] > 1000 and g["total_grad_norm"][1] > 200
    replay(g)  # passes only with the unclipped update


def test_shared_adam_is_not_torch_adam():
    """eps = 1e-3 is added after the sqrt and is NOT scaled by the bias correction (shared_optim.py:161-173)."""
    p = {"w": torch.tensor([1.0, -2.0, 3.0])}
    gr = {"w": torch.tensor([0.5, 0.25, -1.0])}
    st = {}
    a3c_oracle.shared_adam_step(p, gr, st)
    m = 0.1 * gr["w"]
    v = 0.001 * gr["w"] ** 2
    step_size = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    expect = torch.tensor([1.0, -2.0, 3.0]) - step_size * m / (v.sqrt() + 1e-3)
    assert torch.allclose(p["w"], expect, rtol=1e-6, atol=0)
    q = torch.nn.Parameter(torch.tensor([1.0, -2.0, 3.0]))
    opt = torch.optim.Adam([q], lr=1e-3, eps=1e-3, amsgrad=True)
    q.grad = gr["w"].clone()
    opt.step()
    assert not torch.allclose(q.data, p["w"], rtol=1e-5, atol=0)
