"""The hand-written forward / backward of the batched learner (active_tracking_rl_b200/learner.py over csrc/track2d_a3c.cu) against
plain PyTorch float32 references of the same ops (kernel level) and against the autograd implementation in model.py /
player_util.py driven through the same rollout (system level).  Float32 vs float32: relative tolerance 1e-5 on activations,
2e-4 on losses, 1e-3 on gradients (different summation orders over up to E x T rows)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.fixture(scope="module")
def lib():
    from active_tracking_rl_b200 import _lib
    return _lib.load()


@pytest.mark.parametrize("E,mode", [(1, "forced"), (37, "forced"), (1000, "greedy"), (4096, "sample")])
def test_lstm_heads_forward_matches_torch(lib, E, mode):
    from active_tracking_rl_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(E)
    r = lambda *s: torch.randn(*s, generator=g, device=DEV)  # noqa: E731
    gates, b_ih, b_hh, c_prev = r(E, 512) * 1.5, r(512) * 0.1, r(512) * 0.1, r(E, 128)
    w_head, b_head = r(8, 128) * 0.3, r(8) * 0.1
    act, c_next, h_out = torch.zeros(E, 512, device=DEV), torch.zeros(E, 128, device=DEV), torch.zeros(E, 128, device=DEV)
    xh_next = torch.zeros(E, 384, device=DEV)
    out8 = torch.zeros(E, 8, device=DEV)
    action = torch.full((E, 2), -1, dtype=torch.int32, device=DEV)
    forced = torch.randint(0, 4, (E, 2), generator=g, device=DEV, dtype=torch.int32)
    value, logp, ent = [torch.zeros(E, 2, device=DEV) for _ in range(3)]
    logp_all = torch.zeros(E, 4, device=DEV)
    step = torch.tensor([5], dtype=torch.int64, device=DEV)
    col = 1
    off = lambda t: C.c_void_p(t.data_ptr() + 4 * col)  # noqa: E731
    _lib.check(lib.track2d_lstm_heads_forward(_p(gates), _p(b_ih), _p(b_hh), _p(c_prev), _p(act), _p(c_next), _p(h_out), C.c_void_p(xh_next.data_ptr() + 1024), 384,
                                              _p(w_head), _p(b_head), _p(out8), off(action), off(forced) if mode == "forced" else None, off(value), off(logp),
                                              off(ent), _p(logp_all) if mode == "greedy" else None, _p(step), 1234, 1, int(mode == "greedy"), E, 0, _stream()), lib)
    z = gates + b_ih + b_hh
    i, f, gg, o = z.chunk(4, 1)
    c2 = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(gg)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    ref8 = h2.double() @ w_head.double().t() + b_head.double()
    assert torch.allclose(c_next, c2, rtol=1e-5, atol=1e-6) and torch.allclose(h_out, h2, rtol=1e-5, atol=1e-6)
    assert torch.equal(xh_next[:, 256:], h_out) and float(xh_next[:, :256].abs().max()) == 0
    assert torch.allclose(act, torch.cat([torch.sigmoid(i), torch.sigmoid(f), torch.tanh(gg), torch.sigmoid(o)], 1), rtol=1e-5, atol=1e-6)
    assert torch.allclose(out8.double(), ref8, rtol=1e-5, atol=2e-5)
    lsm = F.log_softmax(out8[:, :4], 1)
    prob = F.softmax(out8[:, :4], 1)
    a = action[:, col].long()
    assert (action[:, 0] == -1).all(), "the other agent's column must stay untouched"
    assert torch.allclose(value[:, col], out8[:, 4]) and float(value[:, 0].abs().max()) == 0
    assert torch.allclose(ent[:, col], -(lsm * prob).sum(1), rtol=1e-5, atol=1e-6)
    assert torch.allclose(logp[:, col], lsm.gather(1, a.view(-1, 1)).squeeze(1), rtol=1e-5, atol=1e-6)
    if mode == "forced":
        assert torch.equal(a, forced[:, col].long())
    elif mode == "greedy":
        assert torch.equal(a, prob.argmax(1)) and torch.allclose(logp_all, lsm, rtol=1e-5, atol=1e-6)
    else:
        # the sampler draws from softmax(logits): frequencies per probability bucket, and a different counter gives other draws
        assert a.min() >= 0 and a.max() <= 3
        chosen_p = prob.gather(1, a.view(-1, 1)).squeeze(1)
        assert abs(float(chosen_p.mean()) - float((prob * prob).sum(1).mean())) < 0.03  # E[p_a] = sum p^2 under the right law
        action2 = torch.zeros_like(action)
        step.fill_(6)
        _lib.check(lib.track2d_lstm_heads_forward(_p(gates), _p(b_ih), _p(b_hh), _p(c_prev), None, _p(c_next), None, None, 384, _p(w_head), _p(b_head), None,
                                                  off(action2), None, None, None, None, None, _p(step), 1234, 1, 0, E, 0, _stream()), lib)
        assert 0.3 < float((action2[:, col] != action[:, col]).float().mean()) < 0.9


def test_post_step_and_embed_add(lib):
    from active_tracking_rl_b200 import _lib
    E = 1003
    g = torch.Generator(device=DEV).manual_seed(0)
    done = (torch.rand(E, generator=g, device=DEV) < 0.3).to(torch.uint8)
    xh0, xh1 = torch.randn(E, 384, generator=g, device=DEV), torch.randn(E, 384, generator=g, device=DEV)
    c0, c1 = torch.randn(E, 128, generator=g, device=DEV), torch.randn(E, 128, generator=g, device=DEV)
    ref = [t.clone() for t in (xh0, xh1, c0, c1)]
    eps = torch.arange(E, dtype=torch.int32, device=DEV)
    step = torch.tensor([41], dtype=torch.int64, device=DEV)
    _lib.check(lib.track2d_policy_post_step(_p(done), C.c_void_p(xh0.data_ptr() + 1024), C.c_void_p(xh1.data_ptr() + 1024), 384, _p(c0), _p(c1), _p(eps), _p(step), E,
                                            _stream()), lib)
    d = done.bool()
    for got, want, is_xh in ((xh0, ref[0], True), (xh1, ref[1], True), (c0, ref[2], False), (c1, ref[3], False)):
        want = want.clone()
        if is_xh:
            want[d, 256:] = 0
        else:
            want[d] = 0
        assert torch.equal(got, want)
    assert torch.equal(eps, torch.where(d, torch.zeros_like(eps), torch.arange(E, dtype=torch.int32, device=DEV) + 1)) and int(step.item()) == 42
    # embed_add
    x = torch.randn(E, 256, generator=g, device=DEV)
    out = torch.zeros(E, 384, device=DEV)
    w, b = torch.randn(256, 4, generator=g, device=DEV), torch.randn(256, generator=g, device=DEV)
    acts = torch.randint(0, 4, (E, 2), generator=g, device=DEV, dtype=torch.int32)
    _lib.check(lib.track2d_embed_add(_p(x), 256, _p(out), 384, _p(w), _p(b), _p(acts), 256, E, _stream()), lib)
    want = x + F.linear(F.one_hot(acts[:, 0].long(), 4).float(), w, b)
    assert torch.allclose(out[:, :256], want, rtol=1e-6, atol=1e-6) and float(out[:, 256:].abs().max()) == 0


@pytest.mark.parametrize("mode,use_aux", [(-1, 1), (0, 1), (1, 0)])
def test_a3c_loss_grad_matches_autograd(lib, mode, use_aux):
    """returns / GAE recursion + loss statistics + d(loss)/d(head outputs) vs player_util.py:117-155 written with torch autograd"""
    from active_tracking_rl_b200 import _lib
    T, E, gamma, tau, w_ent = 20, 257, 0.9, 1.0, (0.01, 0.2)
    g = torch.Generator(device=DEV).manual_seed(3)
    out8 = [(torch.randn(T + 1, E, 8, generator=g, device=DEV)).requires_grad_(True) for _ in range(2)]
    actions = torch.randint(0, 4, (T, E, 2), generator=g, device=DEV, dtype=torch.int32)
    rew = torch.randn(T, E, 2, generator=g, device=DEV)
    done = (torch.rand(T, E, generator=g, device=DEV) < 0.07).to(torch.uint8)
    dout = [torch.zeros(T, E, 8, device=DEV) for _ in range(2)]
    stats = torch.zeros(7, E, device=DEV)
    ret, gae = torch.zeros(T, E, 2, device=DEV), torch.zeros(T, E, 2, device=DEV)
    scale = 1.0 / E
    _lib.check(lib.track2d_a3c_loss_grad(_p(out8[0]), _p(out8[1]), _p(dout[0]), _p(dout[1]), _p(actions), _p(rew), _p(done), _p(stats), _p(ret), _p(gae), T, E,
                                         gamma, tau, w_ent[0], w_ent[1], scale, int(mode in (-1, 0)), int(mode in (-1, 1)), 2 if (use_aux and mode != 0) else 1, _stream()), lib)
    # reference
    vals = torch.stack([out8[0][:, :, 4], out8[1][:, :, 4]], 2)  # (T+1, E, 2)
    R, G, vnext = vals[T].detach().clone(), torch.zeros(E, 2, device=DEV), vals[T].detach().clone()
    Rs, Gs = [None] * T, [None] * T
    for t in reversed(range(T)):
        d = done[t].bool().unsqueeze(1)
        R, vnext, G = torch.where(d, 0 * R, R), torch.where(d, 0 * vnext, vnext), torch.where(d, 0 * G, G)
        R = gamma * R + rew[t]
        G = G * gamma * tau + rew[t] + gamma * vnext - vals[t].detach()
        vnext = vals[t].detach()
        Rs[t], Gs[t] = R, G
    Rs, Gs = torch.stack(Rs), torch.stack(Gs)
    assert torch.allclose(ret, Rs, rtol=1e-5, atol=1e-5) and torch.allclose(gae, Gs, rtol=1e-5, atol=1e-5)
    pls, vls, ents = [], [], []
    for a in range(2):
        lsm, prob = F.log_softmax(out8[a][:T, :, :4], 2), F.softmax(out8[a][:T, :, :4], 2)
        ent = -(lsm * prob).sum(2)
        lp = lsm.gather(2, actions[:, :, a].long().unsqueeze(2)).squeeze(2)
        pls.append((-(lp * Gs[:, :, a]) - w_ent[a] * ent).sum(0))
        vls.append((0.5 * (Rs[:, :, a] - vals[:T, :, a]).pow(2)).sum(0))
        ents.append(ent.sum(0))
    prl = (out8[1][:T, :, 5] - rew[:, :, 0]).abs().sum(0)
    loss = 0
    if mode in (-1, 0):
        loss = loss + pls[0] + 0.5 * vls[0]
    if mode in (-1, 1):
        loss = loss + pls[1] + 0.5 * vls[1]
    if use_aux and mode != 0:
        loss = loss + prl
    g0, g1 = torch.autograd.grad(loss.sum() * scale, out8, allow_unused=True)
    for a, gr in enumerate((g0, g1)):
        gr = torch.zeros_like(out8[a]) if gr is None else gr
        assert float(gr[T].abs().max()) == 0
        assert torch.allclose(dout[a], gr[:T], rtol=1e-4, atol=1e-7), (a, (dout[a] - gr[:T]).abs().max())
        assert torch.allclose(stats[a], pls[a], rtol=2e-5, atol=2e-5) and torch.allclose(stats[2 + a], vls[a], rtol=2e-5, atol=2e-5)
        assert torch.allclose(stats[4 + a], ents[a], rtol=2e-5, atol=2e-5)
    assert torch.allclose(stats[6], prl, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("E", [1, 29, 2048])
def test_lstm_heads_backward_matches_autograd(lib, E):
    from active_tracking_rl_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(E)
    r = lambda *s: torch.randn(*s, generator=g, device=DEV)  # noqa: E731
    gates = (r(E, 512) * 1.5).requires_grad_(True)
    c_prev = r(E, 128).requires_grad_(True)
    w_head = r(8, 128) * 0.3
    dout8, dh_rec, dc_rec = r(E, 8), r(E, 128), r(E, 128)
    done = (torch.rand(E, generator=g, device=DEV) < 0.3).to(torch.uint8)
    i, f, gg, o = gates.chunk(4, 1)
    si, sf, tg, so = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(gg), torch.sigmoid(o)
    c2 = sf * c_prev + si * tg
    h2 = so * torch.tanh(c2)
    keep = (1 - done.float()).unsqueeze(1)
    obj = ((h2 @ w_head.t()) * dout8).sum() + (h2 * dh_rec * keep).sum() + (c2 * dc_rec * keep).sum()
    dg_ref, dc_ref = torch.autograd.grad(obj, [gates, c_prev], retain_graph=True)
    act = torch.cat([si, sf, tg, so], 1).detach().contiguous()
    dgates = torch.zeros(E, 512, device=DEV)
    dc = dc_rec.clone()
    _lib.check(lib.track2d_lstm_heads_backward(_p(dout8), _p(w_head), _p(dh_rec), _p(dc), _p(done), _p(act), _p(c_prev.detach()), _p(dgates), E, _stream()), lib)
    assert torch.allclose(dgates, dg_ref, rtol=1e-4, atol=1e-5), (dgates - dg_ref).abs().max()
    assert torch.allclose(dc, dc_ref, rtol=1e-4, atol=1e-5)
    # last step of the sweep: no recurrent gradient at all
    obj2 = ((h2 @ w_head.t()) * dout8).sum()
    dg2, dc2 = torch.autograd.grad(obj2, [gates, c_prev])
    dc = torch.full((E, 128), 7.0, device=DEV)
    _lib.check(lib.track2d_lstm_heads_backward(_p(dout8), _p(w_head), None, _p(dc), None, _p(act), _p(c_prev.detach()), _p(dgates), E, _stream()), lib)
    assert torch.allclose(dgates, dg2, rtol=1e-4, atol=1e-5) and torch.allclose(dc, dc2, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("M,grouped", [(5, True), (1000, True), (40001, False), (40001, True)])
def test_relu_backward_groupsum_matches_torch(lib, M, grouped):
    from active_tracking_rl_b200 import _lib
    N = 256
    g = torch.Generator(device=DEV).manual_seed(M)
    dy = torch.randn(M, N, generator=g, device=DEV)
    ybuf = torch.randn(M, 384, generator=g, device=DEV)
    grp = torch.randint(0, 4, (M, 2), generator=g, device=DEV, dtype=torch.int32)
    want_dz = dy * (ybuf[:, :N] > 0)
    want_gw = torch.stack([(dy.double() * (grp[:, 0] == a).unsqueeze(1)).sum(0) for a in range(4)], 1)  # [N][4]
    n_ws = int(lib.track2d_relu_backward_workspace_floats(M, N))
    ws = torch.empty(n_ws, device=DEV)
    db, gw, gb = torch.zeros(N, device=DEV), torch.zeros(N, 4, device=DEV), torch.zeros(N, device=DEV)
    got = dy.clone()
    _lib.check(lib.track2d_relu_backward_groupsum(_p(got), _p(ybuf), 384, _p(grp) if grouped else None, M, N, _p(db), _p(gw) if grouped else None,
                                                  _p(gb) if grouped else None, _p(ws), n_ws, _stream()), lib)
    assert torch.equal(got, want_dz)
    tol = 1e-5 * float(np.sqrt(M)) * 4
    assert torch.allclose(db.double(), want_dz.double().sum(0), rtol=1e-4, atol=tol)
    if grouped:
        assert torch.allclose(gw.double(), want_gw, rtol=1e-4, atol=tol) and torch.allclose(gb.double(), dy.double().sum(0), rtol=1e-4, atol=tol)


@pytest.mark.parametrize("N", [1, 9, 2051])
def test_conv_stack_uint8_strided_matches_float32(lib, N):
    """the image-source variants of the conv stack: uint8 / strided images run conv2 on the tensor cores (3xTF32, track2d_conv_tc.cu), float32 /
    packed images on the FP32 CUDA-core kernels (track2d_policy.cu) -- same results to float32 accuracy"""
    from active_tracking_rl_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(N)
    obs = torch.tensor([0, 1, 2, 4], device=DEV, dtype=torch.uint8)[torch.randint(0, 4, (N, 2, 169), generator=g, device=DEV)]
    w1, b1 = torch.rand(16, 1, 3, 3, generator=g, device=DEV) - 0.5, torch.rand(16, generator=g, device=DEV) - 0.5
    w2, b2 = (torch.rand(32, 16, 3, 3, generator=g, device=DEV) - 0.5) * 0.3, torch.rand(32, generator=g, device=DEV) - 0.5
    gy = torch.randn(N, 512, generator=g, device=DEV)
    for agent in (0, 1):
        xf = obs[:, agent].float().contiguous()
        y_ref, y_got = torch.zeros(N, 512, device=DEV), torch.zeros(N, 512, device=DEV)
        _lib.check(lib.track2d_maze_conv_forward(_p(xf), N, _p(w1), _p(b1), _p(w2), _p(b2), _p(y_ref), _stream()), lib)
        _lib.check(lib.track2d_maze_conv_forward_ex(C.c_void_p(obs.data_ptr() + 169 * agent), 1, 338, N, _p(w1), _p(b1), _p(w2), _p(b2), _p(y_got), _stream()), lib)
        assert torch.allclose(y_ref, y_got, rtol=3e-5, atol=3e-5), (y_ref - y_got).abs().max()
        if N in (9, 2051):
            gr = [torch.zeros_like(t) for t in (w1, b1, w2, b2)]
            gg = [torch.zeros_like(t) for t in (w1, b1, w2, b2)]
            _lib.check(lib.track2d_maze_conv_backward(_p(xf), _p(y_ref), _p(gy), N, _p(w1), _p(b1), _p(w2), *[_p(t) for t in gr], _stream()), lib)
            _lib.check(lib.track2d_maze_conv_backward_ex(C.c_void_p(obs.data_ptr() + 169 * agent), 1, 338, _p(y_ref), _p(gy), N, _p(w1), _p(b1), _p(w2),
                                                         *[_p(t) for t in gg], _stream()), lib)
            for a, b in zip(gr, gg):
                assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max()) + 1e-6, (float((a - b).abs().max()), float(a.abs().max()))


@pytest.mark.parametrize("env_id,network,aux,mode,E", [("Track2D-BlockPartialPZR-v0", "tat-maze-lstm", "reward", -1, 64),
                                                       ("Track2D-BlockPartialPZR-v0", "tat-maze-lstm", "reward", -1, 6),
                                                       ("Track2D-MazePartialAdv-v0", "maze-lstm", "none", -1, 40),
                                                       ("Track2D-BlockPartialRam-v0", "tat-maze-lstm", "reward", 0, 24),
                                                       ("Track2D-BlockPartialPZR-v0", "tat-maze-lstm", "reward", 1, 24)])
def test_fused_learner_matches_autograd_learner(env_id, network, aux, mode, E):
    """two Trainers on identical worlds (same seed) and weights, one through learner.FusedA3C, one through model.py's autograd path,
    driven with the same forced actions: same values / log-probs / entropies per step, same loss statistics, same gradient, same
    updated weights, same recurrent state carried into the next window (which includes episode ends and auto-resets)"""
    from active_tracking_rl_b200.train import Trainer, default_args
    T = 20
    mk = lambda fused: Trainer(default_args(env=env_id, network=network, aux=aux, train_mode=mode, num_envs=E, num_steps=T, seed=5, fused=fused,  # noqa: E731
                                            entropy_target=0.2 if "tat" in network else 0.01, max_grad_norm=0.0), DEV)
    ta, tb = mk(True), mk(False)
    assert ta.player.fused and not tb.player.fused
    tb.model.load_state_dict(ta.model.state_dict())
    rs = np.random.RandomState(1)
    n_done = 0
    for it in range(3):
        ta.player.update_rnn_hiden()
        tb.player.update_rnn_hiden()
        for t in range(T):
            f = rs.randint(0, 4, size=(E, 2))
            if t % 3 and it > 0:
                f[:, 0] = 0
            f = torch.from_numpy(f).to(DEV)
            ta.player.action_train(f)
            tb.player.action_train(f)
            assert torch.equal(ta.player.obs_buf[t + 1].float(), tb.player.obs_buf[t + 1]) and torch.equal(ta.player.done_buf[t], tb.player.done_buf[t])
            eng = ta.player.engine
            assert torch.allclose(eng.values[t], tb.player.values[-1], rtol=1e-4, atol=2e-5), (it, t)
            assert torch.allclose(eng.logp[t], tb.player.log_probs[-1], rtol=1e-4, atol=2e-5), (it, t)
            assert torch.allclose(eng.entropy[t], tb.player.entropies[-1], rtol=1e-4, atol=2e-5), (it, t)
        n_done += int(ta.player.done_buf.sum())
        boot = torch.from_numpy(rs.randint(0, 4, size=(E, 2))).to(DEV)
        sa = ta.player.optimize(None, ta.optimizer, ta.model, mode, None, boot_forced_actions=boot, apply=False)
        sb = tb.player.optimize(None, tb.optimizer, tb.model, mode, None, boot_forced_actions=boot, apply=False)
        for x, y in zip(sa, sb):
            assert torch.allclose(x, y, rtol=2e-4, atol=2e-4), (it, (x - y).abs().max())
        ga, gb = ta.optimizer.fp.grad, tb.optimizer.fp.grad
        for (name, pa), (_, pb) in zip(ta.model.named_parameters(), tb.model.named_parameters()):
            if pa.grad is None:
                continue
            # (the two paths run conv2 with different arithmetic -- 3xTF32 on the tensor cores vs FP32 FMA -- so activations differ by ~5e-6
            # and a ReLU gate sitting at zero can flip in one of them: with only E x T = a few hundred images that moves a conv weight
            # gradient by up to a few 1e-3 of its scale; same tolerance as the oracle comparison in test_gpu_learner.py)
            scale = float(pb.grad.abs().max()) + 1e-8
            assert float((pa.grad - pb.grad).abs().max()) <= 5e-3 * scale + 1e-7, (it, name, float((pa.grad - pb.grad).abs().max()), scale)
        assert float((ga - gb).norm()) <= 2e-3 * float(gb.norm()) + 1e-7
        ta.player.apply_update(ta.optimizer)
        tb.player.apply_update(tb.optimizer)
        assert torch.allclose(ta.optimizer.fp.flat, tb.optimizer.fp.flat, rtol=0, atol=2e-5)
        assert torch.allclose(ta.player.hxs, tb.player.hxs, rtol=1e-4, atol=1e-5) and torch.allclose(ta.player.cxs, tb.player.cxs, rtol=1e-4, atol=1e-5)
        tb.model.load_state_dict(ta.model.state_dict())  # stay in lock step (updates agree to 2e-5, not bit for bit)
        tb.player.hx_store.copy_(ta.player.hxs)
        tb.player.cx_store.copy_(ta.player.cxs)
    assert n_done > 0
    ta.env.close()
    tb.env.close()


def test_fused_sampling_trains_and_replays_as_a_cuda_graph():
    """the sampled (not forced) fused path end to end: losses fall, weights move on every replay, actions are distributed like pi"""
    from active_tracking_rl_b200.train import Trainer, default_args
    tr = Trainer(default_args(num_envs=1024, seed=4), DEV)
    assert tr.player.fused
    vls = []
    for _ in range(3):
        vls.append(float(tr.iteration()[1].mean()))
    eng = tr.player.engine
    a = eng.actions[:20].reshape(-1, 2).long()
    freq = torch.stack([torch.bincount(a[:, k], minlength=4).float() / a.shape[0] for k in range(2)])
    prob = torch.stack([F.softmax(eng.nets[k].out8[:20, :, :4], 2).mean((0, 1)) for k in range(2)])
    assert float((freq - prob).abs().max()) < 0.02, (freq, prob)
    tr.capture(warmup=1)
    w_prev = tr.optimizer.fp.flat.clone()
    for k in range(30):
        pl, vl, ent, prl = tr.replay()
        vls.append(float(vl.mean()))
        if k % 10 == 9:
            assert float((tr.optimizer.fp.flat - w_prev).abs().max()) > 0
            w_prev = tr.optimizer.fp.flat.clone()
    assert np.isfinite(vls).all() and np.mean(vls[-5:]) < np.mean(vls[:3]), vls
    assert int(eng.rng_step.item()) == 34 * 20, int(eng.rng_step.item())
    assert tr.env.status() == 0
    tr.env.close()


@pytest.mark.parametrize("network,E,slices", [("tat-maze-lstm", 50000, ((0, 24576), (24576, 50000))), ("maze-lstm", 50000, ((0, 25000), (25000, 50000))),
                                              ("tat-maze-lstm", 1000, ((0, 304), (304, 305), (305, 1000)))])
def test_forward_in_env_slices_equals_whole_batch_forward(network, E, slices):
    """FusedA3C.forward(envs=(first, end)): every kernel is row-wise and the sampler's Philox counter is the env index, so the slices of
    a batch add up to the whole-batch policy step, sampled actions included.  Bit for bit when the GEMM launch plan of the slices is the
    whole batch's (>= 96 super-tiles: no split-K, which is the case of the pipelined host path at the bench sizes); with split-K the
    reduction order over K differs and the values agree to rounding."""
    from active_tracking_rl_b200.train import Trainer, default_args
    T = 3
    exact = E >= 49152
    mk = lambda: Trainer(default_args(network=network, aux="reward" if "tat" in network else "none", num_envs=E, num_steps=T, seed=11), DEV)  # noqa: E731
    ta, tb = mk(), mk()
    ea, eb = ta.player.engine, tb.player.engine
    same = (lambda x, y: torch.equal(x, y)) if exact else (lambda x, y: torch.allclose(x, y, rtol=1e-4, atol=1e-5))
    for t in range(T):
        ea.forward(t)
        for lo, hi in slices:
            eb.forward(t, envs=(lo, hi))
        if exact:
            assert torch.equal(ea.actions[t], eb.actions[t])
        else:  # a probability that moved in the last bit can move a sample across its threshold: vanishingly rare
            assert float((ea.actions[t] != eb.actions[t]).float().mean()) < 0.002
            eb.actions[t].copy_(ea.actions[t])
            eb.logp[t].copy_(ea.logp[t])
        assert same(ea.values[t], eb.values[t]) and same(ea.logp[t], eb.logp[t]) and same(ea.entropy[t], eb.entropy[t])
        for na, nb in zip(ea.nets, eb.nets):
            assert same(na.xh[t + 1], nb.xh[t + 1]) and same(na.c[t + 1], nb.c[t + 1]) and same(na.out8[t], nb.out8[t])
            assert same(na.act[t], nb.act[t]) and torch.equal(na.convout[t], nb.convout[t])
        for tr in (ta, tb):
            p = tr.player
            p.env.step_into(p.engine.actions[t], p.obs_buf[t + 1], p.rew_buf[t], p.done_buf[t])
            p.engine.post_step(t, p.done_buf[t])
        assert torch.equal(ta.player.obs_buf[t + 1], tb.player.obs_buf[t + 1])
    with pytest.raises(ValueError):
        eb.forward(0, envs=(10, E + 1))
    ta.env.close()
    tb.env.close()


def test_host_rollout_of_a_small_batch_equals_device_rollout():
    """the defaults of the host path for a small batch (one transfer chunk, the prefetched policy step over the whole batch)"""
    from active_tracking_rl_b200.train import Trainer, default_args
    mk = lambda: Trainer(default_args(num_envs=1003, num_steps=5, seed=4), DEV)  # noqa: E731
    ta, tb = mk(), mk()
    hb = tb.env.alloc_host_buffers(obs_dtype=torch.float32)
    for it in range(3):
        sa, sb = ta.iteration(), tb.iteration(host=hb)
        for x, y in zip(sa, sb):
            assert torch.equal(x, y), it
        assert torch.equal(ta.optimizer.fp.flat, tb.optimizer.fp.flat), it
    assert ta.env.status() == 0 and tb.env.status() == 0
    ta.env.close()
    tb.env.close()


@pytest.mark.parametrize("obs_dtype,chunks", [(torch.float32, 8), (torch.uint8, 4)])
def test_pipelined_host_rollout_equals_device_rollout(obs_dtype, chunks):
    """Agent.action_train(host=...): env.step through the host-buffer ABI with the next policy step started on the first half of the envs
    while the second half's observations are still arriving -- same actions, rewards, statistics and updated weights, bit for bit, as
    the device-resident rollout and as the host path without the prefetch"""
    from active_tracking_rl_b200.train import Trainer, default_args
    E, T = 50000, 5
    mk = lambda: Trainer(default_args(num_envs=E, num_steps=T, seed=3), DEV)  # noqa: E731
    ta, tb, tc = mk(), mk(), mk()
    hb, hc = tb.env.alloc_host_buffers(obs_dtype=obs_dtype), tc.env.alloc_host_buffers(obs_dtype=obs_dtype)
    hb["chunks"], hb["forward_slices"] = chunks, 2
    hc["prefetch"] = False
    for it in range(3):
        sa, sb, sc = ta.iteration(), tb.iteration(host=hb), tc.iteration(host=hc)
        for x, y, z in zip(sa, sb, sc):
            assert torch.equal(x, y) and torch.equal(x, z), it
        assert torch.equal(ta.optimizer.fp.flat, tb.optimizer.fp.flat) and torch.equal(ta.optimizer.fp.flat, tc.optimizer.fp.flat), it
    assert ta.env.status() == 0 and tb.env.status() == 0 and tc.env.status() == 0
    for tr in (ta, tb, tc):
        tr.env.close()
