"""FusedA3C -- the batched actor-learner's policy step and update without autograd.

`A3C_Dueling.forward` + `Agent.optimize` (model.py:212-264, player_util.py:108-161) for E envs, written out as explicit
forward / backward passes over libtrack2d kernels, because the structure of the problem is fixed and autograd hides it:

  * forward, per env-step and agent: conv stack -> fc GEMM (+bias, ReLU) [-> + tracker-action embedding] -> ONE gate GEMM over
    the concatenated [feature | h] input (K = 384) -> ONE fused kernel: LSTM cell, actor / critic / reward_aux heads, softmax,
    entropy, action sample (csrc/track2d_a3c.cu).  ~12 launches per env-step for both agents (the autograd path: ~180).
  * backward: only the LSTM recurrence is sequential.  One kernel turns rewards / values into returns, GAE and the gradient of
    every head output; a T-step sweep (cell backward + dh = dgates W_hh) produces dgates for all steps; everything else -- the
    input-side dgrads, EVERY weight gradient, the bias sums, the conv backward -- runs ONCE over all T x E rows, as long-K /
    tall-M tensor-core GEMMs instead of T small ones.
  * observations stay uint8 (the env's lossless encoding, values 0 / 1 / 2 / 4): a quarter of the rollout-buffer traffic.

The module's nn.Parameters remain the single source of truth (reference `state_dict` format); gradients are written into
`p.grad` (views of SharedAdam's flat buffer).  Numerics: float32 throughout, GEMMs 3xTF32 (fp32-accurate).  Validated against
the CPU restatement of the reference learner (tests/test_gpu_learner.py) and against the autograd implementation in model.py (tests/test_gpu_fused.py).
"""
import ctypes as C

import torch

from . import _lib, gemm

H = 128
NOUT = 8


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def supported(model, env, args):
    """the configuration the fused path is written for: two agents, maze encoder, LSTM(128), 13 x 13 partial observations,
    one frame per step (every 2D README command)"""
    from .model import A3C_Dueling
    shp = tuple(env.observation_space[0].shape)
    return (isinstance(model, A3C_Dueling) and not model.single and int(args.stack_frames) == 1 and int(args.rnn_out) == H
            and shp == (1, 13, 13) and len(env.observation_space) == 2 and next(model.parameters()).is_cuda)


class _Net(object):
    """one agent's parameters, activations and backward scratch"""

    def __init__(self, module, agent, tat, T, E, device):
        self.m, self.agent, self.tat = module, agent, tat
        self.frames = 2 if tat else 1
        self.F = 512 * self.frames
        f32 = dict(dtype=torch.float32, device=device)
        self.convout = torch.zeros((T + 1, E, self.F), **f32)
        self.yfc = torch.zeros((T + 1, E, 256), **f32) if tat else None
        self.xh = torch.zeros((T + 2, E, 384), **f32)       # [feature | h_prev] per step; slot T + 1 = scratch of the bootstrap step
        self.c = torch.zeros((T + 2, E, H), **f32)
        self.act = torch.zeros((T, E, 4 * H), **f32)
        self.hout = torch.zeros((T, E, H), **f32)
        self.out8 = torch.zeros((T + 1, E, NOUT), **f32)
        self.gates = torch.zeros((E, 4 * H), **f32)
        self.w_cat = torch.zeros((4 * H, 384), **f32)
        self.w_head = torch.zeros((NOUT, H), **f32)
        self.b_head = torch.zeros(NOUT, **f32)
        self.bwd = None

    def alloc_backward(self, T, E, device):
        if self.bwd is not None:
            return
        f32 = dict(dtype=torch.float32, device=device)
        self.dout8 = torch.zeros((T, E, NOUT), **f32)
        self.dgates = torch.zeros((T, E, 4 * H), **f32)
        self.dfeat = torch.zeros((T * E, 256), **f32)
        self.dconv = torch.zeros((T * E, self.F), **f32)
        self.dh = torch.zeros((E, H), **f32)
        self.dc = torch.zeros((E, H), **f32)
        self.g_head = torch.zeros((NOUT, H), **f32)
        self.g_bhead = torch.zeros(NOUT, **f32)
        self.g_blstm = torch.zeros(4 * H, **f32)
        self.bwd = True

    def heads(self):
        m = self.m
        hs = [m.actor.actor_linear, m.critic.critic_linear]
        if self.tat:
            hs.append(m.reward_aux)
        return hs

    @torch.no_grad()
    def pack(self):
        """[W_ih | W_hh] and the packed head weights -- after every weight change"""
        m = self.m
        self.w_cat[:, :256].copy_(m.lstm.weight_ih)
        self.w_cat[:, 256:].copy_(m.lstm.weight_hh)
        o = 0
        for l in self.heads():
            n = l.weight.shape[0]
            self.w_head[o:o + n].copy_(l.weight)
            self.b_head[o:o + n].copy_(l.bias)
            o += n

    def trainable(self):
        return all(p.grad is not None for p in self.m.parameters())


class FusedA3C(object):
    def __init__(self, model, num_envs, num_steps, device, seed=1):
        self.model, self.E, self.T, self.device = model, int(num_envs), int(num_steps), torch.device(device)
        self.lib = _lib.load()
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        E, T = self.E, self.T
        self.nets = [_Net(model.player0, 0, False, T, E, self.device), _Net(model.player1, 1, bool(model.tat), T, E, self.device)]
        # uint8 observation slots; the step kernel stores 16-byte vectors, so a slot (E x 338 bytes) is padded to whole groups of 8 envs
        self.Epad = (E + 7) // 8 * 8
        self.obs_store = torch.zeros((T + 1, self.Epad, 2, 1, 13, 13), dtype=torch.uint8, device=self.device)
        self.obs = self.obs_store[:, :E]
        self.actions = torch.zeros((T + 1, E, 2), dtype=torch.int32, device=self.device)
        self.values = torch.zeros((T + 1, E, 2), dtype=torch.float32, device=self.device)
        self.logp = torch.zeros((T + 1, E, 2), dtype=torch.float32, device=self.device)
        self.entropy = torch.zeros((T + 1, E, 2), dtype=torch.float32, device=self.device)
        self.logp_all = torch.zeros((2, E, 4), dtype=torch.float32, device=self.device)  # test mode: log pi(.|s) per agent
        self.stats = torch.zeros((7, E), dtype=torch.float32, device=self.device)
        self.returns = torch.zeros((T, E, 2), dtype=torch.float32, device=self.device)
        self.gae = torch.zeros((T, E, 2), dtype=torch.float32, device=self.device)
        self.eps_len = torch.zeros(E, dtype=torch.int32, device=self.device)
        self.rng_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._ws = {}
        # packed weight copies are refreshed lazily: after an update (Agent.apply_update) and after load_state_dict
        self._dirty = True
        model.register_load_state_dict_post_hook(lambda module, incompatible: self.mark_dirty())

    # ---- plumbing ------------------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def mark_dirty(self):
        """the module's weights changed: re-pack [W_ih | W_hh] and the head weights before the next forward"""
        self._dirty = True

    def pack(self):
        for n in self.nets:
            n.pack()
        self._dirty = False

    def reset_state(self):
        for n in self.nets:
            n.xh[0, :, 256:].zero_()
            n.c[0].zero_()
        self.eps_len.zero_()

    def hidden(self, slot=0):
        """(hx, cx) as the reference shapes them, (E, 2, 128) each: the recurrent state ENTERING step `slot`"""
        hx = torch.stack([n.xh[slot, :, 256:] for n in self.nets], 1)
        cx = torch.stack([n.c[slot] for n in self.nets], 1)
        return hx, cx

    def set_hidden(self, hx, cx, slot=0):
        for a, n in enumerate(self.nets):
            n.xh[slot, :, 256:].copy_(hx[:, a])
            n.c[slot].copy_(cx[:, a])

    def _workspace(self, key, n):
        ws = self._ws.get(key)
        if ws is None or ws.numel() < n:
            ws = self._ws[key] = torch.empty(max(n, 1), dtype=torch.float32, device=self.device)
        return ws

    # ---- forward -------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, t, forced=None, greedy=False, bootstrap=False, envs=None):
        """policy step `t` from self.obs[t] and the recurrent state in slot t: fills actions[t], values[t], logp[t],
        entropy[t] and the state of slot t + 1.  `bootstrap`: the value-only forward of Agent.optimize (player_util.py:110-116;
        it still samples, as the reference does) -- nothing is kept for the backward.  `envs` = (first, end): only that slice of the
        batch (every kernel is row-wise and the sampler's counter is the env index, so slices add up to exactly the whole-batch step;
        Agent uses it to start on the envs whose observations have already arrived over PCIe)."""
        with _lib.on_device(self.device):
            return self._forward(t, forced, greedy, bootstrap, envs)

    def _forward(self, t, forced, greedy, bootstrap, envs=None):
        lib, st = self.lib, self._stream()
        e0, e1 = (0, self.E) if envs is None else (int(envs[0]), int(envs[1]))
        if not 0 <= e0 < e1 <= self.E:
            raise ValueError("forward: env slice [%d, %d) outside [0, %d)" % (e0, e1, self.E))
        E = e1 - e0
        if self._dirty:
            self.pack()
        obs = self.obs[t]
        row = lambda buf, width: C.c_void_p(buf.data_ptr() + 4 * width * e0)  # noqa: E731  row e0 of a float32 / int32 [E][width] array
        for n in self.nets:
            a, m = n.agent, n.m
            enc = m.encoder
            if n.tat:
                x_ptr, stride, n_img = obs.data_ptr() + 338 * e0, 169, 2 * E
            else:
                x_ptr, stride, n_img = obs.data_ptr() + 338 * e0 + 169 * a, 338, E
            convout, xh_t = n.convout[t, e0:e1], n.xh[t, e0:e1]
            gates = n.gates[e0:e1]
            _lib.check(lib.track2d_maze_conv_forward_ex(C.c_void_p(x_ptr), 1, stride, n_img, _p(enc.conv1.weight), _p(enc.conv1.bias),
                                                        _p(enc.conv2.weight), _p(enc.conv2.bias), _p(convout), st), lib)
            feat = xh_t[:, :256]
            if n.tat:
                yfc = n.yfc[t, e0:e1]
                gemm.gemm(convout, False, n.F, enc.fc.weight, False, n.F, E, 256, n.F, bias=enc.fc.bias, relu=True, out=yfc)
                # model.py:198-199: + fc_action_tracker(one_hot(tracker action)), added after the ReLU
                _lib.check(lib.track2d_embed_add(_p(yfc), 256, _p(feat), 384, _p(m.fc_action_tracker.weight), _p(m.fc_action_tracker.bias),
                                                 row(self.actions[t], 2), 256, E, st), lib)
            else:
                gemm.gemm(convout, False, n.F, enc.fc.weight, False, n.F, E, 256, n.F, bias=enc.fc.bias, relu=True, out=feat)
            gemm.gemm(xh_t, False, 384, n.w_cat, False, 384, E, 4 * H, 384, out=gates)
            keep = not bootstrap
            col = lambda buf: C.c_void_p(buf[t].data_ptr() + 4 * (2 * e0 + a))  # noqa: E731  column `a` of an [E][2] array, from row e0
            _lib.check(lib.track2d_lstm_heads_forward(
                _p(gates), _p(m.lstm.bias_ih), _p(m.lstm.bias_hh), row(n.c[t], H), row(n.act[t], 4 * H) if keep else None, row(n.c[t + 1], H),
                row(n.hout[t], H) if keep else None, C.c_void_p(n.xh[t + 1].data_ptr() + 4 * (384 * e0 + 256)), 384, _p(n.w_head), _p(n.b_head),
                row(n.out8[t], NOUT), col(self.actions), C.c_void_p(forced.data_ptr() + 4 * (2 * e0 + a)) if forced is not None else None,
                col(self.values), col(self.logp), col(self.entropy), row(self.logp_all[a], 4) if greedy else None, _p(self.rng_step), self.seed,
                a | (2 if bootstrap else 0), int(greedy), E, e0, st), lib)
        return self.actions[t]

    @torch.no_grad()
    def post_step(self, t, done):
        """after env.step of step t: zero the recurrent state of finished envs in slot t + 1, eps_len, sampling counter"""
        n0, n1 = self.nets
        with _lib.on_device(self.device):
            _lib.check(self.lib.track2d_policy_post_step(_p(done), C.c_void_p(n0.xh[t + 1].data_ptr() + 4 * 256), C.c_void_p(n1.xh[t + 1].data_ptr() + 4 * 256),
                                                         384, _p(n0.c[t + 1]), _p(n1.c[t + 1]), _p(self.eps_len), _p(self.rng_step), self.E, self._stream()), self.lib)

    @torch.no_grad()
    def carry_over(self, T, obs=True):
        """the next rollout starts from the state after step T - 1 (train.py:76 update_rnn_hiden: same values, no graph)"""
        if obs:
            self.obs[0].copy_(self.obs[T])
        for n in self.nets:
            n.xh[0, :, 256:].copy_(n.xh[T, :, 256:])
            n.c[0].copy_(n.c[T])

    # ---- backward ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def backward(self, T, rewards, done, training_mode, w_ent, use_aux, gamma, tau, scale):
        """Fills p.grad of the trained agents' parameters with the gradient of
        scale * sum_e [loss_tracker + loss_target (+ pred_loss)] (player_util.py:141-155) and returns the per-env statistics
        (policy_loss (E,2), value_loss (E,2), entropy sums (E,2), pred_loss (E,)).  Needs forward(T, bootstrap=True) first."""
        with _lib.on_device(self.device):
            return self._backward(T, rewards, done, training_mode, w_ent, use_aux, gamma, tau, scale)

    def _backward(self, T, rewards, done, training_mode, w_ent, use_aux, gamma, tau, scale):
        lib, E, st = self.lib, self.E, self._stream()
        M2 = T * E
        train = [training_mode in (-1, 0) and self.nets[0].trainable(), training_mode in (-1, 1) and self.nets[1].trainable()]
        aux = bool(use_aux) and self.nets[1].tat and training_mode != 0 and self.nets[1].trainable()
        for n in self.nets:
            n.alloc_backward(self.T, E, self.device)
        n0, n1 = self.nets
        _lib.check(lib.track2d_a3c_loss_grad(_p(n0.out8), _p(n1.out8), _p(n0.dout8), _p(n1.dout8), _p(self.actions), _p(rewards), _p(done), _p(self.stats),
                                             _p(self.returns), _p(self.gae), T, E, float(gamma), float(tau), float(w_ent[0]), float(w_ent[1]), float(scale),
                                             int(train[0]), int(train[1]), 2 if aux else int(bool(use_aux) and self.nets[1].tat), st), lib)
        for n in self.nets:
            if not (train[n.agent] or (n.agent == 1 and aux)):
                continue
            m, enc = n.m, n.m.encoder
            dout8 = n.dout8[:T].reshape(M2, NOUT)
            dgates = n.dgates[:T].reshape(M2, 4 * H)
            # heads: dW = dout8^T h, db = column sums
            gemm.gemm(dout8, True, NOUT, n.hout[:T].reshape(M2, H), True, H, NOUT, H, M2, out=n.g_head)
            n.g_bhead.copy_(gemm.colsum(dout8))
            o = 0
            for l in n.heads():
                k = l.weight.shape[0]
                l.weight.grad.copy_(n.g_head[o:o + k])
                l.bias.grad.copy_(n.g_bhead[o:o + k])
                o += k
            # the BPTT sweep: cell backward, then the recurrent gradient of the previous step
            for t in range(T - 1, -1, -1):
                last = t == T - 1
                _lib.check(lib.track2d_lstm_heads_backward(_p(n.dout8[t]), _p(n.w_head), None if last else _p(n.dh), _p(n.dc), _p(done[t]),
                                                           _p(n.act[t]), _p(n.c[t]), _p(n.dgates[t]), E, st), lib)
                if t > 0:
                    gemm.gemm(n.dgates[t], False, 4 * H, m.lstm.weight_hh, True, H, E, H, 4 * H, out=n.dh)
            # everything below is batched over all T x E rows
            n.g_blstm.copy_(gemm.colsum(dgates))
            m.lstm.bias_ih.grad.copy_(n.g_blstm)
            m.lstm.bias_hh.grad.copy_(n.g_blstm)
            xh = n.xh[:T].reshape(M2, 384)
            gemm.gemm(dgates, True, 4 * H, xh, True, 384, 4 * H, 256, M2, out=m.lstm.weight_ih.grad)
            gemm.gemm(dgates, True, 4 * H, xh[:, 256:], True, 384, 4 * H, H, M2, out=m.lstm.weight_hh.grad)
            gemm.gemm(dgates, False, 4 * H, m.lstm.weight_ih, True, 256, M2, 256, 4 * H, out=n.dfeat)
            # ReLU backward of the fc (+ the embedding's gradients, TAT), bias gradient in the same pass
            n_ws = int(lib.track2d_relu_backward_workspace_floats(M2, 256))
            ws = self._workspace("relu", n_ws)
            if n.tat:
                y, y_ld, grp = n.yfc, 256, _p(self.actions)
                gw, gb = _p(m.fc_action_tracker.weight.grad), _p(m.fc_action_tracker.bias.grad)
            else:
                y, y_ld, grp, gw, gb = n.xh, 384, None, None, None
            _lib.check(lib.track2d_relu_backward_groupsum(_p(n.dfeat), _p(y), y_ld, grp, M2, 256, _p(enc.fc.bias.grad), gw, gb, _p(ws), n_ws, st), lib)
            conv = n.convout[:T].reshape(M2, n.F)
            gemm.gemm(n.dfeat, True, 256, conv, True, n.F, 256, n.F, M2, out=enc.fc.weight.grad)
            gemm.gemm(n.dfeat, False, 256, enc.fc.weight, True, n.F, M2, n.F, 256, out=n.dconv)
            # conv stack backward over every image of the rollout (one launch when the observation slots are contiguous)
            chunks = [(0, T)] if self.Epad == E else [(t, 1) for t in range(T)]
            for t0, nt in chunks:
                rows = nt * E
                if n.tat:
                    x_ptr, stride, n_img = self.obs[t0].data_ptr(), 169, 2 * rows
                else:
                    x_ptr, stride, n_img = self.obs[t0].data_ptr() + 169 * n.agent, 338, rows
                _lib.check(lib.track2d_maze_conv_backward_ex(C.c_void_p(x_ptr), 1, stride, _p(n.convout[t0]), _p(n.dconv[t0 * E:]), n_img, _p(enc.conv1.weight),
                                                             _p(enc.conv1.bias), _p(enc.conv2.weight), _p(enc.conv1.weight.grad), _p(enc.conv1.bias.grad),
                                                             _p(enc.conv2.weight.grad), _p(enc.conv2.bias.grad), st), lib)
        s = self.stats
        return (torch.stack([s[0], s[1]], 1), torch.stack([s[2], s[3]], 1), torch.stack([s[4], s[5]], 1), s[6])
