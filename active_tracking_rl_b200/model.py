"""Batched maze-lstm / tat-maze-lstm policies: the reference's A3C_Dueling (model.py:212-264) with its
hard-wired batch of 1 (perception.py:89 `x.view(1, -1)`, model.py:242 `hx[:1]`) lifted to a leading env
axis.  Per env the arithmetic is the reference's; parameter names and shapes are the reference's, so
`state_dict()` files interchange both ways (test.py:112-127 checkpoints):

    player{0,1}.encoder.{conv1,conv2,fc}.{weight,bias}     CNN_maze          perception.py:68-92
    player{0,1}.lstm.{weight_ih,weight_hh,bias_ih,bias_hh} LSTMCell(256,128) model.py:116-118
    player{0,1}.actor.actor_linear / critic.critic_linear  heads             model.py:55-99
    player1.fc_action_tracker, player1.reward_aux          TAT only          model.py:175-182

Layouts: observations (E, 2, 1, 13, 13) float32 (agent axis second); recurrent state hx, cx (E, 2, 128).
Only the discrete, maze-encoder, LSTM configuration the 2D README commands use is built.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Conv3x3S2(torch.autograd.Function):
    """3x3 / stride 2 / pad 1 convolution (perception.py:71-72) with a GEMM-shaped backward.

    cuDNN's fp32 backward kernels for these shapes (1->16 and 16->32 channels on 13x13 / 7x7 images, a few hundred
    thousand images per call) run at a few percent of the machine (dgrad2d_grouped_direct 14.8 ms, wgrad_alg1 5 ms
    per call at 196,608 images; profiles/).  Here the backward recomputes the im2col matrix directly in
    [C*9, N*L] layout (one gather), and both gradients are plain SGEMMs with a long reduction axis:
        dW [OC, C*9]   = dY [OC, N*L] @ cols^T
        dcols [C*9, N*L] = W^T @ dY, scattered back with 9 strided adds (col2im)
    Nothing but x and w is kept between forward and backward."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return F.conv2d(x, w, b, stride=2, padding=1)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        N, C, H, W = x.shape
        OC, OH, OW = w.shape[0], gy.shape[2], gy.shape[3]
        K, NL = C * 9, N * OH * OW
        g = gy.permute(1, 0, 2, 3).reshape(OC, NL)  # [OC, N*L]
        gb = g.sum(1)
        xp = F.pad(x, (1, 1, 1, 1))
        win = xp.unfold(2, 3, 2).unfold(3, 3, 2)  # [N, C, OH, OW, 3, 3] view
        cols = win.permute(1, 4, 5, 0, 2, 3).reshape(K, NL)  # one gather, [C*9, N*L]
        gw = torch.mm(g, cols.t()).view_as(w)
        gx = None
        if ctx.needs_input_grad[0]:
            dcols = torch.mm(w.view(OC, K).t(), g).view(C, 3, 3, N, OH, OW)
            gxp = x.new_zeros((C, N, H + 2, W + 2))
            for ki in range(3):
                for kj in range(3):
                    gxp[:, :, ki:ki + 2 * OH:2, kj:kj + 2 * OW:2] += dcols[:, ki, kj]
            gx = gxp[:, :, 1:H + 1, 1:W + 1].permute(1, 0, 2, 3)
        return gx, gw, gb


def conv3x3s2(x, conv):
    return _Conv3x3S2.apply(x, conv.weight, conv.bias)


class _MazeConvStack(torch.autograd.Function):
    """conv1 + ReLU + conv2 + ReLU of CNN_maze as ONE hand-written CUDA kernel each way (csrc/track2d_policy.cu,
    track2d_maze_conv_forward / _backward).  x (N, 1, 13, 13) -> (N, 512)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        x = x.contiguous()
        N = x.shape[0]
        y2 = torch.empty((N, 512), dtype=torch.float32, device=x.device)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        with _lib.on_device(x.device):
            _lib.check(lib.track2d_maze_conv_forward(p(x), N, p(w1), p(b1), p(w2), p(b2), p(y2),
                                                     C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), lib)
        ctx.save_for_backward(x, y2, w1, b1, w2)
        return y2

    @staticmethod
    def backward(ctx, gy2):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        x, y2, w1, b1, w2 = ctx.saved_tensors
        gy2 = gy2.contiguous()
        dw1, db1, dw2 = torch.zeros_like(w1), torch.zeros_like(b1), torch.zeros_like(w2)
        db2 = torch.zeros(32, dtype=torch.float32, device=x.device)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        with _lib.on_device(x.device):
            _lib.check(lib.track2d_maze_conv_backward(p(x), p(y2), p(gy2), x.shape[0], p(w1), p(b1), p(w2), p(dw1), p(db1), p(dw2), p(db2),
                                                      C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), lib)
        return None, dw1, db1, dw2, db2


# "fused": the hand-written conv-stack kernels (CUDA tensors); "gemm_backward": cuDNN forward + im2col/SGEMM backward;
# "cudnn": plain F.conv2d autograd.  CPU tensors (tests of the host-side logic) always take the F.conv2d route.
CONV_IMPL = "fused"
# "tf32x3": the fc / LSTM GEMMs (forward, dgrad, wgrad) run on the tcgen05 tensor cores with fp32 accuracy
# (csrc/track2d_gemm.cu, 3xTF32 split); "cublas": torch's own fp32 linear (SIMT SGEMM).  CPU tensors always take torch's.
GEMM_IMPL = "tf32x3"


def _linear(x, lin, relu=False):
    if GEMM_IMPL == "tf32x3" and x.is_cuda:
        from . import gemm
        return gemm.linear(x, lin.weight, lin.bias, relu)
    y = lin(x)
    return F.relu(y) if relu else y


def _lstm_cell(lstm, x, hx, cx):
    """nn.LSTMCell.forward (model.py:116 `self.lstm(feature, (hx, cx))`): two gate GEMMs + the fused pointwise cell"""
    if GEMM_IMPL == "tf32x3" and x.is_cuda:
        from . import gemm
        igates = gemm.linear(x, lstm.weight_ih)
        hgates = gemm.linear(hx, lstm.weight_hh)
        if gemm.lstm_pointwise_supported(igates, cx):
            return gemm.lstm_pointwise(igates, hgates, cx, lstm.bias_ih, lstm.bias_hh)
        hy, cy, _ = torch.ops.aten._thnn_fused_lstm_cell(igates, hgates, cx, lstm.bias_ih, lstm.bias_hh)
        return hy, cy
    return lstm(x, (hx, cx))


def weights_init(m):
    """utils.py:47-62: U(+-sqrt(6 / (fan_in + fan_out))) for every Conv / Linear, zero bias.  It is the
    LAST init applied (model.py:130,187 `self.apply(weights_init)`), so it overrides norm_col_init etc."""
    if isinstance(m, nn.Conv2d):
        shape = list(m.weight.shape)
        fan_in = shape[1] * shape[2] * shape[3]
        fan_out = shape[2] * shape[3] * shape[0]
    elif isinstance(m, nn.Linear):
        fan_out, fan_in = m.weight.shape
    else:
        return
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    with torch.no_grad():
        m.weight.uniform_(-bound, bound)
        m.bias.zero_()


def _heads(hx, *linears):
    """the 4- / 1-wide output heads of one agent (actor, critic[, reward_aux]) on the recurrent state.  On the GPU they are ONE
    GEMM against the concatenated weights, zero-padded to 8 rows (cuBLAS ran each head's dW = dy^T h over 65,536 rows as a
    107 us SIMT kernel); the parameters stay separate tensors, so state_dict files are unchanged."""
    if GEMM_IMPL == "tf32x3" and hx.is_cuda:
        from . import gemm
        n = sum(l.weight.shape[0] for l in linears)
        pad = (-n) % 8
        ws, bs = [l.weight for l in linears], [l.bias for l in linears]
        if pad:
            ws.append(hx.new_zeros((pad, hx.shape[1])))
            bs.append(hx.new_zeros(pad))
        out = gemm.linear(hx, torch.cat(ws, 0), torch.cat(bs, 0))
        res, o = [], 0
        for l in linears:
            res.append(out[:, o:o + l.weight.shape[0]])
            o += l.weight.shape[0]
        return res
    return [l(hx) for l in linears]


class CNN_maze(nn.Module):
    """perception.py:68-92.  `frames` observations per env are run through the convs as separate images
    and their features concatenated before fc (the reference gets that from view(1, -1) over the
    stack axis): fc sees frames * 32 * 4 * 4 inputs."""

    def __init__(self, obs_shape, frames):
        super().__init__()
        c, h, w = obs_shape
        self.frames = frames
        self.conv1 = nn.Conv2d(c, 16, 3, stride=2, padding=1)
        self.conv2 = nn.Conv2d(16, 32, 3, stride=2, padding=1)
        h2 = ((h + 2 - 3) // 2 + 1 + 2 - 3) // 2 + 1
        w2 = ((w + 2 - 3) // 2 + 1 + 2 - 3) // 2 + 1
        self.fc = nn.Linear(frames * 32 * h2 * w2, 256)
        self.outdim = 256

    def forward(self, x):
        """x: (B, frames, C, H, W) -> (B, 256)"""
        B = x.shape[0]
        y = x.reshape((B * self.frames,) + tuple(x.shape[2:]))
        if CONV_IMPL == "fused" and y.is_cuda and y.shape[1:] == (1, 13, 13) and not y.requires_grad:
            y = _MazeConvStack.apply(y, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias)
        elif CONV_IMPL == "gemm_backward":
            y = F.relu(conv3x3s2(y, self.conv1))
            y = F.relu(conv3x3s2(y, self.conv2))
        else:
            y = F.relu(self.conv1(y))
            y = F.relu(self.conv2(y))
        return _linear(y.reshape(B, -1), self.fc, relu=True)


class PolicyNet(nn.Module):
    def __init__(self, input_dim, n_actions):
        super().__init__()
        self.actor_linear = nn.Linear(input_dim, n_actions)

    def forward(self, x):
        return self.actor_linear(x)


class ValueNet(nn.Module):
    def __init__(self, input_dim):
        super().__init__()
        self.critic_linear = nn.Linear(input_dim, 1)

    def forward(self, x):
        return self.critic_linear(x)


def sample_action(logit, test=False, forced=None):
    """model.py:41-50 for a batch: returns (action int64 (B,), entropy (B,1), log_prob (B,1) [train] or
    (B, n) [test]).  `forced` replays given actions instead of sampling (parity tests)."""
    prob = F.softmax(logit, dim=1)
    log_prob = F.log_softmax(logit, dim=1)
    entropy = -(log_prob * prob).sum(1, keepdim=True)
    if test:
        action = prob.max(1)[1]
    else:
        action = forced.reshape(-1, 1) if forced is not None else prob.multinomial(1)
        log_prob = log_prob.gather(1, action)
        action = action.squeeze(1)
    return action.detach(), entropy, log_prob


class A3C(nn.Module):
    """model.py:102-145 (tracker; also the naive-dueling target)."""

    def __init__(self, obs_shape, n_actions, rnn_out=128, frames=1):
        super().__init__()
        self.encoder = CNN_maze(obs_shape, frames)
        self.lstm = nn.LSTMCell(self.encoder.outdim, rnn_out)
        self.actor = PolicyNet(rnn_out, n_actions)
        self.critic = ValueNet(rnn_out)
        self.apply(weights_init)
        with torch.no_grad():
            self.lstm.bias_ih.zero_()
            self.lstm.bias_hh.zero_()

    def forward(self, x, hx, cx, test=False, forced=None):
        feature = self.encoder(x)
        hx, cx = _lstm_cell(self.lstm, feature, hx, cx)
        logit, value = _heads(hx, self.actor.actor_linear, self.critic.critic_linear)
        action, entropy, log_prob = sample_action(logit, test, forced)
        return value, action, entropy, log_prob, hx, cx


class TAT(nn.Module):
    """model.py:148-209 tracker-aware target: sees (tracker obs, own obs) and the tracker's action
    one-hot; auxiliary head predicts the tracker's reward."""

    def __init__(self, obs_shape, n_actions, rnn_out=128, frames=2, dim_action_tracker=4):
        super().__init__()
        self.encoder = CNN_maze(obs_shape, frames)
        self.lstm = nn.LSTMCell(self.encoder.outdim, rnn_out)
        self.actor = PolicyNet(rnn_out, n_actions)
        self.critic = ValueNet(rnn_out)
        self.fc_action_tracker = nn.Linear(dim_action_tracker, self.encoder.outdim)
        self.reward_aux = nn.Linear(rnn_out, 1)
        self.apply(weights_init)
        with torch.no_grad():
            self.lstm.bias_ih.zero_()
            self.lstm.bias_hh.zero_()

    def forward(self, x, hx, cx, action_tracker_onehot, test=False, forced=None):
        feature = self.encoder(x) + _linear(action_tracker_onehot, self.fc_action_tracker)
        hx, cx = _lstm_cell(self.lstm, feature, hx, cx)
        logit, value, r_pred = _heads(hx, self.actor.actor_linear, self.critic.critic_linear, self.reward_aux)
        action, entropy, log_prob = sample_action(logit, test, forced)
        return value, action, entropy, log_prob, hx, cx, r_pred


class A3C_Dueling(nn.Module):
    """model.py:212-264.  forward(obs (E,2,1,13,13), hx (E,2,128), cx (E,2,128)) ->
    values (E,2), actions int64 (E,2), entropies (E,2), log_probs (E,2) [train], (hx, cx), R_pred (E,1) | None"""

    def __init__(self, obs_space, action_space, args, device=None):
        super().__init__()
        self.num_agents = len(obs_space)
        head_name = args.network
        if 'maze' not in head_name or 'lstm' not in head_name or 'continuous' in head_name:
            raise NotImplementedError("only the maze-lstm / tat-maze-lstm configurations of the 2D path are built (got %r)" % head_name)
        frames = int(args.stack_frames)
        self.single = bool(getattr(args, 'single', False))
        self.action_dim_tracker = action_space[0].n
        self.player0 = A3C(obs_space[0].shape, action_space[0].n, args.rnn_out, frames)
        self.tat = 'tat' in head_name
        if not self.single:
            if self.tat:
                self.player1 = TAT(obs_space[1].shape, action_space[1].n, args.rnn_out, frames * 2, self.action_dim_tracker)
            else:
                self.player1 = A3C(obs_space[1].shape, action_space[1].n, args.rnn_out, frames)

    def forward(self, inputs, test=False, forced_actions=None):
        obs, (hx, cx) = inputs
        f0 = forced_actions[:, 0] if forced_actions is not None else None
        f1 = forced_actions[:, 1] if forced_actions is not None else None
        v0, a0, e0, lp0, h0, c0 = self.player0(obs[:, 0:1], hx[:, 0], cx[:, 0], test, f0)
        if self.single:
            return v0, a0.unsqueeze(1), e0, lp0, (h0.unsqueeze(1), c0.unsqueeze(1)), None
        R_pred = None
        if self.tat:
            onehot = F.one_hot(a0, self.action_dim_tracker).to(obs.dtype)  # model.py:251-252
            v1, a1, e1, lp1, h1, c1, R_pred = self.player1(obs, hx[:, 1], cx[:, 1], onehot, test, f1)  # cat(states[0], states[1]) :253
        else:
            v1, a1, e1, lp1, h1, c1 = self.player1(obs[:, 1:2], hx[:, 1], cx[:, 1], test, f1)
        values = torch.cat([v0, v1], 1)
        actions = torch.stack([a0, a1], 1)
        entropies = torch.cat([e0, e1], 1)
        if test:
            log_probs = torch.stack([lp0, lp1], 1)  # (E, 2, n_actions), as the reference returns in test mode
        else:
            log_probs = torch.cat([lp0, lp1], 1)
        return values, actions, entropies, log_probs, (torch.stack([h0, h1], 1), torch.stack([c0, c1], 1)), R_pred


def build_model(obs_space, action_space, args, device):
    """model.py:12-15"""
    model = A3C_Dueling(obs_space, action_space, args, device)
    model.train()
    return model
