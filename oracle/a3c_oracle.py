"""CPU ORACLE for the learner half of the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, float32) restatement of the reference's per-env, batch-of-one A3C arithmetic:

    forward           model.py:133-145 (A3C), :190-209 (TAT), :238-264 (A3C_Dueling); perception.py:86-92
    optimize          player_util.py:108-161 (n-step return, GAE, policy/value/entropy/aux-L1 losses)
    shared_adam_step  shared_optim.py:122-175 (+ clip_grad_norm_, player_util.py:157)
    worker            train.py:69-109: one reference worker (rollout <= num_steps, optimize), used by
                      bench.py's cpu_baseline / --impl reference legs

Weights are a reference-format state_dict (player0.encoder.conv1.weight, ...), so reference checkpoints
plug in.  Parity status: PINNED by tests/test_learner_oracle.py against tests/golden/learner_*.npz,
recorded from the unmodified reference (oracle/refharness/make_golden_learner.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def det_state_dict(tat=True, seed=1234, scale=0.08):
    """A deterministic reference-format state_dict (numpy legacy RNG, stable everywhere) so that
    fixtures need not store 800k floats."""
    shapes = []
    for p, is_tat in (("player0", False), ("player1", tat)):
        fc_in = 1024 if is_tat else 512
        shapes += [(p + ".encoder.conv1.weight", (16, 1, 3, 3)), (p + ".encoder.conv1.bias", (16,)),
                   (p + ".encoder.conv2.weight", (32, 16, 3, 3)), (p + ".encoder.conv2.bias", (32,)),
                   (p + ".encoder.fc.weight", (256, fc_in)), (p + ".encoder.fc.bias", (256,)),
                   (p + ".lstm.weight_ih", (512, 256)), (p + ".lstm.weight_hh", (512, 128)),
                   (p + ".lstm.bias_ih", (512,)), (p + ".lstm.bias_hh", (512,)),
                   (p + ".actor.actor_linear.weight", (4, 128)), (p + ".actor.actor_linear.bias", (4,)),
                   (p + ".critic.critic_linear.weight", (1, 128)), (p + ".critic.critic_linear.bias", (1,))]
        if is_tat:
            shapes += [(p + ".fc_action_tracker.weight", (256, 4)), (p + ".fc_action_tracker.bias", (256,)),
                       (p + ".reward_aux.weight", (1, 128)), (p + ".reward_aux.bias", (1,))]
    rs = np.random.RandomState(seed)
    sd = {}
    for name, shape in shapes:
        sd[name] = torch.from_numpy(rs.uniform(-scale, scale, size=shape).astype(np.float32))
    return sd


def _encoder(sd, p, x):
    """CNN_maze.forward (perception.py:86-92): x (frames, 1, 13, 13) -> (1, 256)"""
    x = F.relu(F.conv2d(x, sd[p + ".encoder.conv1.weight"], sd[p + ".encoder.conv1.bias"], stride=2, padding=1))
    x = F.relu(F.conv2d(x, sd[p + ".encoder.conv2.weight"], sd[p + ".encoder.conv2.bias"], stride=2, padding=1))
    x = x.view(1, -1)
    return F.relu(F.linear(x, sd[p + ".encoder.fc.weight"], sd[p + ".encoder.fc.bias"]))


def _lstm(sd, p, x, hx, cx):
    """nn.LSTMCell: gates i, f, g, o"""
    gates = F.linear(x, sd[p + ".lstm.weight_ih"], sd[p + ".lstm.bias_ih"]) + F.linear(hx, sd[p + ".lstm.weight_hh"], sd[p + ".lstm.bias_hh"])
    i, f, g, o = gates.chunk(4, 1)
    cx = torch.sigmoid(f) * cx + torch.sigmoid(i) * torch.tanh(g)
    hx = torch.sigmoid(o) * torch.tanh(cx)
    return hx, cx


def _heads(sd, p, feature, test, forced):
    value = F.linear(feature, sd[p + ".critic.critic_linear.weight"], sd[p + ".critic.critic_linear.bias"])
    logit = F.linear(feature, sd[p + ".actor.actor_linear.weight"], sd[p + ".actor.actor_linear.bias"])
    prob = F.softmax(logit, dim=1)
    log_prob = F.log_softmax(logit, dim=1)
    entropy = -(log_prob * prob).sum(1, keepdim=True)
    if test:
        action = prob.max(1)[1]
    else:
        action = torch.tensor([[int(forced)]]) if forced is not None else prob.multinomial(1)
        log_prob = log_prob.gather(1, action)
        action = action.view(-1)
    return value, int(action.item()), entropy, log_prob


def forward(sd, state, hx, cx, tat=True, test=False, forced=None):
    """A3C_Dueling.forward for ONE env.  state (2, 1, 1, 13, 13); hx, cx (2, 128).
    Returns values (2,1), [a0, a1], entropies (2,1), log_probs (2,1)|(2,4), (hx, cx), R_pred (1,1)|0"""
    f0 = _encoder(sd, "player0", state[0])
    h0, c0 = _lstm(sd, "player0", f0, hx[:1], cx[:1])
    v0, a0, e0, lp0 = _heads(sd, "player0", h0, test, None if forced is None else forced[0])
    R_pred = 0
    if tat:
        onehot = torch.zeros(4)
        onehot[a0] = 1
        st = torch.cat((state[0], state[1]), 0)  # model.py:253
        f1 = _encoder(sd, "player1", st) + F.linear(onehot, sd["player1.fc_action_tracker.weight"], sd["player1.fc_action_tracker.bias"])
    else:
        f1 = _encoder(sd, "player1", state[1])
    h1, c1 = _lstm(sd, "player1", f1, hx[1:], cx[1:])
    v1, a1, e1, lp1 = _heads(sd, "player1", h1, test, None if forced is None else forced[1])
    if tat:
        R_pred = F.linear(h1, sd["player1.reward_aux.weight"], sd["player1.reward_aux.bias"])
    return torch.cat([v0, v1]), [a0, a1], torch.cat([e0, e1]), torch.cat([lp0, lp1]), (torch.cat((h0, h1)), torch.cat((c0, c1))), R_pred


def segment_loss(values, log_probs, entropies, preds, rewards, R_boot, gamma=0.9, tau=1.0, w_entropy=0.01, w_entropy_target=0.2,
                 aux=True, training_mode=-1):
    """player_util.py:117-154 for one rollout (segment): lists over time of (2,1) tensors; rewards list of
    (2,1); R_boot (2,1) = 0 if the rollout ended the episode else V(s_T).data.  Returns (loss, policy_loss (2,1),
    value_loss (2,1), pred_loss (1,1))."""
    values = list(values) + [R_boot]
    policy_loss = torch.zeros(2, 1)
    value_loss = torch.zeros(2, 1)
    pred_loss = torch.zeros(1, 1)
    w = torch.tensor([[float(w_entropy)], [float(w_entropy_target)]])
    R = R_boot.clone()
    gae = torch.zeros(1, 1)
    for i in reversed(range(len(rewards))):
        if aux:
            pred_loss = pred_loss + (preds[i][0] - rewards[i][0]).abs().mean()
        R = gamma * R + rewards[i]
        advantage = R - values[i]
        value_loss = value_loss + 0.5 * advantage.pow(2)
        delta_t = rewards[i] + gamma * values[i + 1].detach() - values[i].detach()
        gae = gae * gamma * tau + delta_t
        policy_loss = policy_loss - (log_probs[i] * gae.detach()) - (w * entropies[i])
    loss_tracker = (policy_loss[0] + 0.5 * value_loss[0]).mean()
    loss_target = (policy_loss[1] + 0.5 * value_loss[1]).mean()
    loss = loss_tracker if training_mode == 0 else (loss_target if training_mode == 1 else loss_tracker + loss_target)
    if aux and training_mode != 0:
        loss = loss + pred_loss.mean()
    return loss, policy_loss, value_loss, pred_loss


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ over a list of tensors (in place); returns the total norm"""
    total = torch.sqrt(sum((g.detach() ** 2).sum() for g in grads))
    coef = min(1.0, float(max_norm) / (float(total) + 1e-6))
    for g in grads:
        g.mul_(coef)
    return float(total)


def shared_adam_step(params, grads, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-3):
    """SharedAdam.step with amsgrad=True (shared_optim.py:122-175).  state: dict name -> (step, m, v, vmax)"""
    b1, b2 = betas
    for name in params:
        g = grads[name]
        if name not in state:
            state[name] = [0, torch.zeros_like(params[name]), torch.zeros_like(params[name]), torch.zeros_like(params[name])]
        st = state[name]
        st[0] += 1
        st[1].mul_(b1).add_(g, alpha=1 - b1)
        st[2].mul_(b2).addcmul_(g, g, value=1 - b2)
        torch.max(st[3], st[2], out=st[3])
        denom = st[3].sqrt().add_(eps)
        step_size = lr * math.sqrt(1 - b2 ** st[0]) / (1 - b1 ** st[0])
        params[name].addcdiv_(st[1], denom, value=-step_size)


class Worker(object):
    """One reference A3C worker (train.py:69-109) on an oracle env: rollout of <= num_steps steps (stops at done),
    optimize, SharedAdam step on the shared weights.  Used as the CPU baseline of the whole path."""

    def __init__(self, env, sd, adam_state, tat=True, num_steps=20, gamma=0.9, tau=1.0, entropy=0.01, entropy_target=0.2, lr=1e-3,
                 training_mode=-1, seed=1):
        self.env, self.sd, self.adam_state, self.tat = env, sd, adam_state, tat
        self.num_steps, self.gamma, self.tau, self.entropy, self.entropy_target, self.lr = num_steps, gamma, tau, entropy, entropy_target, lr
        self.training_mode = training_mode
        torch.manual_seed(seed)
        self.done = True
        self.n_steps = 0
        self.first_update = True

    def _reset(self):
        self.state = torch.from_numpy(np.float32(self.env.reset())).unsqueeze(1)  # frame_stack: (2, 1, 1, 13, 13)
        self.hx, self.cx = torch.zeros(2, 128), torch.zeros(2, 128)

    def iteration(self):
        local = {k: v.clone().requires_grad_(True) for k, v in self.sd.items()}  # load_state_dict(shared_model.state_dict())
        if self.done:
            self._reset()
        self.hx, self.cx = self.hx.detach(), self.cx.detach()
        values, log_probs, entropies, preds, rewards = [], [], [], [], []
        for _ in range(self.num_steps):
            v, acts, ent, lp, (self.hx, self.cx), rp = forward(local, self.state, self.hx, self.cx, self.tat)
            obs, rew, self.done, _ = self.env.step(acts)
            self.state = torch.from_numpy(np.float32(obs)).unsqueeze(1)
            self.n_steps += 1
            values.append(v); log_probs.append(lp); entropies.append(ent); preds.append(rp)
            rewards.append(torch.tensor(rew).float().unsqueeze(1))
            if self.done:
                break
        R = torch.zeros(2, 1)
        if not self.done:
            with torch.no_grad():
                R = forward(local, self.state, self.hx, self.cx, self.tat)[0]
        loss, _, _, _ = segment_loss(values, log_probs, entropies, preds, rewards, R, self.gamma, self.tau, self.entropy,
                                     self.entropy_target, self.tat, self.training_mode)
        names = list(local)
        grads = torch.autograd.grad(loss, [local[n] for n in names], allow_unused=True)
        gd = {n: (g if g is not None else torch.zeros_like(local[n])) for n, g in zip(names, grads)}
        # No clipping: the reference's clip_grad_norm_(params, 50) (player_util.py:157) is inert -- `params` is
        # the generator train.py:39-44 creates once; the first call consumes it while the shared grads are still
        # None, every later call sees an exhausted generator (pinned by tests/golden/learner_*.npz).
        shared_adam_step(self.sd, gd, self.adam_state, lr=self.lr)
        return float(loss.detach())
