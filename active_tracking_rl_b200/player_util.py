"""Agent (player_util.py:9-161) for E envs at once: rollout buffers, action_train / action_test, reset,
update_rnn_hiden and optimize, with the reference's per-env arithmetic.

Differences forced by batching (hard part 3):
  * a reference rollout stops at `done` and the worker resets before the next one (train.py:73-88);
    here every rollout has exactly num_steps steps, finished envs are reset inside env.step, and the
    return / GAE recursions, the bootstrap value and the LSTM state are cut at those boundaries -- so
    for each env the loss is the sum of the reference's losses over the episode segments in the window;
  * the 16 Hogwild workers' separate updates become ONE update with the mean of the per-env gradients.
"""
import ctypes as C

import torch

from . import _lib, learner


def host_pipeline_plan(num_envs, obs_bytes, chunks=None, forward_slices=None):
    """(number of transfer chunks, set of chunk counts after which a forward slice starts) of Agent._step_through_host.
    Defaults by batch size: ~2 MB per transfer chunk, at most 16 (a small batch is bound by launches, not by the bus: one chunk); two
    forward slices once half a batch still fills the GPU (>= 96 GEMM super-tiles of 256 rows), else one.  forward_slices: a count of
    equal slices, or the chunk counts at which a slice ends, e.g. (8, 12, 16)."""
    C_ = int(chunks) if chunks is not None else max(1, min(16, int(obs_bytes) // (2 << 20)))
    if not 1 <= C_ <= 16:
        raise ValueError("host pipeline: 1..16 transfer chunks")
    S_ = forward_slices if forward_slices is not None else (2 if num_envs >= 2 * 96 * 256 else 1)
    if isinstance(S_, int):
        n = max(1, min(S_, C_))
        ends = {(k + 1) * C_ // n for k in range(n)}
    else:
        ends = {int(k) for k in S_}
        if not ends or min(ends) < 1 or max(ends) > C_:
            raise ValueError("host pipeline: slice boundaries are chunk counts in 1..%d" % C_)
    ends.add(C_)
    return C_, ends


class Agent(object):
    def __init__(self, model, env, args, state, device):
        self.model = model
        self.env = env
        self.args = args
        self.device = torch.device(device)
        self.num_agents = len(env.observation_space)
        if self.num_agents != 2 or bool(getattr(args, 'single', False)):
            # the rollout buffers, the GAE / loss kernels and optimize() are written for the tracker-target pair of the 2D path
            raise NotImplementedError("single-agent mode (--single) is not built: the 2D path always has a tracker and a target "
                                      "(scripted targets Ram / Nav / RPF ignore the second policy's action)")
        self.num_envs = env.num_envs
        self.dim_action = 1
        self.rnn_out = args.rnn_out
        self.w_entropy_target = getattr(args, 'entropy_target', 0.2)
        # the reference's clip_grad_norm_(params, 50) never clips: `params` is the generator train.py:39-44 makes
        # once, consumed by the first call while the shared grads are still None (pinned by
        # the golden-vector tests).  0 reproduces that effective behaviour, 50 is the written intent.
        self.max_grad_norm = float(getattr(args, 'max_grad_norm', 0.0))
        self.gpu_id = self.device.index if self.device.type == 'cuda' else -1
        self.state = state
        self.eps_len = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        self.n_steps = 0  # env-steps taken (all envs)
        self.done = None
        self.reward = None
        self.info = None
        # persistent recurrent-state storage: a rollout starts from it and writes its final state back in place, so the
        # whole iteration reads and writes fixed addresses (needed for CUDA-graph replay, harmless otherwise)
        self.hx_store = torch.zeros(self.num_envs, self.num_agents, self.rnn_out, device=self.device)
        self.cx_store = torch.zeros(self.num_envs, self.num_agents, self.rnn_out, device=self.device)
        self.hxs, self.cxs = self.hx_store, self.cx_store
        self._w_ent_key = (float(args.entropy), float(self.w_entropy_target))
        self.w_ent = torch.tensor(list(self._w_ent_key), device=self.device)
        self.lib = _lib.load() if self.device.type == 'cuda' else None
        # the hand-written forward / backward (learner.FusedA3C) for the configurations it covers; the autograd path of model.py
        # otherwise (CPU tensors, single-agent, stacked frames, Full observations)
        self.engine = None
        if self.device.type == 'cuda' and getattr(args, 'fused', True) and learner.supported(model, env, args):
            self.engine = learner.FusedA3C(model, self.num_envs, int(args.num_steps), self.device, seed=int(getattr(args, 'seed', 1)) + 7919 * max(self.gpu_id, 0))
            self.eps_len = self.engine.eps_len
        self._alloc_rollout()
        self.clear_actions()

    # ---- rollout storage -------------------------------------------------------------------------
    def _alloc_rollout(self):
        T, E = int(self.args.num_steps), self.num_envs
        obs_shape = tuple(self.env.obs.shape[1:])
        # the env writes straight into these (no copies): obs_buf[t + 1] <- step t
        if self.engine is not None:
            self.obs_buf = self.engine.obs  # uint8: the env's lossless observation encoding (values 0, 1, 2, 4)
        else:
            if E % 2:  # the step kernel stores 16-byte vectors: a float32 slot (E x 1,352 bytes) must keep that alignment
                raise _lib.Track2DError("num_envs must be even for device-resident rollouts (16-byte aligned observation slots)")
            self.obs_buf = torch.zeros((T + 1, E) + obs_shape, dtype=torch.float32, device=self.device)
        self.rew_buf = torch.zeros((T, E, 2), dtype=torch.float32, device=self.device)
        self.done_buf = torch.zeros((T, E), dtype=torch.uint8, device=self.device)
        self.val_buf = torch.zeros((T + 1, E, 2), dtype=torch.float32, device=self.device)
        self.ret_buf = torch.zeros((T, E, 2), dtype=torch.float32, device=self.device)
        self.gae_buf = torch.zeros((T, E, 2), dtype=torch.float32, device=self.device)
        self.t = 0

    def clear_actions(self):
        self.values, self.log_probs, self.entropies, self.preds = [], [], [], []
        self.t = 0
        self._prefetched = None  # (slot, bootstrap) of a forward the pipelined host path has already run for the next call
        return self

    # ---- episode / recurrent state ---------------------------------------------------------------
    def reset(self):
        """env.reset() for every env + zero LSTM state (player_util.py:84-102)"""
        self.env.reset()
        self.obs_buf[0].copy_(self.env.obs)
        self.state = self.obs_buf[0]
        self.eps_len.zero_()
        self.reset_rnn_hiden()
        self.clear_actions()

    def reset_rnn_hiden(self):
        self.hx_store.zero_()
        self.cx_store.zero_()
        self.hxs, self.cxs = self.hx_store, self.cx_store
        if self.engine is not None:
            self.engine.reset_state()

    def update_rnn_hiden(self):
        """truncate BPTT at the rollout boundary (player_util.py:104-106)"""
        if self.engine is not None:  # the fused path keeps no graph: slot 0 already holds the state the rollout starts from
            return
        if self.hxs is not self.hx_store:
            self.hx_store.copy_(self.hxs.detach())
            self.cx_store.copy_(self.cxs.detach())
        self.hxs, self.cxs = self.hx_store, self.cx_store

    # ---- acting ----------------------------------------------------------------------------------
    def action_train(self, forced_actions=None, host=None):
        """one step of every env: policy forward -> env.step -> buffers (player_util.py:44-67).
        host=None: device-resident (the env kernel writes straight into the rollout buffers).
        host=<pinned buffers from env.alloc_host_buffers()>: the reference's data flow -- actions to host
        numpy, env.step on host buffers (C ABI: H2D, kernels, D2H), observation back to the device
        (player_util.py:54-59 `torch.from_numpy(state_multi).float().to(device)`)."""
        t = self.t
        assert t < self.rew_buf.shape[0], "rollout buffer full: call optimize()"
        if self.engine is not None:
            return self._action_train_fused(t, forced_actions, host)
        value, action, entropy, log_prob, (hxs, cxs), R_pred = self.model((self.state, (self.hxs, self.cxs)), False, forced_actions)
        actions32 = action.to(torch.int32).contiguous()
        if host is None:
            self.env.step_into(actions32, self.obs_buf[t + 1], self.rew_buf[t], self.done_buf[t])
        else:
            host['actions'].copy_(actions32)  # D2H, synchronous
            self.env.step_host(host['actions'], host['obs'], host['reward'], host['done'])
            # H2D from pinned memory.  Through .data: the rollout slots share one allocation (one autograd version
            # counter) and slot t is already saved for backward when slot t + 1 is filled.
            if host['obs'].dtype == torch.uint8:  # upload a quarter of the bytes, widen to float32 on the device
                if getattr(self, '_obs_u8', None) is None:
                    self._obs_u8 = torch.empty(host['obs'].shape, dtype=torch.uint8, device=self.device)
                self._obs_u8.copy_(host['obs'], non_blocking=True)
                self.obs_buf.data[t + 1].copy_(self._obs_u8)
            else:
                self.obs_buf.data[t + 1].copy_(host['obs'], non_blocking=True)
            self.rew_buf.data[t].copy_(host['reward'], non_blocking=True)
            self.done_buf.data[t].copy_(host['done'], non_blocking=True)
        self.reward = self.rew_buf[t]
        self.done = self.done_buf[t]
        self.state = self.obs_buf[t + 1]
        # envs that finished were reset by the env: their next step starts from zero LSTM state (train.py:73-74)
        keep = (1 - self.done.to(hxs.dtype)).view(-1, 1, 1)
        self.hxs, self.cxs = hxs * keep, cxs * keep
        self.eps_len.add_(1).mul_(1 - self.done.to(torch.int32))
        self.n_steps += self.num_envs
        self.values.append(value)
        self.entropies.append(entropy)
        self.log_probs.append(log_prob)
        self.preds.append(R_pred)
        self.last_actions = actions32
        self.t = t + 1
        return self

    def _action_train_fused(self, t, forced_actions, host):
        eng = self.engine
        if forced_actions is not None:
            forced_actions = forced_actions.to(device=self.device, dtype=torch.int32).contiguous()
        if forced_actions is None and self._prefetched == (t, False):
            actions32 = eng.actions[t]  # this step's forward already ran, slice by slice, while its observations were arriving
        else:
            actions32 = eng.forward(t, forced=forced_actions)
        self._prefetched = None
        if host is None:
            self.env.step_into(actions32, self.obs_buf[t + 1], self.rew_buf[t], self.done_buf[t])
            eng.post_step(t, self.done_buf[t])
        else:
            self._step_through_host(t, actions32, host)
        self.reward = self.rew_buf[t]
        self.done = self.done_buf[t]
        self.state = self.obs_buf[t + 1]
        self.n_steps += self.num_envs
        self.last_actions = actions32
        self.t = t + 1
        return self

    def _step_through_host(self, t, actions32, host):
        """env.step through the host-buffer C ABI -- the host round trip of the reference's data flow (player_util.py:54-59) -- as a
        three-stage pipeline over slices of consecutive envs: the env's D2H of observation chunk c + 1 (library stream), the re-upload
        of chunk c (copy stream; PCIe is full duplex, the dependency is a device-side event) and, once the chunks of a slice of the
        batch are back on the device, the NEXT policy step of that slice (this stream) while the rest is still on the bus.  The next
        action_train / optimize call finds its forward done (`_prefetched`); the arithmetic is the whole-batch forward's, row for row."""
        eng, env, dev = self.engine, self.env, self.device
        host['actions'].copy_(actions32)  # D2H, synchronous
        C_, ends = host_pipeline_plan(self.num_envs, host['obs'].numel() * host['obs'].element_size(), host.get('chunks'), host.get('forward_slices'))
        env.step_host_begin(host['actions'], host['obs'], host['reward'], host['done'], C_)
        f32 = host['obs'].dtype != torch.uint8
        if f32 and getattr(self, '_obs_f32', None) is None:
            self._obs_f32 = torch.empty(host['obs'].shape, dtype=torch.float32, device=dev)
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        main, copy = torch.cuda.current_stream(dev), self._copy_stream
        copy.wait_stream(main)  # the slots about to be overwritten are no longer read (previous rollout's backward, previous narrow)
        landed, ev_flags = [], None
        with torch.cuda.stream(copy):  # nothing but copies on this stream: the H2D engine never waits for a kernel
            for c in range(C_):
                lo, hi = env.chunk_bounds(c, C_)
                env.host_chunk_wait(c, on_stream=True)  # device-side dependency: the H2D engine trails the D2H engine by one chunk
                if c == 0:
                    self.rew_buf[t].copy_(host['reward'], non_blocking=True)
                    self.done_buf[t].copy_(host['done'], non_blocking=True)
                    ev_flags = copy.record_event()
                if hi > lo:
                    (self._obs_f32 if f32 else self.obs_buf[t + 1])[lo:hi].copy_(host['obs'][lo:hi], non_blocking=True)
                if c + 1 in ends:  # chunk c completes a forward slice
                    landed.append((hi, copy.record_event()))
        main.wait_event(ev_flags)
        eng.post_step(t, self.done_buf[t])
        prefetch = bool(host.get('prefetch', True))
        boot = t + 1 == eng.T  # slot T is the value-only forward of optimize (player_util.py:110-116)
        first = 0
        for hi, ev in landed:
            main.wait_event(ev)
            if hi <= first:
                continue
            if f32:  # float32 on the host side (the reference's dtype): narrowed on the device (lossless: the values are 0, 1, 2, 4)
                self.obs_buf[t + 1, first:hi].copy_(self._obs_f32[first:hi])
            if prefetch:
                eng.forward(t + 1, bootstrap=boot, envs=(first, hi))
            first = hi
        self._prefetched = (t + 1, boot) if prefetch else None

    @property
    def fused(self):
        return self.engine is not None

    # recurrent state as the reference exposes it, (E, 2, 128) each; with the fused engine: the state entering the current step
    @property
    def hxs(self):
        return self.engine.hidden(self.t)[0] if self.engine is not None else self._hxs

    @hxs.setter
    def hxs(self, v):
        self._hxs = v

    @property
    def cxs(self):
        return self.engine.hidden(self.t)[1] if self.engine is not None else self._cxs

    @cxs.setter
    def cxs(self, v):
        self._cxs = v

    def _action_test_fused(self):
        eng = self.engine
        actions32 = eng.forward(0, greedy=True, bootstrap=True)
        obs, reward, done = self.env.step(actions32)
        self.reward, self.done = reward, done
        self.obs_buf[0].copy_(obs)
        eng.post_step(0, done)
        eng.carry_over(1, obs=False)
        self.state = self.obs_buf[0]
        self.n_steps += self.num_envs
        return self

    def action_test(self):
        """greedy step (player_util.py:69-82); used by the evaluator"""
        if self.engine is not None:
            return self._action_test_fused()
        with torch.no_grad():
            value, action, entropy, log_prob, (hxs, cxs), R_pred = self.model((self.state, (self.hxs, self.cxs)), True)
        actions32 = action.to(torch.int32).contiguous()
        obs, reward, done = self.env.step(actions32)
        self.reward, self.done = reward, done
        self.obs_buf[0].copy_(obs)
        self.state = self.obs_buf[0]
        keep = (1 - done.to(hxs.dtype)).view(-1, 1, 1)
        self.hxs, self.cxs = hxs * keep, cxs * keep
        self.eps_len.add_(1).mul_(1 - done.to(torch.int32))
        self.n_steps += self.num_envs
        return self

    # ---- learning --------------------------------------------------------------------------------
    def _returns_and_gae(self, T):
        E = self.num_envs
        if self.lib is not None:
            p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
            _lib.check(self.lib.track2d_gae_returns(p(self.rew_buf), p(self.done_buf), p(self.val_buf), p(self.ret_buf), p(self.gae_buf),
                                                    T, E, float(self.args.gamma), float(self.args.tau),
                                                    C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), self.lib)
        else:
            raise _lib.Track2DError("Agent.optimize needs the CUDA library; there is no CPU fallback")
        return self.ret_buf[:T], self.gae_buf[:T]

    def optimize(self, params, optimizer, shared_model, training_mode, device_share=None, world_size=1, allreduce=None,
                 boot_forced_actions=None, apply=True):
        """Agent.optimize (player_util.py:108-161): bootstrap, n-step return + GAE, A3C losses of both agents,
        aux reward-prediction L1, backward, (all-reduce), [clip] + SharedAdam.  `params`, `shared_model` and
        `device_share` are accepted for call compatibility: the model IS the shared model here."""
        T, E = self.t, self.num_envs
        assert T > 0
        self.env.join()  # side-stream work of the env (standby worlds, plans made ahead) rejoins the learner's stream once per rollout
        if self.engine is not None:
            return self._optimize_fused(T, optimizer, training_mode, world_size, allreduce, boot_forced_actions, apply)
        with torch.no_grad():  # player_util.py:110-116, value only matters where the episode is still running
            v_boot, _, _, _, _, _ = self.model((self.state, (self.hxs, self.cxs)), False, boot_forced_actions)
            self.val_buf[T].copy_(v_boot)
            self.val_buf[:T].copy_(torch.stack([v.detach() for v in self.values], 0))
        returns, gae = self._returns_and_gae(T)

        values = torch.stack(self.values, 0)        # (T, E, 2)
        log_probs = torch.stack(self.log_probs, 0)  # (T, E, 2)
        entropies = torch.stack(self.entropies, 0)  # (T, E, 2)
        key = (float(self.args.entropy), float(self.w_entropy_target))
        if key != self._w_ent_key:  # w_entropy_target is assigned after construction (train.py:53)
            self._w_ent_key = key
            self.w_ent = torch.tensor(list(key), device=self.device)
        w_ent = self.w_ent
        advantage = returns - values
        value_loss = (0.5 * advantage.pow(2)).sum(0)                       # (E, 2)   :133
        policy_loss = (-(log_probs * gae) - w_ent * entropies).sum(0)      # (E, 2)   :137-139
        loss_tracker = policy_loss[:, 0] + 0.5 * value_loss[:, 0]          # :143
        loss_target = policy_loss[:, 1] + 0.5 * value_loss[:, 1]
        pred_loss = torch.zeros(E, device=self.device)
        use_aux = 'reward' in self.args.aux and self.preds[0] is not None
        if use_aux:  # L1 between the target's prediction and the TRACKER's reward (:128-129)
            preds = torch.stack(self.preds, 0).squeeze(-1)                 # (T, E)
            pred_loss = (preds - self.rew_buf[:T, :, 0]).abs().sum(0)
        if training_mode == 0:
            loss = loss_tracker
        elif training_mode == 1:
            loss = loss_target
        else:
            loss = loss_tracker + loss_target
        if use_aux and training_mode != 0:
            loss = loss + pred_loss
        optimizer.zero_grad()
        loss.mean().backward()  # mean over envs == average of the per-worker gradients
        self._stats = (policy_loss.detach(), value_loss.detach(), entropies.detach().sum(0), pred_loss.detach())
        if apply:
            self.apply_update(optimizer, world_size, allreduce)
        return self._stats

    def _optimize_fused(self, T, optimizer, training_mode, world_size, allreduce, boot_forced_actions, apply):
        eng = self.engine
        if boot_forced_actions is not None:
            boot_forced_actions = boot_forced_actions.to(device=self.device, dtype=torch.int32).contiguous()
        if boot_forced_actions is not None or self._prefetched != (T, True):
            eng.forward(T, forced=boot_forced_actions, bootstrap=True)  # player_util.py:110-116 (it samples, like the reference)
        self._prefetched = None
        key = (float(self.args.entropy), float(self.w_entropy_target))
        optimizer.zero_grad()
        use_aux = 'reward' in self.args.aux
        self._stats = eng.backward(T, self.rew_buf, self.done_buf, training_mode, key, use_aux, float(self.args.gamma), float(self.args.tau),
                                   1.0 / self.num_envs)  # loss.mean() over envs == average of the per-worker gradients
        if apply:
            self.apply_update(optimizer, world_size, allreduce)
        return self._stats

    def apply_update(self, optimizer, world_size=1, allreduce=None, skip_allreduce=False):
        """second half of optimize(): [all-reduce of the flat gradient], [clip] + SharedAdam, and the hand-over of observation /
        recurrent state to the next rollout.  Split off so that a multi-GPU run can replay the two halves as CUDA graphs with the
        NCCL all-reduce launched eagerly between them."""
        fused_exchange = None
        if allreduce is not None and world_size > 1 and not skip_allreduce:
            if hasattr(allreduce, 'h') and hasattr(optimizer, 'fp'):  # peer.PeerAllReduce: the exchange rides in the optimizer kernel
                fused_exchange = allreduce
            else:
                allreduce(optimizer.fp.grad)
        if fused_exchange is not None:
            optimizer.step(max_grad_norm=self.max_grad_norm, grad_scale=1.0 / world_size, peer=fused_exchange)
        else:
            optimizer.step(max_grad_norm=self.max_grad_norm, grad_scale=1.0 / world_size)
        if self.engine is not None:
            T = self.t
            self.engine.mark_dirty()
            self.engine.carry_over(T)
            self.clear_actions()
            self.state = self.obs_buf[0]
            return self._stats
        self.clear_actions()
        self.obs_buf.data[0].copy_(self.state)
        self.state = self.obs_buf[0]
        self.hx_store.copy_(self.hxs.detach())  # the next rollout starts from the stored state (fixed addresses)
        self.cx_store.copy_(self.cxs.detach())
        self.hxs, self.cxs = self.hx_store, self.cx_store
        return self._stats
