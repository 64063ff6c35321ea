"""Batched, device-resident gym-track2d: the host-side mirror of the reference's env interface over the
libtrack2d C ABI.

    env = Track2DVecEnv('Track2D-BlockPartialPZR-v0', num_envs=65536, device='cuda:0', seed=1)
    obs = env.reset()                       # float32 cuda tensor (E, 2, 1, 13, 13)
    obs, reward, done = env.step(actions)   # actions: int32 cuda tensor (E, 2)

Per env this is `gym.make(id)` of the reference (Track1v1Env + TimeLimit(500),
gym_track2d/__init__.py:3-18) followed by environment.py:128-156 frame_stack with stack_frames=1 in its
float32 cast: the per-env observation (2, 1, 13, 13) gains a leading env axis.  PyTorch only provides
the device buffers and the CUDA stream; every transition is computed by the CUDA kernels in csrc/.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .spaces import Box, Discrete


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _np(a):
    return a.ctypes.data_as(C.c_void_p)


class Track2DVecEnv(object):
    """E independent Track1v1Env instances advanced by one kernel launch per step."""

    def __init__(self, env_id=None, num_envs=1, device="cuda:0", seed=1, rng="philox", auto_reset=True,
                 keep_f64=False, obs_dtype=torch.float32, map_type=None, obs_type=None, target_mode=None, level=0,
                 max_episode_steps=500, plan_ahead=False):
        if env_id is not None:
            map_type, obs_type, target_mode, level = _lib.parse_env_id(env_id)
        if obs_type not in _lib.OBS:
            raise TypeError("Observation type must be either 'full' or 'partial'")  # track_1v1.py:261
        if not torch.cuda.is_available():
            raise _lib.Track2DError("Track2DVecEnv needs a CUDA device; there is no CPU fallback")
        self.lib = _lib.load()
        self.env_id = env_id or "Track2D-%s%s%s-v%d" % (map_type, obs_type, target_mode, level)
        self.map_type, self.obs_type, self.target_mode, self.level = map_type, obs_type, target_mode, level
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.rng = rng
        self.auto_reset = bool(auto_reset)
        assert obs_dtype in (torch.float32, torch.uint8)
        self.obs_dtype = obs_dtype
        # plan_ahead: prepare next-episode worlds / Nav plans ahead of time on a side stream instead of inside step() (same results)
        flags = ((_lib.FLAG_AUTO_RESET if auto_reset else 0) | (_lib.FLAG_KEEP_F64 if keep_f64 else 0)
                 | (_lib.FLAG_PLAN_AHEAD if plan_ahead else 0))
        cfg = _lib.Config(_lib.ABI_VERSION, self.num_envs, _lib.MAP[map_type], _lib.OBS[obs_type], _lib.TARGET[target_mode],
                          int(level), _lib.RNG[rng], self.device.index, int(max_episode_steps), flags, int(seed) & (2 ** 64 - 1))
        h = C.c_void_p(0)
        _lib.check(self.lib.track2d_create(C.byref(cfg), C.byref(h)), self.lib)
        self.h = h
        self.H, self.W = self.lib.track2d_map_height(h), self.lib.track2d_map_width(h)
        self.cells = self.lib.track2d_obs_cells(h)
        hw = (13, 13) if obs_type == "Partial" else (self.H, self.W)
        E = self.num_envs
        # caller-owned device buffers, handed to the library as raw pointers
        self.obs = torch.zeros((E, 2, 1) + hw, dtype=obs_dtype, device=self.device)
        self.reward = torch.zeros((E, 2), dtype=torch.float32, device=self.device)
        self.done = torch.zeros((E,), dtype=torch.uint8, device=self.device)
        # the reference exposes LISTS of per-agent spaces (track_1v1.py:48-55)
        self.observation_space = [Box(0, 6, (1,) + hw, np.float32) for _ in range(2)]
        self.action_space = [Discrete(4) for _ in range(2)]
        self.num_agents = 2

    # ---- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.track2d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- device API ----------------------------------------------------------------------------
    def reset(self, mask=None):
        """Reset every env (mask=None) or those with a non-zero mask byte; returns the obs tensor."""
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            assert mask.numel() == self.num_envs
        fn = self.lib.track2d_reset if self.obs_dtype == torch.float32 else self.lib.track2d_reset_u8
        _lib.check(fn(self.h, _ptr(mask), _ptr(self.obs), self._stream()), self.lib)
        return self.obs

    def step(self, actions):
        """actions: int32 tensor (E, 2) on the env's device.  Returns (obs, reward, done) -- views of the
        env-owned output buffers, overwritten by the next step."""
        if actions.dtype != torch.int32 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        if actions.shape != (self.num_envs, 2):
            raise TypeError("actions must have shape (num_envs, 2)")
        fn = self.lib.track2d_step if self.obs_dtype == torch.float32 else self.lib.track2d_step_u8
        _lib.check(fn(self.h, _ptr(actions), _ptr(self.obs), _ptr(self.reward), _ptr(self.done), self._stream()), self.lib)
        return self.obs, self.reward, self.done

    def step_into(self, actions, obs, reward, done):
        """Same as step() but writes into caller-provided tensors (rollout buffers): no copies."""
        fn = self.lib.track2d_step if obs.dtype == torch.float32 else self.lib.track2d_step_u8
        _lib.check(fn(self.h, _ptr(actions), _ptr(obs), _ptr(reward), _ptr(done), self._stream()), self.lib)

    def init_maze(self, mask=None):
        _lib.check(self.lib.track2d_init_maze(self.h, _ptr(mask), self._stream()), self.lib)

    # ---- host-buffer API -----------------------------------------------------------------------
    def alloc_host_buffers(self, obs_dtype=torch.float32):
        """pinned host buffers for step_host / reset_host (obs_dtype uint8: the quarter-traffic variant)"""
        E = self.num_envs
        hw = tuple(self.obs.shape[1:])
        return dict(actions=torch.zeros((E, 2), dtype=torch.int32).pin_memory(),
                    obs=torch.zeros((E,) + hw, dtype=obs_dtype).pin_memory(),
                    reward=torch.zeros((E, 2), dtype=torch.float32).pin_memory(),
                    done=torch.zeros((E,), dtype=torch.uint8).pin_memory())

    def reset_host(self, obs_host, mask_host=None):
        fn = self.lib.track2d_reset_host if obs_host.dtype == torch.float32 else self.lib.track2d_reset_host_u8
        _lib.check(fn(self.h, _ptr(mask_host), _ptr(obs_host)), self.lib)
        return obs_host

    def step_host(self, actions_host, obs_host, reward_host, done_host):
        """numpy-facing step: host actions in, host obs/reward/done out (H2D + kernels + D2H inside)."""
        fn = self.lib.track2d_step_host if obs_host.dtype == torch.float32 else self.lib.track2d_step_host_u8
        _lib.check(fn(self.h, _ptr(actions_host), _ptr(obs_host), _ptr(reward_host), _ptr(done_host)), self.lib)
        return obs_host, reward_host, done_host

    def join(self):
        """make the current stream wait for the handle's side-stream work (standby worlds, plans made ahead): call before a CUDA-graph
        capture containing step() ends; a no-op for handles without side work"""
        _lib.check(self.lib.track2d_join(self.h, self._stream()), self.lib)

    def step_host_begin(self, actions_host, obs_host, reward_host, done_host, n_chunks=8):
        """pipelined step_host: enqueue H2D actions, kernels and the D2H of obs in n_chunks pieces; host_chunk_wait(c) blocks until
        envs chunk_bounds(c, n_chunks) of obs_host (and, with chunk 0, reward / done) have arrived"""
        _lib.check(self.lib.track2d_step_host_begin(self.h, _ptr(actions_host), _ptr(obs_host), int(obs_host.dtype == torch.uint8), _ptr(reward_host),
                                                    _ptr(done_host), int(n_chunks)), self.lib)

    def host_chunk_wait(self, chunk, on_stream=False):
        """block the host until chunk `chunk` is in the host buffers; on_stream: make the current CUDA stream wait for it instead (what
        is enqueued next -- the re-upload -- starts the moment the chunk has landed, with no host wake-up in between)"""
        if on_stream:
            _lib.check(self.lib.track2d_host_chunk_wait_stream(self.h, int(chunk), self._stream()), self.lib)
        else:
            _lib.check(self.lib.track2d_host_chunk_wait(self.h, int(chunk)), self.lib)

    def chunk_bounds(self, chunk, n_chunks):
        E = self.num_envs
        return E * chunk // n_chunks, E * (chunk + 1) // n_chunks

    # ---- state read-back / injection (synchronous; tests and the single-env shim) ----------------
    def get_maps(self, first=0, count=None):
        count = self.num_envs - first if count is None else count
        out = np.zeros((count, self.H, self.W), np.uint8)
        _lib.check(self.lib.track2d_get_maps(self.h, first, count, _np(out)), self.lib)
        return out

    def set_maps(self, maps, first=0):
        maps = np.ascontiguousarray(maps, np.uint8).reshape(-1, self.H, self.W)
        _lib.check(self.lib.track2d_set_maps(self.h, first, maps.shape[0], _np(maps)), self.lib)

    def get_agents(self, first=0, count=None):
        count = self.num_envs - first if count is None else count
        pos, ctr = np.zeros((count, 2, 2), np.int32), np.zeros((count, 2), np.int32)
        _lib.check(self.lib.track2d_get_agents(self.h, first, count, _np(pos), _np(ctr)), self.lib)
        return pos, ctr

    def set_agents(self, pos=None, counters=None, first=0):
        n = None
        if pos is not None:
            pos = np.ascontiguousarray(pos, np.int32).reshape(-1, 2, 2)
            n = pos.shape[0]
        if counters is not None:
            counters = np.ascontiguousarray(counters, np.int32).reshape(-1, 2)
            n = counters.shape[0]
        _lib.check(self.lib.track2d_set_agents(self.h, first, n, _np(pos) if pos is not None else None,
                                               _np(counters) if counters is not None else None), self.lib)

    def get_goals(self, first=0, count=None):
        count = self.num_envs - first if count is None else count
        g = np.zeros((count, 2, 2), np.int32)
        _lib.check(self.lib.track2d_get_goals(self.h, first, count, _np(g)), self.lib)
        return g

    def get_ram(self, first=0, count=None):
        count = self.num_envs - first if count is None else count
        plan = np.zeros((count, _lib.RAM_MAXPLAN), np.int32)
        ln, idx = np.zeros(count, np.int32), np.zeros(count, np.int32)
        _lib.check(self.lib.track2d_get_ram(self.h, first, count, _np(plan), _np(ln), _np(idx)), self.lib)
        return plan, ln, idx

    def set_ram(self, plan, length, idx, first=0):
        plan = np.ascontiguousarray(plan, np.int32).reshape(-1, _lib.RAM_MAXPLAN)
        length = np.ascontiguousarray(length, np.int32).reshape(-1)
        idx = np.ascontiguousarray(idx, np.int32).reshape(-1)
        _lib.check(self.lib.track2d_set_ram(self.h, first, plan.shape[0], _np(plan), _np(length), _np(idx)), self.lib)

    def get_nav(self, first=0, count=None):
        count = self.num_envs - first if count is None else count
        plan = np.zeros((count, _lib.NAV_MAXPLAN), np.int32)
        ln, idx, goal = np.zeros(count, np.int32), np.zeros(count, np.int32), np.zeros((count, 2), np.int32)
        _lib.check(self.lib.track2d_get_nav(self.h, first, count, _np(plan), _np(ln), _np(idx), _np(goal)), self.lib)
        return plan, ln, idx, goal

    def astar_solve(self, start, goal, first=0):
        """AstarSolver.solve + get_actions (Astar_solver.py:102-149) on the generator maze of env first + i, from start[i] to
        goal[i]; returns (plans int32 [n][NAV_MAXPLAN], lengths int32 [n], -1 = unreachable).  Overwrites those envs' plans."""
        start = np.ascontiguousarray(start, np.int32).reshape(-1, 2)
        goal = np.ascontiguousarray(goal, np.int32).reshape(-1, 2)
        n = start.shape[0]
        plan, ln = np.zeros((n, _lib.NAV_MAXPLAN), np.int32), np.zeros(n, np.int32)
        _lib.check(self.lib.track2d_astar_solve(self.h, first, n, _np(start), _np(goal), _np(plan), _np(ln)), self.lib)
        return plan, ln

    def set_nav(self, plan, length, idx, goal, first=0):
        plan = np.ascontiguousarray(plan, np.int32).reshape(-1, _lib.NAV_MAXPLAN)
        length = np.ascontiguousarray(length, np.int32).reshape(-1)
        idx = np.ascontiguousarray(idx, np.int32).reshape(-1)
        goal = np.ascontiguousarray(goal, np.int32).reshape(-1, 2)
        _lib.check(self.lib.track2d_set_nav(self.h, first, plan.shape[0], _np(plan), _np(length), _np(idx), _np(goal)), self.lib)

    def get_rewards_f64(self, first=0, count=None):
        count = self.num_envs - first if count is None else count
        r = np.zeros((count, 2), np.float64)
        _lib.check(self.lib.track2d_get_rewards_f64(self.h, first, count, _np(r)), self.lib)
        return r

    def get_target_actions(self, first=0, count=None):
        count = self.num_envs - first if count is None else count
        a = np.zeros(count, np.int32)
        _lib.check(self.lib.track2d_get_target_actions(self.h, first, count, _np(a)), self.lib)
        return a

    def seed_env(self, index, seed):
        """T2D_RNG_NUMPY: np.random.seed(seed) for env `index`"""
        _lib.check(self.lib.track2d_seed_env(self.h, int(index), int(seed) & 0xFFFFFFFF), self.lib)

    def get_rng_numpy(self, index):
        key, pos = np.zeros(624, np.uint32), C.c_int32(0)
        _lib.check(self.lib.track2d_get_rng_numpy(self.h, int(index), _np(key), C.byref(pos)), self.lib)
        return key, int(pos.value)

    def status(self):
        s = C.c_uint32(0)
        _lib.check(self.lib.track2d_get_status(self.h, C.byref(s), self._stream()), self.lib)
        return int(s.value)

    def counters(self):
        ep, st = C.c_uint64(0), C.c_uint64(0)
        _lib.check(self.lib.track2d_get_counters(self.h, C.byref(ep), C.byref(st)), self.lib)
        return int(ep.value), int(st.value)


class Track1v1Env(object):
    """Single-env, numpy-facing drop-in for `gym.make('Track2D-...')` of the reference (Track1v1Env wrapped
    in TimeLimit(500)): same reset()/step()/seed()/close() surface, return types and dtypes
    (track_1v1.py:71-168), computed by the CUDA kernels with num_envs=1.

    The reference draws from the process-global numpy RNG; there is no such thing to share here, so the
    stream is owned by the env: seed(s) plays the role of np.random.seed(s) (rng='numpy' replays the
    reference's draw order bit for bit)."""

    metadata = {'render.modes': ['human', 'rgb_array']}

    def __init__(self, env_id=None, map_type='Block', obs_type='Partial', target_mode='PZR', level=0, device='cuda:0',
                 seed=None, rng='numpy', max_episode_steps=500):
        if env_id is not None:
            map_type, obs_type, target_mode, level = _lib.parse_env_id(env_id)
        self.map_type, self.obs_type, self.target_mode, self.level = map_type, obs_type, target_mode, level
        self.num_agents_max = self.num_agents = 2
        self.pob_size = 6
        self.action_type = 'VonNeumann'
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1)[0])  # the reference's default is OS entropy too
        self.vec = Track2DVecEnv(env_id, 1, device, seed, rng, auto_reset=False, keep_f64=True, map_type=map_type,
                                 obs_type=obs_type, target_mode=target_mode, level=level, max_episode_steps=max_episode_steps)
        self.observation_space = self.vec.observation_space
        self.action_space = self.vec.action_space
        self._host = self.vec.alloc_host_buffers()
        self._obs_np_dtype = np.int64 if map_type == 'Maze' else np.float64  # track_1v1.py:234 / generators.py:145,175
        self.traces, self.distance = [], 0.0
        if rng == 'numpy':
            self.vec.init_maze()  # Track1v1Env.__init__ draws one map before any reset (track_1v1.py:45)

    @property
    def unwrapped(self):
        return self

    # reference attributes some callers read
    @property
    def state(self):
        return self.vec.get_agents()[0][0].tolist()

    @property
    def maze(self):
        return self.vec.get_maps()[0].astype(self._obs_np_dtype)

    @property
    def goal_states(self):
        return self.vec.get_goals()[0].tolist()

    @property
    def C_far(self):
        return int(self.vec.get_agents()[1][0][0])

    def seed(self, seed=None):
        if seed is not None and self.vec.rng == 'numpy':
            self.vec.seed_env(0, seed)
        return [seed]

    def _obs(self):
        return self._host['obs'].numpy()[0].astype(self._obs_np_dtype)

    def reset(self):
        self.vec.reset_host(self._host['obs'])
        st = self.state
        self.traces = [st[0]]
        self.distance = float(abs(st[0][0] - st[1][0]) + abs(st[0][1] - st[1][1]))  # track_1v1.py:154
        return self._obs()

    def step(self, action):
        action = list(action)
        if len(action) < 2:
            raise TypeError("step() needs one action per agent")
        a = self._host['actions']
        a[0, 0], a[0, 1] = (int(np.asarray(x).reshape(-1)[0]) for x in action[:2])  # scalars, 0-d or 1-element arrays, as gym callers pass them
        self.vec.step_host(a, self._host['obs'], self._host['reward'], self._host['done'])
        rewards = self.vec.get_rewards_f64()[0]
        st = self.state
        self.distance = float(np.linalg.norm(np.array(st[1]) - np.array(st[0])))
        self.traces.append(st[1])
        info = {'distance': self.distance, 'traces': self.traces, 'traces_relative': []}
        return self._obs(), rewards, bool(self._host['done'][0]), info

    def render(self, mode='human', close=False):
        raise NotImplementedError("matplotlib rendering is out of scope (reference: track_1v1.py:170-216)")

    def close(self):
        self.vec.close()


def make(env_id, **kwargs):
    """gym.make for the 72 Track2D ids (gym_track2d/__init__.py:3-18)."""
    _lib.parse_env_id(env_id)
    return Track1v1Env(env_id, **kwargs)
