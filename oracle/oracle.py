"""ctypes front-end of the C oracle (oracle/track2d_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package (active_tracking_rl_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtrack2d_oracle.so")

MAP = {"Block": 0, "Maze": 1, "Empty": 2}
OBS = {"Partial": 0, "Full": 1}
TARGET = {"Adv": 0, "PZR": 1, "Far": 2, "Nav": 3, "Ram": 4, "RPF": 5}
NAV_MAXPLAN = 8192


def build(force=False):
    src = os.path.join(_HERE, "track2d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libtrack2d_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, u32, dbl = C.c_void_p, C.c_int, C.c_uint32, C.c_double
        L.o_create.restype = vp
        L.o_create.argtypes = [i32, i32, i32, i32]
        L.o_destroy.argtypes = [vp]
        L.o_seed.argtypes = [vp, u32]
        L.o_rng_draws.restype = C.c_uint64
        L.o_rng_draws.argtypes = [vp]
        L.o_rng_u32.restype = u32
        L.o_rng_u32.argtypes = [vp]
        L.o_rng_double.restype = dbl
        L.o_rng_double.argtypes = [vp]
        L.o_rng_randint.restype = i32
        L.o_rng_randint.argtypes = [vp, i32, i32]
        L.o_rng_permutation.argtypes = [vp, i32, vp]
        L.o_rng_get_state.argtypes = [vp, vp, vp]
        L.o_rng_set_state.argtypes = [vp, vp, i32]
        L.o_init_maze.argtypes = [vp]
        L.o_reset.argtypes = [vp, vp]
        L.o_step.restype = i32
        L.o_step.argtypes = [vp, vp, vp, vp, vp, vp]
        L.o_obs_cells.restype = i32
        L.o_obs_cells.argtypes = [vp]
        L.o_get_obs.argtypes = [vp, vp]
        L.o_height.restype = i32
        L.o_height.argtypes = [vp]
        L.o_width.restype = i32
        L.o_width.argtypes = [vp]
        L.o_get_maze.argtypes = [vp, vp]
        L.o_get_gen_maze.argtypes = [vp, vp]
        L.o_get_state.argtypes = [vp, vp, vp, vp]
        L.o_set_state.argtypes = [vp, i32, i32, vp, vp, i32, i32]
        L.o_get_ram.argtypes = [vp, vp, vp, vp]
        L.o_set_ram.argtypes = [vp, vp, i32, i32]
        L.o_get_nav.restype = i32
        L.o_get_nav.argtypes = [vp, vp, i32, vp, vp]
        L.o_set_nav.argtypes = [vp, vp, i32, i32, vp]
        L.o_get_nav_stats.argtypes = [vp, vp]
        L.o_astar.restype = i32
        L.o_astar.argtypes = [vp, vp, vp, vp, i32]
        L.o_run_random.restype = C.c_long
        L.o_run_random.argtypes = [i32, i32, i32, i32, u32, C.c_long, vp, vp]
        i64 = C.c_long
        L.o_batch_pipeline.restype = i64
        L.o_batch_pipeline.argtypes = [i32, i32, i32, i32, i64, i64, i64, u32, i32] + [vp] * 12 + [i32]
        L.o_batch_inject_steps.restype = None
        L.o_batch_inject_steps.argtypes = [i32, i32, i32, i32, i64, i64, i64, i32, i32, i32] + [vp] * 9 + [i32]
        _lib = L
    return _lib


def parse_env_id(env_id):
    """'Track2D-BlockPartialPZR-v0' -> (map, obs, target, level); grammar of gym_track2d/__init__.py:3-18"""
    assert env_id.startswith("Track2D-"), env_id
    body, ver = env_id[len("Track2D-"):].rsplit("-v", 1)
    for m in MAP:
        if body.startswith(m):
            rest = body[len(m):]
            for o in OBS:
                if rest.startswith(o):
                    t = rest[len(o):]
                    if t in TARGET:
                        return m, o, t, int(ver)
    raise KeyError(env_id)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleEnv(object):
    """One reference-semantics env.  seed(s) == np.random.seed(s) on the reference's global RNG."""

    def __init__(self, env_id=None, map_type="Block", obs_type="Partial", target_mode="PZR", level=0):
        if env_id is not None:
            map_type, obs_type, target_mode, level = parse_env_id(env_id)
        self.L = lib()
        self.map_type, self.obs_type, self.target_mode, self.level = map_type, obs_type, target_mode, level
        self.h = self.L.o_create(MAP[map_type], OBS[obs_type], TARGET[target_mode], level)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.o_destroy(self.h)
            self.h = None

    # --- RNG -------------------------------------------------------------------------------
    def seed(self, s):
        self.L.o_seed(self.h, int(s) & 0xFFFFFFFF)

    def rng_state(self):
        key = np.zeros(624, np.uint32)
        pos = C.c_int(0)
        self.L.o_rng_get_state(self.h, _p(key), C.byref(pos))
        return key, pos.value

    def set_rng_state(self, key, pos):
        key = np.ascontiguousarray(key, np.uint32)
        self.L.o_rng_set_state(self.h, _p(key), int(pos))

    def rng_draws(self):
        return int(self.L.o_rng_draws(self.h))

    # --- env -------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.L.o_height(self.h), self.L.o_width(self.h)

    def _obs_buf(self):
        n = self.L.o_obs_cells(self.h)
        return np.zeros((2, n), np.uint8)

    def _shape_obs(self, o):
        if self.obs_type == "Full":
            H, W = self.shape
            return o.reshape(2, 1, H, W)
        return o.reshape(2, 1, 13, 13)

    def init_maze(self):
        self.L.o_init_maze(self.h)

    def reset(self):
        self.L.o_init_maze  # noqa: B018 (doc: reset == init_maze + target reset + counters)
        o = np.zeros((2, 82 * 82), np.uint8)
        self.L.o_reset(self.h, _p(o))
        n = self.L.o_obs_cells(self.h)
        return self._shape_obs(np.ascontiguousarray(o.reshape(-1)[: 2 * n].reshape(2, n)))

    def obs(self):
        o = self._obs_buf()
        self.L.o_get_obs(self.h, _p(o))
        return self._shape_obs(o)

    def step(self, actions):
        a = np.asarray([int(actions[0]), int(actions[1])], np.int32)
        o = self._obs_buf()
        rew = np.zeros(2, np.float64)
        de, ta = C.c_int(0), C.c_int(0)
        done = self.L.o_step(self.h, _p(a), _p(o), _p(rew), C.byref(de), C.byref(ta))
        return self._shape_obs(o), rew, bool(done), {"done_env": bool(de.value), "target_action": ta.value}

    def maze(self):
        H, W = self.shape
        m = np.zeros((H, W), np.uint8)
        self.L.o_get_maze(self.h, _p(m))
        return m

    def gen_maze(self):
        H, W = self.shape
        m = np.zeros((H, W), np.uint8)
        self.L.o_get_gen_maze(self.h, _p(m))
        return m

    def state(self):
        st, g, c = np.zeros((2, 2), np.int32), np.zeros((2, 2), np.int32), np.zeros(2, np.int32)
        self.L.o_get_state(self.h, _p(st), _p(g), _p(c))
        return st, g, int(c[0]), int(c[1])

    def set_state(self, maze, state, c_far=0, elapsed=0):
        maze = np.ascontiguousarray(maze, np.uint8)
        st = np.ascontiguousarray(state, np.int32)
        self.L.o_set_state(self.h, maze.shape[0], maze.shape[1], _p(maze), _p(st), int(c_far), int(elapsed))

    def ram(self):
        plan = np.zeros(16, np.int32)
        n, i = C.c_int(0), C.c_int(0)
        self.L.o_get_ram(self.h, _p(plan), C.byref(n), C.byref(i))
        return plan[: n.value].copy(), i.value

    def set_ram(self, plan, idx):
        plan = np.ascontiguousarray(plan, np.int32)
        self.L.o_set_ram(self.h, _p(plan), len(plan), int(idx))

    def nav(self):
        plan = np.zeros(NAV_MAXPLAN, np.int32)
        i = C.c_int(0)
        goal = np.zeros(2, np.int32)
        n = self.L.o_get_nav(self.h, _p(plan), NAV_MAXPLAN, C.byref(i), _p(goal))
        return plan[:n].copy(), i.value, goal

    def set_nav(self, plan, idx, goal):
        plan = np.ascontiguousarray(plan, np.int32)
        goal = np.ascontiguousarray(goal, np.int32)
        self.L.o_set_nav(self.h, _p(plan), len(plan), int(idx), _p(goal))

    def nav_stats(self):
        s = np.zeros(4, np.int32)
        self.L.o_get_nav_stats(self.h, _p(s))
        return dict(replans=int(s[0]), planb=int(s[1]), expansions=int(s[2]), replaces=int(s[3]))

    def astar(self, start, goal):
        """A* on the generator maze; returns list of actions or None if unsolvable."""
        plan = np.zeros(NAV_MAXPLAN, np.int32)
        s = np.asarray(start, np.int32)
        g = np.asarray(goal, np.int32)
        n = self.L.o_astar(self.h, _p(s), _p(g), _p(plan), NAV_MAXPLAN)
        return None if n < 0 else plan[:n].copy()


def run_random(env_id, seed, n_steps):
    """env-only CPU baseline loop (random_agent_multi.py:34-51 shape); returns (steps, reward_sum, episodes)."""
    m, o, t, lvl = parse_env_id(env_id)
    rs = np.zeros(2, np.float64)
    ne = C.c_long(0)
    n = lib().o_run_random(MAP[m], OBS[o], TARGET[t], lvl, int(seed) & 0xFFFFFFFF, int(n_steps), _p(rs), C.byref(ne))
    return int(n), rs, int(ne.value)


def _threads():
    return max(1, min(os.cpu_count() or 1, 64))


def batch_pipeline(env_id, E, seed0, actions, first=0, n=None, want_maze=False):
    """Env e of an E-wide batch == the reference after np.random.seed(seed0 + e): reset, then T = len(actions) steps
    with reset-on-done (libtrack2d's auto-reset convention).  ONE ctypes call for the whole batch (threads inside).
    actions int32 [T][E][2].  Returns a dict of arrays indexed [t][e] (rows outside [first, first + n) stay zero)."""
    m, o, t, lvl = parse_env_id(env_id)
    L = lib()
    actions = np.ascontiguousarray(actions, np.int32)
    T = actions.shape[0]
    assert actions.shape == (T, E, 2)
    n = E - first if n is None else n
    H = 81 if m == "Maze" else 82
    cells = H * H if o == "Full" else 169
    out = dict(
        reset_obs=np.zeros((E, 2, cells), np.uint8), reset_pos=np.zeros((E, 2, 2), np.int32),
        reset_maze=np.zeros((E, H, H), np.uint8) if want_maze else None,
        pos=np.zeros((T, E, 2, 2), np.int32), ctr=np.zeros((T, E, 2), np.int32), rew=np.zeros((T, E, 2), np.float64),
        done=np.zeros((T, E), np.uint8), tgt_act=np.zeros((T, E), np.int32), obs=np.zeros((T, E, 2, cells), np.uint8),
        rng_ckpt=np.zeros((E, 9), np.uint32), plan_meta=np.zeros((T, E, 2), np.int32))
    q = lambda k: _p(out[k]) if out[k] is not None else None  # noqa: E731
    out["n_done"] = int(L.o_batch_pipeline(MAP[m], OBS[o], TARGET[t], lvl, E, first, n, int(seed0) & 0xFFFFFFFF, T, _p(actions),
                                           q("reset_obs"), q("reset_maze"), q("reset_pos"), q("pos"), q("ctr"), q("rew"), q("done"),
                                           q("tgt_act"), q("obs"), q("rng_ckpt"), q("plan_meta"), _threads()))
    return out


def batch_inject_steps(env_id, maze, pos0, ctr0, actions):
    """T steps (no reset) of E envs from injected states; one ctypes call.  maze uint8 [E][H][W], pos0 int32 [E][2][2],
    ctr0 int32 [E][2] = (C_far, elapsed), actions int32 [T][E][2]."""
    m, o, t, lvl = parse_env_id(env_id)
    L = lib()
    maze = np.ascontiguousarray(maze, np.uint8)
    pos0 = np.ascontiguousarray(pos0, np.int32)
    ctr0 = np.ascontiguousarray(ctr0, np.int32)
    actions = np.ascontiguousarray(actions, np.int32)
    E, H, W = maze.shape
    T = actions.shape[0]
    assert actions.shape == (T, E, 2) and pos0.shape == (E, 2, 2) and ctr0.shape == (E, 2)
    cells = H * W if o == "Full" else 169
    out = dict(pos=np.zeros((T, E, 2, 2), np.int32), ctr=np.zeros((T, E, 2), np.int32), rew=np.zeros((T, E, 2), np.float64),
               done=np.zeros((T, E), np.uint8), obs=np.zeros((T, E, 2, cells), np.uint8))
    L.o_batch_inject_steps(MAP[m], OBS[o], TARGET[t], lvl, E, 0, E, H, W, T, _p(maze), _p(pos0), _p(ctr0), _p(actions),
                           _p(out["pos"]), _p(out["ctr"]), _p(out["rew"]), _p(out["done"]), _p(out["obs"]), _threads())
    return out
