"""Pins the CPU oracle (oracle/track2d_oracle.c) to the reference: bit-exact replay of every fixture
recorded from the unmodified reference env by oracle/refharness/make_golden.py, including the numpy
legacy MT19937 state after every reset and episode (=> same number and order of RNG draws)."""
import glob
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

EPISODE_FILES = sorted(glob.glob(os.path.join(GOLDEN, "episodes_*.npz")))


def _checkpoint(env):
    key, pos = env.rng_state()
    return np.concatenate([key[:8], np.asarray([pos], np.uint32)])


def test_fixtures_present():
    assert len(EPISODE_FILES) >= 15
    assert os.path.exists(os.path.join(GOLDEN, "astar_kat.npz"))


@pytest.mark.parametrize("path", EPISODE_FILES, ids=[os.path.basename(p)[9:-4] for p in EPISODE_FILES])
def test_oracle_replays_reference_episodes(path):
    g = np.load(path)
    env_id = "Track2D-" + os.path.basename(path)[9:-4]
    env = oracle.OracleEnv(env_id)
    is_ram, is_nav = "Ram" in env_id, ("Nav" in env_id or "RPF" in env_id)
    t0 = 0
    n_replans = 0
    for k in range(len(g["ep_length"])):
        if g["ep_seed"][k] >= 0:
            env.seed(int(g["ep_seed"][k]))
        obs = env.reset()
        st, goals, c_far, elapsed = env.state()
        assert (env.maze() == g["ep_maze"][k]).all(), "maze, episode %d" % k
        assert (env.gen_maze() == g["ep_gen_maze"][k]).all()
        assert (st == g["ep_init_state"][k]).all(), (st, g["ep_init_state"][k])
        assert (goals == g["ep_goals"][k]).all()
        assert (obs.reshape(2, -1) == g["ep_reset_obs"][k]).all()
        if is_ram:
            plan, idx = env.ram()
            assert list(plan) == list(g["ep_ram_plan"][k][: g["ep_ram_len"][k]]) and idx == 0
        if is_nav:
            plan, idx, goal = env.nav()
            assert list(plan) == list(g["ep_nav_plan"][k][: g["ep_nav_len"][k]])
            assert (goal == g["ep_nav_goal"][k]).all()
        assert (_checkpoint(env) == g["ep_rng_after_reset"][k]).all(), "RNG stream after reset"
        L = int(g["ep_length"][k])
        for t in range(t0, t0 + L):
            obs, rew, done, info = env.step(g["st_actions"][t])
            st, _, c_far, _ = env.state()
            assert (st == g["st_state"][t]).all(), (t, st, g["st_state"][t])
            # float64 rewards must be BIT-exact
            assert rew.tobytes() == g["st_rewards"][t].tobytes(), (t, rew, g["st_rewards"][t])
            assert done == bool(g["st_done"][t]), t
            assert c_far == int(g["st_c_far"][t])
            assert (obs.reshape(2, -1) == g["st_obs"][t]).all(), t
            if is_ram:
                plan, idx = env.ram()
                assert (idx, len(plan)) == (int(g["st_tgt_i"][t]), int(g["st_tgt_len"][t]))
            if is_nav:
                plan, idx, _ = env.nav()
                assert (idx, len(plan)) == (int(g["st_tgt_i"][t]), int(g["st_tgt_len"][t]))
                n_replans += int(idx == 1 and t > t0)
        t0 += L
        assert (_checkpoint(env) == g["ep_rng_after_episode"][k]).all(), "RNG stream after episode %d" % k
    if env_id == "Track2D-BlockPartialNav-v0":
        assert n_replans >= 1, "fixture should exercise a Navigator replan"


def test_oracle_astar_kat():
    g = np.load(os.path.join(GOLDEN, "astar_kat.npz"))
    n_unsolvable = 0
    for i in range(len(g["length"])):
        dim = int(g["dim"][i])
        env = oracle.OracleEnv(map_type="Maze" if dim == 81 else "Block", target_mode="Nav")
        maze = np.ascontiguousarray(g["maze"][i][:dim, :dim])
        env.set_state(maze, [[1, 1], [1, 1]])
        plan = env.astar(g["start"][i], g["goal"][i])
        if g["length"][i] < 0:
            assert plan is None
            n_unsolvable += 1
        else:
            assert plan is not None and list(plan) == list(g["plan"][i][: g["length"][i]]), i
        assert env.nav_stats()["replaces"] == int(g["replaces"][i])
    assert n_unsolvable >= 1


def test_reward_closed_form_anchors():
    """KATs readable straight off track_1v1.py:94-111."""
    maze = np.zeros((82, 82), np.uint8)
    maze[0, :] = maze[-1, :] = maze[:, 0] = maze[:, -1] = 1
    env = oracle.OracleEnv("Track2D-BlockPartialPZR-v0")

    def rewards_at(tr, tg):
        # bump both into... no: place agents and take a blocked move so positions stay put
        m = maze.copy()
        m[tr[0] - 1, tr[1]] = 1
        m[tg[0] - 1, tg[1]] = 1
        env.set_state(m, [tr, tg])
        _, r, _, _ = env.step([0, 0])
        return r

    assert np.allclose(rewards_at([40, 40], [40, 41]), [1 - 2 / 6.0, -(1 - 2 / 6.0)], rtol=0, atol=0)
    assert list(rewards_at([40, 40], [40, 40])) == [1.0, -1.0]
    assert list(rewards_at([40, 40], [40, 46])) == [-1.0, 1.0]
    assert list(rewards_at([40, 40], [40, 49])) == [-1.0, 0.5]  # PZR penalty beyond the FOV radius


def test_done_on_eleventh_far_step_and_timelimit():
    maze = np.zeros((82, 82), np.uint8)
    maze[0, :] = maze[-1, :] = maze[:, 0] = maze[:, -1] = 1
    maze[39, 40] = maze[59, 40] = 1  # both agents blocked upwards -> they stay in place
    env = oracle.OracleEnv("Track2D-BlockPartialAdv-v0")
    env.set_state(maze, [[40, 40], [60, 40]])
    dones = [env.step([0, 0])[2] for _ in range(11)]
    assert dones == [False] * 10 + [True]
    env.set_state(maze, [[40, 40], [60, 40]], c_far=0, elapsed=498)
    assert env.step([0, 0])[2] is False
    assert env.step([0, 0])[2] is True  # elapsed hits 500 (TimeLimit)


def test_obs_out_of_bounds_and_overlap():
    maze = np.zeros((82, 82), np.uint8)
    maze[0, :] = maze[-1, :] = maze[:, 0] = maze[:, -1] = 1
    env = oracle.OracleEnv("Track2D-BlockPartialAdv-v0")
    env.set_state(maze, [[1, 1], [1, 1]])
    o = env.obs()
    assert o[0, 0, 6, 6] == 2 and o[1, 0, 6, 6] == 4  # overlap: each sees its own colour
    assert (o[0, 0, :5, :] == 1).all() and (o[0, 0, :, :5] == 1).all()  # beyond the map reads wall
    assert o[0, 0, 5, 5] == 1 and o[0, 0, 6, 7] == 0
    env.set_state(maze, [[10, 10], [16, 4]])
    o = env.obs()
    assert o[0, 0, 12, 0] == 4 and o[1, 0, 0, 12] == 2  # visible at the FOV corner
    env.set_state(maze, [[10, 10], [17, 4]])
    o = env.obs()
    assert 4 not in o[0] and 2 not in o[1]
