"""GPU parity at the BASELINE.json config sizes: the CUDA path (through the libtrack2d C ABI) against the C oracle run
over the WHOLE batch (one ctypes call per batch, threads inside), bit-exact: positions, counters, float64 rewards,
dones, executed target actions, plan cursors, observations, maps, RNG stream positions.

  config 2/5  Track2D-BlockPartialPZR-v0  @ 65,536 envs   (numpy-RNG pipeline with auto-reset; and injected states, Philox worlds)
  config 3    Track2D-BlockPartialNav-v0  @ 16,384 envs   (A* plans, replans and resets inside the window)
  config 4    Track2D-MazePartialAdv-v0   @ 32,768 envs

plus the direct CUDA A* known-answer test on the 60 plans recorded from the reference's AstarSolver."""
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def t2d():
    from active_tracking_rl_b200 import envs
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return envs


def _actions(T, E, seed, tracker_runs_away=True):
    rs = np.random.RandomState(seed)
    acts = rs.randint(0, 4, size=(T, E, 2)).astype(np.int32)
    if tracker_runs_away:  # most envs: the tracker walks up and the target down, so far-counter dones (and auto-resets) happen
        run = rs.rand(E) < 0.7
        acts[:, run, 0] = 0
        acts[:, run, 1] = 1
    return acts


@pytest.mark.parametrize("env_id,E,T", [("Track2D-BlockPartialPZR-v0", 65536, 24), ("Track2D-BlockPartialNav-v0", 16384, 24),
                                        ("Track2D-MazePartialAdv-v0", 32768, 24)],
                         ids=["PZR-65536", "Nav-16384", "MazeAdv-32768"])
def test_numpy_pipeline_at_config_size_matches_oracle(t2d, env_id, E, T):
    S = 20260000
    acts = _actions(T, E, 5)
    ref = oracle.batch_pipeline(env_id, E, S, acts, want_maze=True)
    env = t2d.Track2DVecEnv(env_id, num_envs=E, seed=S, rng="numpy", auto_reset=True, keep_f64=True, obs_dtype=torch.uint8)
    obs = env.reset().cpu().numpy().reshape(E, 2, -1)
    assert (obs == ref["reset_obs"]).all(), "reset observations"
    assert (env.get_agents()[0] == ref["reset_pos"]).all(), "spawn positions"
    assert (env.get_maps() == ref["reset_maze"]).all(), "maps"
    is_nav = env.target_mode in ("Nav", "RPF")
    n_done = 0
    for t in range(T):
        o, r, d = env.step(torch.from_numpy(acts[t]).cuda())
        assert (o.cpu().numpy().reshape(E, 2, -1) == ref["obs"][t]).all(), ("obs", t)
        assert env.get_rewards_f64().tobytes() == ref["rew"][t].tobytes(), ("float64 rewards", t)
        assert r.cpu().numpy().tobytes() == ref["rew"][t].astype(np.float32).tobytes(), ("float32 rewards", t)
        assert (d.cpu().numpy() == ref["done"][t]).all(), ("done", t)
        pos, ctr = env.get_agents()
        assert (pos == ref["pos"][t]).all(), ("positions", t)
        assert (ctr == ref["ctr"][t]).all(), ("counters", t)
        if is_nav:
            assert (env.get_target_actions() == ref["tgt_act"][t]).all(), ("executed target action", t)
            _, ln, idx, _ = env.get_nav()
            assert (idx == ref["plan_meta"][t, :, 0]).all() and (ln == ref["plan_meta"][t, :, 1]).all(), ("plan cursor", t)
        n_done += int(ref["done"][t].sum())
    assert n_done == ref["n_done"] and n_done > E // 10, "the window must exercise auto-reset (%d finished)" % n_done
    if is_nav:  # plans of length < T were exhausted inside the window: replans happened
        assert (ref["plan_meta"][0, :, 1] < T - 2).sum() > 100
    for e in (0, 1, E // 2, E - 1):
        key, p = env.get_rng_numpy(e)
        assert p == int(ref["rng_ckpt"][e, 8]) and (key[:8] == ref["rng_ckpt"][e, :8]).all(), "RNG stream diverged for env %d" % e
    assert env.status() == 0
    assert env.counters()[0] == n_done
    env.close()


@pytest.mark.parametrize("env_id,E", [("Track2D-BlockPartialPZR-v0", 65536), ("Track2D-MazePartialAdv-v0", 32768)], ids=["PZR-65536", "MazeAdv-32768"])
def test_step_kernel_at_config_size_matches_oracle_on_philox_worlds(t2d, env_id, E):
    """the throughput configuration (Philox worlds, float32 observations, the contract kernel): read the device's own worlds
    back, replay T steps of every env in the oracle from that state, compare everything"""
    T = 12
    env = t2d.Track2DVecEnv(env_id, num_envs=E, seed=3, rng="philox", auto_reset=False, keep_f64=True)
    env.reset()
    maps = env.get_maps()
    pos0, ctr0 = env.get_agents()
    rs = np.random.RandomState(11)
    ctr0 = np.stack([rs.randint(0, 11, E), rs.randint(0, 499, E)], 1).astype(np.int32)  # incl. envs about to hit C_far > 10 / step 500
    ctr0[::7, 1] = 499
    env.set_agents(pos0, ctr0)
    acts = _actions(T, E, 13)
    ref = oracle.batch_inject_steps(env_id, maps, pos0, ctr0, acts)
    dead = np.zeros(E, bool)  # the oracle keeps stepping a finished env; so does the device without auto-reset
    for t in range(T):
        o, r, d = env.step(torch.from_numpy(acts[t]).cuda())
        o = o.cpu().numpy().reshape(E, 2, -1)
        assert o.dtype == np.float32 and (o == ref["obs"][t]).all(), ("obs", t)
        assert env.get_rewards_f64().tobytes() == ref["rew"][t].tobytes(), ("float64 rewards", t)
        assert r.cpu().numpy().tobytes() == ref["rew"][t].astype(np.float32).tobytes(), ("float32 rewards", t)
        assert (d.cpu().numpy() == ref["done"][t]).all(), ("done", t)
        pos, ctr = env.get_agents()
        assert (pos == ref["pos"][t]).all() and (ctr == ref["ctr"][t]).all(), ("state", t)
        dead |= ref["done"][t].astype(bool)
    assert dead.sum() > E // 20
    assert env.status() == 0
    env.close()


def test_cuda_astar_matches_reference_kats(t2d):
    """csrc/track2d_nav.cuh astar_plan called directly (track2d_astar_solve) on the 60 (maze, start, goal) -> plan known answers
    recorded from the reference's AstarSolver (oracle/refharness/make_golden.py), incl. the unreachable and start == goal cases"""
    g = np.load(os.path.join(GOLDEN, "astar_kat.npz"))
    n_unsolvable = n_empty = 0
    for dim, env_id in ((82, "Track2D-BlockPartialNav-v0"), (81, "Track2D-MazePartialNav-v0")):
        sel = np.nonzero(g["dim"] == dim)[0]
        if len(sel) == 0:
            continue
        env = t2d.Track2DVecEnv(env_id, num_envs=len(sel), seed=1, rng="philox", auto_reset=False)
        env.reset()
        env.set_maps(np.ascontiguousarray(g["maze"][sel][:, :dim, :dim]))
        plan, ln = env.astar_solve(g["start"][sel], g["goal"][sel])
        for k, i in enumerate(sel):
            L = int(g["length"][i])
            if L < 0:
                assert ln[k] == -1, (i, ln[k])
                n_unsolvable += 1
            else:
                assert ln[k] == L and list(plan[k][:L]) == list(g["plan"][i][:L]), i
                n_empty += L == 0
        assert not (env.status() & 4), "Frontier.replace fired"
        assert int(g["replaces"][sel].sum()) == 0
        env.close()
    assert n_unsolvable >= 1


@pytest.mark.parametrize("obs_dtype,n_chunks", [(torch.float32, 8), (torch.uint8, 3), (torch.float32, 16)])
def test_pipelined_host_step_equals_host_step(t2d, obs_dtype, n_chunks):
    """track2d_step_host_begin / track2d_host_chunk_wait (chunked, overlappable D2H) deliver exactly what track2d_step_host[_u8] does"""
    E = 1003
    a = t2d.Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=E, seed=9, rng="philox", auto_reset=True)
    b = t2d.Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=E, seed=9, rng="philox", auto_reset=True)
    ha, hb = a.alloc_host_buffers(obs_dtype=obs_dtype), b.alloc_host_buffers(obs_dtype=obs_dtype)
    a.reset_host(ha["obs"])
    b.reset_host(hb["obs"])
    assert torch.equal(ha["obs"], hb["obs"])
    rs = np.random.RandomState(2)
    for t in range(30):
        acts = torch.from_numpy(rs.randint(0, 4, size=(E, 2)).astype(np.int32))
        ha["actions"].copy_(acts)
        hb["actions"].copy_(acts)
        a.step_host(ha["actions"], ha["obs"], ha["reward"], ha["done"])
        hb["obs"].fill_(77)
        b.step_host_begin(hb["actions"], hb["obs"], hb["reward"], hb["done"], n_chunks)
        for c in range(n_chunks):
            b.host_chunk_wait(c)
            lo, hi = b.chunk_bounds(c, n_chunks)
            assert torch.equal(hb["obs"][lo:hi], ha["obs"][lo:hi]), (t, c)
            if c == 0:
                assert torch.equal(hb["reward"], ha["reward"]) and torch.equal(hb["done"], ha["done"])
    # the device-side form of the chunk dependency (track2d_host_chunk_wait_stream): the re-upload of chunk c is ordered after its D2H
    # by an event wait on the consumer's stream, the host never blocks in between
    dev = torch.empty(hb["obs"].shape, dtype=obs_dtype, device="cuda")
    for t in range(5):
        acts = torch.from_numpy(rs.randint(0, 4, size=(E, 2)).astype(np.int32))
        ha["actions"].copy_(acts)
        hb["actions"].copy_(acts)
        a.step_host(ha["actions"], ha["obs"], ha["reward"], ha["done"])
        hb["obs"].fill_(77)
        dev.fill_(55)
        torch.cuda.synchronize()
        b.step_host_begin(hb["actions"], hb["obs"], hb["reward"], hb["done"], n_chunks)
        for c in range(n_chunks):
            b.host_chunk_wait(c, on_stream=True)
            lo, hi = b.chunk_bounds(c, n_chunks)
            dev[lo:hi].copy_(hb["obs"][lo:hi], non_blocking=True)
        torch.cuda.synchronize()
        assert torch.equal(dev.cpu(), ha["obs"]) and torch.equal(hb["reward"], ha["reward"]) and torch.equal(hb["done"], ha["done"]), t
    assert a.status() == 0 and b.status() == 0
    a.close()
    b.close()


@pytest.mark.parametrize("env_id,E,T", [("Track2D-BlockPartialNav-v0", 384, 260), ("Track2D-MazePartialNav-v0", 200, 200),
                                        ("Track2D-MazePartialAdv-v0", 512, 120), ("Track2D-MazePartialRam-v0", 256, 120)])
def test_planning_ahead_equals_synchronous_planning(t2d, env_id, E, T):
    """T2D_FLAG_PLAN_AHEAD: Philox + auto-reset handles prepare next-episode worlds and Nav plans AHEAD of time on a side stream (standby
    worlds, plan ring) instead of inside step().  Counter-based RNG: WHEN something is computed must not change WHAT comes out --
    positions, executed target actions, rewards, dones and observations are identical step by step, across many resets and replans."""
    a = t2d.Track2DVecEnv(env_id, num_envs=E, seed=77, rng="philox", auto_reset=True, obs_dtype=torch.uint8, plan_ahead=True)
    b = t2d.Track2DVecEnv(env_id, num_envs=E, seed=77, rng="philox", auto_reset=True, obs_dtype=torch.uint8)
    oa, ob = a.reset().clone(), b.reset().clone()
    assert torch.equal(oa, ob) and (a.get_maps() == b.get_maps()).all() and (a.get_agents()[0] == b.get_agents()[0]).all()
    rs = np.random.RandomState(4)
    n_done = 0
    for t in range(T):
        pos, _ = b.get_agents()
        d = pos[:, 1] - pos[:, 0]  # most trackers chase their target (long episodes: plans run out), the rest wander off (resets)
        chase = np.where(np.abs(d[:, 0]) >= np.abs(d[:, 1]), np.where(d[:, 0] < 0, 0, 1), np.where(d[:, 1] < 0, 2, 3))
        a0 = np.where(np.arange(E) % 4 == 0, rs.randint(0, 4, E), chase)
        acts = torch.from_numpy(np.stack([a0, rs.randint(0, 4, E)], 1).astype(np.int32)).cuda()
        xa, ra, da = a.step(acts)
        xb, rb, db = b.step(acts)
        a.join()
        assert torch.equal(da, db), t
        assert torch.equal(ra, rb), t
        assert torch.equal(xa, xb), t
        pa, ca = a.get_agents()
        pb, cb = b.get_agents()
        assert (pa == pb).all() and (ca == cb).all(), t
        if a.target_mode in ("Nav", "Ram"):
            assert (a.get_target_actions() == b.get_target_actions()).all(), t
        n_done += int(da.sum())
    assert n_done > E // 4, n_done
    assert a.status() == 0 and b.status() == 0
    assert a.counters() == b.counters()
    a.close()
    b.close()
