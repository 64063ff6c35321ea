"""GPU parity tests: the CUDA path (through the libtrack2d C ABI) against the CPU oracle and against the
golden vectors recorded from the unmodified reference.  Integer / byte / index results are compared
bit-exactly; float64 rewards bit-exactly; float32 rewards must equal the float32 rounding of the
oracle's float64 (what player_util.py:58 does)."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

EPISODE_FILES = sorted(glob.glob(os.path.join(GOLDEN, "episodes_*.npz")))


@pytest.fixture(scope="module")
def t2d():
    from active_tracking_rl_b200 import envs
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return envs


def _ckpt(key, pos):
    return np.concatenate([key[:8], np.asarray([pos], np.uint32)])


# ------------------------------------------------------------------------------------------------
# 1. golden replay: the single-env gym shim in numpy-RNG mode reproduces the reference bit for bit
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", EPISODE_FILES, ids=[os.path.basename(p)[9:-4] for p in EPISODE_FILES])
def test_cuda_replays_reference_episodes(t2d, path):
    g = np.load(path)
    env_id = "Track2D-" + os.path.basename(path)[9:-4]
    env = t2d.make(env_id, seed=0, rng="numpy")
    is_ram, is_nav = "Ram" in env_id, ("Nav" in env_id or "RPF" in env_id)
    t0 = 0
    for k in range(len(g["ep_length"])):
        if g["ep_seed"][k] >= 0:
            env.seed(int(g["ep_seed"][k]))
        obs = env.reset()
        assert obs.shape == (2, 1) + ((13, 13) if "Partial" in env_id else env.maze.shape)
        assert (env.maze == g["ep_maze"][k]).all(), "maze, episode %d" % k
        assert (np.asarray(env.state) == g["ep_init_state"][k]).all()
        if not is_nav or "RPF" not in env_id:
            assert (np.asarray(env.goal_states) == g["ep_goals"][k]).all()
        assert (obs.reshape(2, -1) == g["ep_reset_obs"][k]).all()
        if is_ram:
            plan, ln, idx = env.vec.get_ram()
            assert list(plan[0][: ln[0]]) == list(g["ep_ram_plan"][k][: g["ep_ram_len"][k]]) and idx[0] == 0
        if is_nav:
            plan, ln, idx, goal = env.vec.get_nav()
            assert ln[0] == g["ep_nav_len"][k] and list(plan[0][: ln[0]]) == list(g["ep_nav_plan"][k][: ln[0]])
            assert (goal[0] == g["ep_nav_goal"][k]).all()
        assert (_ckpt(*env.vec.get_rng_numpy(0)) == g["ep_rng_after_reset"][k]).all(), "RNG stream after reset"
        L = int(g["ep_length"][k])
        for t in range(t0, t0 + L):
            obs, rew, done, info = env.step(g["st_actions"][t])
            assert (np.asarray(env.state) == g["st_state"][t]).all(), t
            assert rew.dtype == np.float64 and rew.tobytes() == g["st_rewards"][t].tobytes(), (t, rew, g["st_rewards"][t])
            assert done == bool(g["st_done"][t]), t
            assert env.C_far == int(g["st_c_far"][t])
            assert (obs.reshape(2, -1) == g["st_obs"][t]).all(), t
            if is_nav:
                _, ln, idx, _ = env.vec.get_nav()
                assert (idx[0], ln[0]) == (int(g["st_tgt_i"][t]), int(g["st_tgt_len"][t])), t
            if is_ram:
                _, ln, idx = env.vec.get_ram()
                assert (idx[0], ln[0]) == (int(g["st_tgt_i"][t]), int(g["st_tgt_len"][t])), t
        t0 += L
        assert (_ckpt(*env.vec.get_rng_numpy(0)) == g["ep_rng_after_episode"][k]).all(), "RNG stream after episode %d" % k
    assert env.vec.status() == 0
    env.close()


# ------------------------------------------------------------------------------------------------
# 2. batched transition parity on injected states (the hot kernel), vs the oracle
# ------------------------------------------------------------------------------------------------
def _oracle_batch(env_id, E, seed):
    envs = []
    for e in range(E):
        o = oracle.OracleEnv(env_id)
        o.seed(seed + e)
        envs.append(o)
    return envs


@pytest.mark.parametrize("env_id,E", [("Track2D-BlockPartialPZR-v0", 1000), ("Track2D-BlockPartialAdv-v0", 130),
                                      ("Track2D-BlockPartialFar-v0", 257), ("Track2D-MazePartialPZR-v0", 300),
                                      ("Track2D-EmptyPartialAdv-v0", 33)])
def test_step_kernel_matches_oracle_on_injected_states(t2d, env_id, E):
    rs = np.random.RandomState(7)
    orc = _oracle_batch(env_id, E, 100)
    for o in orc:
        o.reset()
    env = t2d.Track2DVecEnv(env_id, num_envs=E, seed=5, rng="philox", auto_reset=False, keep_f64=True)
    env.reset()
    maps = np.stack([o.maze() for o in orc])
    pos = np.zeros((E, 2, 2), np.int32)
    ctr = np.zeros((E, 2), np.int32)
    for e, o in enumerate(orc):
        free = np.argwhere(maps[e] == 0)
        a = free[rs.randint(len(free))]
        # half the envs: target close to the tracker (inside the FOV, incl. overlap); rest: anywhere
        if e % 2 == 0:
            near = free[(np.abs(free - a).max(1) <= 7)]
            b = near[rs.randint(len(near))]
        else:
            b = free[rs.randint(len(free))]
        pos[e] = [a, b]
        ctr[e] = [rs.randint(0, 11), rs.randint(0, 499)]
        o.set_state(maps[e], pos[e], c_far=int(ctr[e][0]), elapsed=int(ctr[e][1]))
    env.set_maps(maps)
    env.set_agents(pos, ctr)
    for t in range(40):
        acts = rs.randint(0, 4, size=(E, 2)).astype(np.int32)
        obs, rew, done = env.step(torch.from_numpy(acts).cuda())
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        rew64 = env.get_rewards_f64()
        gpos, gctr = env.get_agents()
        for e, o in enumerate(orc):
            oo, orew, odone, _ = o.step(acts[e])
            st, _, cfar, elapsed = o.state()
            assert (gpos[e] == st).all(), (t, e)
            assert (gctr[e] == [cfar, elapsed]).all(), (t, e)
            assert rew64[e].tobytes() == orew.tobytes(), (t, e, rew64[e], orew)
            assert rew[e].tobytes() == orew.astype(np.float32).tobytes(), (t, e)
            assert bool(done[e]) == odone, (t, e)
            assert (obs[e].reshape(2, 1, 13, 13) == oo).all(), (t, e)
    assert env.status() == 0
    env.close()


# ------------------------------------------------------------------------------------------------
# 3. whole pipeline in numpy-RNG mode, batched, with auto-reset: env e == reference seeded (S + e)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env_id,E,T", [("Track2D-BlockPartialPZR-v0", 96, 120), ("Track2D-BlockPartialRam-v0", 64, 120),
                                        ("Track2D-MazePartialAdv-v0", 64, 100), ("Track2D-MazePartialRam-v0", 32, 60),
                                        ("Track2D-BlockPartialNav-v0", 24, 80), ("Track2D-BlockPartialRPF-v0", 16, 60),
                                        ("Track2D-BlockPartialPZR-v1", 32, 60), ("Track2D-EmptyPartialRam-v0", 16, 40)])
def test_numpy_mode_pipeline_matches_oracle(t2d, env_id, E, T):
    S = 4242
    rs = np.random.RandomState(3)
    orc = _oracle_batch(env_id, E, S)
    env = t2d.Track2DVecEnv(env_id, num_envs=E, seed=S, rng="numpy", auto_reset=True, keep_f64=True)
    obs = env.reset().cpu().numpy()
    for e, o in enumerate(orc):
        assert (obs[e] == o.reset()).all(), e
    assert (env.get_maps() == np.stack([o.maze() for o in orc])).all()
    n_done = 0
    for t in range(T):
        acts = rs.randint(0, 4, size=(E, 2)).astype(np.int32)
        if t % 3 == 0:
            acts[:, 0] = 0  # bias the tracker away so episodes end (far counter) and auto-reset is exercised
        obs, rew, done = env.step(torch.from_numpy(acts).cuda())
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        rew64 = env.get_rewards_f64()
        tgt = env.get_target_actions()
        for e, o in enumerate(orc):
            oo, orew, odone, info = o.step(acts[e])
            assert rew64[e].tobytes() == orew.tobytes(), (t, e)
            assert bool(done[e]) == odone, (t, e)
            if env.target_mode in ("Ram", "Nav", "RPF"):
                assert tgt[e] == info["target_action"], (t, e)
            if odone:
                oo = o.reset()
                n_done += 1
            assert (obs[e] == oo).all(), (t, e)
        gpos, gctr = env.get_agents()
        for e, o in enumerate(orc):
            st, _, cfar, elapsed = o.state()
            assert (gpos[e] == st).all() and (gctr[e] == [cfar, elapsed]).all(), (t, e)
    assert n_done > 0, "no episode finished; the test did not exercise auto-reset"
    for e in (0, E - 1):
        key, pos = env.get_rng_numpy(e)
        okey, opos = orc[e].rng_state()
        assert pos == opos and (key == okey).all(), "RNG stream diverged for env %d" % e
    assert env.status() == 0
    assert env.counters()[0] == n_done
    env.close()


# ------------------------------------------------------------------------------------------------
# 4. Philox mode: same distributions as the reference's generators (checked against the oracle)
# ------------------------------------------------------------------------------------------------
def test_philox_block_reset_invariants_and_distributions(t2d):
    E = 8192
    env = t2d.Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=E, seed=11, rng="philox", auto_reset=False)
    obs = env.reset().cpu().numpy()
    maps = env.get_maps()
    pos, ctr = env.get_agents()
    assert (maps[:, 0, :] == 1).all() and (maps[:, -1, :] == 1).all() and (maps[:, :, 0] == 1).all() and (maps[:, :, -1] == 1).all()
    k = maps[:, 1:-1, 1:-1].reshape(E, -1).sum(1)
    assert k.min() >= 0 and k.max() <= 959  # k = int(0.15 * u * 6400)
    # k / 6400 ~ U[0, 0.15): mean 479.5, KS-ish bound on the empirical CDF
    ks = np.abs(np.sort(k) / 960.0 - (np.arange(E) + 0.5) / E).max()
    assert ks < 0.03, ks
    assert (ctr == 0).all()
    er = np.arange(E)
    assert (maps[er, pos[:, 0, 0], pos[:, 0, 1]] == 0).all() and (maps[er, pos[:, 1, 0], pos[:, 1, 1]] == 0).all()
    d = pos[:, 0] - pos[:, 1]  # target in {r-1, r} x {c-1, c}
    assert ((d >= 0) & (d <= 1)).all()
    # tracker uniform over the interior: row/col means near 40.5, spread near uniform on 1..80
    assert abs(pos[:, 0, 0].mean() - 40.5) < 1.0 and abs(pos[:, 0, 1].mean() - 40.5) < 1.0
    assert abs(pos[:, 0, 0].std() - np.sqrt((80 ** 2 - 1) / 12.0)) < 0.7
    # obstacle cells uniform over the interior: per-cell occupancy ~ mean density
    occ = maps[:, 1:-1, 1:-1].mean(0)
    assert abs(occ.mean() - 479.5 / 6400) < 0.002 and occ.std() < 0.006
    # target offset law: among envs whose four 2x2 cells are all free the offset is uniform over the 4
    allfree = np.array([maps[e, pos[e, 0, 0] - 1:pos[e, 0, 0] + 1, pos[e, 0, 1] - 1:pos[e, 0, 1] + 1].sum() == 0 for e in range(E)])
    hist = np.bincount((d[allfree][:, 0] * 2 + d[allfree][:, 1]), minlength=4) / allfree.sum()
    assert np.abs(hist - 0.25).max() < 0.03, hist
    # reset observation: centre cells and value set
    assert (obs[:, 0, 0, 6, 6] == 2).all() and (obs[:, 1, 0, 6, 6] == 4).all()
    assert set(np.unique(obs)) <= {0.0, 1.0, 2.0, 4.0}
    # and it equals what the oracle computes for the same injected state
    o = oracle.OracleEnv("Track2D-BlockPartialPZR-v0")
    for e in range(0, E, 257):
        o.set_state(maps[e], pos[e])
        assert (o.obs() == obs[e]).all(), e
    # determinism: same seed, same world; different seed, different world
    env2 = t2d.Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=E, seed=11, rng="philox", auto_reset=False)
    env2.reset()
    assert (env2.get_maps() == maps).all() and (env2.get_agents()[0] == pos).all()
    env3 = t2d.Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=64, seed=12, rng="philox", auto_reset=False)
    env3.reset()
    assert (env3.get_maps() != maps[:64]).any()
    for x in (env, env2, env3):
        x.close()


def test_philox_maze_and_ram_against_oracle_statistics(t2d):
    E = 2048
    env = t2d.Track2DVecEnv("Track2D-MazePartialRam-v0", num_envs=E, seed=3, rng="philox", auto_reset=False)
    env.reset()
    maps = env.get_maps()
    assert maps.shape[1:] == (81, 81)
    assert (maps[:, 0, :] == 1).all() and (maps[:, -1, :] == 1).all() and (maps[:, :, 0] == 1).all() and (maps[:, :, -1] == 1).all()
    walls = maps[:, 1:-1, 1:-1].reshape(E, -1).sum(1)
    ow = []
    for s in range(300):
        o = oracle.OracleEnv("Track2D-MazePartialRam-v0")
        o.seed(900 + s)
        o.reset()
        ow.append(o.maze()[1:-1, 1:-1].sum())
    assert abs(walls.mean() - np.mean(ow)) < 4 * np.std(ow) / np.sqrt(300) + 3, (walls.mean(), np.mean(ow))
    plan, ln, idx = env.get_ram()
    assert ln.min() >= 1 and ln.max() <= 9 and (idx == 0).all()
    assert np.abs(np.bincount(ln, minlength=10)[1:] / E - 1 / 9.0).max() < 0.03
    a0 = plan[:, 0]
    assert np.abs(np.bincount(a0, minlength=4) / E - 0.25).max() < 0.04
    # run: the executed target actions follow the stored plans
    acts = torch.zeros((E, 2), dtype=torch.int32, device="cuda")
    for t in range(30):
        plan, ln, idx = env.get_ram()
        env.step(acts)
        tgt = env.get_target_actions()
        cont = idx + 1 < ln  # plan not exhausted by this step: the action is the planned one
        assert (tgt[cont] == plan[np.arange(E), idx][cont]).all()
    env.close()


# ------------------------------------------------------------------------------------------------
# 5. size-independent properties at the BASELINE sizes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("E", [4096, 65536])
def test_full_size_properties(t2d, E):
    env = t2d.Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=E, seed=1, rng="philox", auto_reset=True)
    obs = env.reset().clone()
    g = torch.Generator(device="cuda").manual_seed(0)
    ep0 = 0
    for t in range(60):
        acts = torch.randint(0, 4, (E, 2), generator=g, device="cuda", dtype=torch.int32)
        ppos, pctr = env.get_agents()
        obs, rew, done = env.step(acts)
        pos, ctr = env.get_agents()
        maps = env.get_maps() if t % 20 == 0 else None
        o, r, d = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy().astype(bool)
        # finished envs were reset in the same call: counters zero, fresh spawn adjacent
        assert (ctr[d] == 0).all()
        live = ~d
        # moves are unit steps or stays
        assert (np.abs(pos[live] - ppos[live]).sum(-1) <= 1).all()
        # rewards recomputed in float64 from the positions, rounded once to float32
        a = acts.cpu().numpy()
        dd = np.sqrt(((pos[live, 1] - pos[live, 0]).astype(np.float64) ** 2).sum(1))
        rt = np.maximum(1 - 2 * dd / 6.0, -1)
        rg = np.maximum(-rt - 1.0 * np.maximum(dd - 6.0, 0) / 6.0, -1)
        assert (r[live, 0] == rt.astype(np.float32)).all() and (r[live, 1] == rg.astype(np.float32)).all()
        # far counter
        far = dd > 6
        assert (ctr[live, 0][~far] == 0).all() and (ctr[live, 0][far] == pctr[live, 0][far] + 1).all()
        assert (ctr[live, 1] == pctr[live, 1] + 1).all()
        # observation structure
        assert (o[:, 0, 0, 6, 6] == 2).all() and (o[:, 1, 0, 6, 6] == 4).all()
        vis = (np.abs(pos[:, 1] - pos[:, 0]).max(1) <= 6) & ((pos[:, 1] != pos[:, 0]).any(1))
        ee = np.arange(E)[vis]
        dr, dc = (pos[vis, 1, 0] - pos[vis, 0, 0]), (pos[vis, 1, 1] - pos[vis, 0, 1])
        assert (o[ee, 0, 0, 6 + dr, 6 + dc] == 4).all() and (o[ee, 1, 0, 6 - dr, 6 - dc] == 2).all()
        assert ((o == 4).reshape(E, -1).sum(1) == 1 + vis).all() and ((o == 2).reshape(E, -1).sum(1) == 1 + vis).all()
        if maps is not None:
            er = np.arange(E)
            assert (maps[er, pos[:, 0, 0], pos[:, 0, 1]] == 0).all() and (maps[er, pos[:, 1, 0], pos[:, 1, 1]] == 0).all()
            # the tracker's window equals the map slice (walls) wherever it is inside the map
            for e in range(0, E, max(1, E // 64)):
                r0, c0 = pos[e, 0]
                win = np.ones((13, 13), np.uint8)
                rr0, rr1, cc0, cc1 = max(r0 - 6, 0), min(r0 + 7, 82), max(c0 - 6, 0), min(c0 + 7, 82)
                win[rr0 - (r0 - 6):rr1 - (r0 - 6), cc0 - (c0 - 6):cc1 - (c0 - 6)] = maps[e, rr0:rr1, cc0:cc1]
                got = o[e, 0, 0].copy()
                got[got > 1] = 0
                assert (got == win).all(), e
        ep0 += int(d.sum())
    assert ep0 > 0
    assert env.counters() == (ep0, 60 * E)
    assert env.status() == 0
    env.close()


def test_uint8_observations_equal_float32(t2d):
    E = 1030
    a = t2d.Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=E, seed=9, rng="philox", auto_reset=True)
    b = t2d.Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=E, seed=9, rng="philox", auto_reset=True, obs_dtype=torch.uint8)
    assert (a.reset().cpu().numpy() == b.reset().cpu().numpy()).all()
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in range(50):
        acts = torch.randint(0, 4, (E, 2), generator=g, device="cuda", dtype=torch.int32)
        oa, ra, da = a.step(acts)
        ob, rb, db = b.step(acts)
        assert ob.dtype == torch.uint8
        assert (oa.cpu().numpy() == ob.cpu().numpy()).all() and torch.equal(ra, rb) and torch.equal(da, db)
    a.close()
    b.close()


@pytest.mark.parametrize("host_dtype", [torch.float32, torch.uint8])
def test_host_buffer_api_equals_device_api(t2d, host_dtype):
    E = 515
    a = t2d.Track2DVecEnv("Track2D-BlockPartialRam-v0", num_envs=E, seed=2, rng="philox", auto_reset=True)
    b = t2d.Track2DVecEnv("Track2D-BlockPartialRam-v0", num_envs=E, seed=2, rng="philox", auto_reset=True)
    hb = b.alloc_host_buffers(obs_dtype=host_dtype)
    assert hb["obs"].dtype == host_dtype
    oa = a.reset()
    b.reset_host(hb["obs"])
    assert (oa.cpu() == hb["obs"]).all()
    rs = np.random.RandomState(0)
    for t in range(30):
        acts = rs.randint(0, 4, size=(E, 2)).astype(np.int32)
        oa, ra, da = a.step(torch.from_numpy(acts).cuda())
        hb["actions"].copy_(torch.from_numpy(acts))
        b.step_host(hb["actions"], hb["obs"], hb["reward"], hb["done"])
        assert (oa.cpu() == hb["obs"]).all() and (ra.cpu() == hb["reward"]).all() and (da.cpu() == hb["done"]).all()
    a.close()
    b.close()


def test_error_behaviour(t2d):
    from active_tracking_rl_b200 import _lib
    with pytest.raises(TypeError):  # track_1v1.py:261
        t2d.Track2DVecEnv(map_type="Block", obs_type="Sideways", target_mode="PZR", num_envs=2)
    with pytest.raises(KeyError):
        t2d.make("Track2D-BlockPartialNope-v0")
    env = t2d.Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=8, seed=1)
    acts = torch.zeros((8, 2), dtype=torch.int32, device="cuda")
    with pytest.raises(_lib.Track2DError):  # gym TimeLimit: "Cannot call env.step() before calling reset()"
        env.step(acts)
    env.reset()
    with pytest.raises(TypeError):
        env.step(torch.zeros((8, 3), dtype=torch.int32, device="cuda"))
    env.step(torch.full((8, 2), 9, dtype=torch.int32, device="cuda"))
    assert env.status() & _lib.STATUS_BAD_ACTION
    env.close()


def test_sharedadam_kernel_matches_reference_formula(t2d):
    import ctypes as C
    from active_tracking_rl_b200 import _lib
    lib = _lib.load()
    n = 801291
    g = torch.Generator(device="cuda").manual_seed(5)
    p = torch.randn(n, device="cuda", generator=g)
    m, v, vmax = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    pr, mr, vr, xr = p.clone().double(), m.clone().double(), v.clone().double(), vmax.clone().double()
    scratch = torch.zeros(2, device="cuda")
    lr, b1, b2, eps, maxn = 1e-3, 0.9, 0.999, 1e-3, 50.0
    for step in range(1, 6):
        grad = torch.randn(n, device="cuda", generator=g) * (0.2 if step % 2 else 0.01)
        _lib.check(lib.track2d_sharedadam_step(C.c_void_p(p.data_ptr()), C.c_void_p(grad.data_ptr()), C.c_void_p(m.data_ptr()),
                                               C.c_void_p(v.data_ptr()), C.c_void_p(vmax.data_ptr()), n, step, lr, b1, b2, eps, maxn, 1.0,
                                               C.c_void_p(scratch.data_ptr()), None, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        # shared_optim.py:122-175 + clip_grad_norm_, in float64
        gd = grad.double()
        total = gd.norm()
        gd = gd * min(1.0, maxn / (total.item() + 1e-6))
        mr = mr * b1 + (1 - b1) * gd
        vr = vr * b2 + (1 - b2) * gd * gd
        xr = torch.maximum(xr, vr)
        step_size = lr * np.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        pr = pr - step_size * mr / (xr.sqrt() + eps)
        assert torch.allclose(p.double(), pr, rtol=0, atol=5e-6), (step, (p.double() - pr).abs().max())  # fp32 state vs fp64 restatement
        assert torch.allclose(vmax.double(), xr, rtol=1e-5, atol=1e-9)


def _bfs_dist(maze, start, goal):
    from collections import deque
    H, W = maze.shape
    dist = -np.ones((H, W), np.int32)
    dist[start[0], start[1]] = 0
    q = deque([tuple(start)])
    while q:
        r, c = q.popleft()
        if (r, c) == tuple(goal):
            return int(dist[r, c])
        for dr, dc in ((-1, 0), (1, 0), (0, -1), (0, 1)):
            nr, nc = r + dr, c + dc
            if maze[nr, nc] == 0 and dist[nr, nc] < 0:
                dist[nr, nc] = dist[r, c] + 1
                q.append((nr, nc))
    return -1


@pytest.mark.parametrize("env_id", ["Track2D-BlockPartialNav-v0", "Track2D-MazePartialNav-v0"])
def test_philox_navigator_plans_are_valid_shortest_paths(t2d, env_id):
    """Philox mode has no reference stream to match; the Navigator's contract is checked instead: every plan is a
    wall-free path from the target to its goal of BFS-optimal length (A* with an admissible heuristic), and the
    target executes it action by action, replanning when it runs out (navigator.py:11-36)."""
    E = 96
    env = t2d.Track2DVecEnv(env_id, num_envs=E, seed=21, rng="philox", auto_reset=True)
    env.reset()
    D = {0: (-1, 0), 1: (1, 0), 2: (0, -1), 3: (0, 1)}

    def check_plans(envs):
        maps, (pos, _), (plan, ln, idx, goal) = env.get_maps(), env.get_agents(), env.get_nav()
        for e in envs:
            if ln[e] == 10 and _bfs_dist(maps[e], pos[e, 1], goal[e]) != 10:
                continue  # plan B (10 random actions) after 6 failed goals: astronomically rare, but legal
            r, c = pos[e, 1]
            for a in plan[e, idx[e]:ln[e]]:
                r, c = r + D[int(a)][0], c + D[int(a)][1]
                assert maps[e, r, c] == 0, "plan walks into a wall"
            assert (r, c) == tuple(goal[e])
            if idx[e] == 0:
                assert ln[e] == _bfs_dist(maps[e], pos[e, 1], goal[e]), "A* plan is not a shortest path"
    check_plans(range(E))
    g = torch.Generator(device="cuda").manual_seed(3)
    replans = 0
    for t in range(120):
        pos, _ = env.get_agents()
        d = pos[:, 1] - pos[:, 0]  # tracker chases the target so episodes live long enough to exhaust plans
        a0 = np.where(np.abs(d[:, 0]) >= np.abs(d[:, 1]), np.where(d[:, 0] < 0, 0, 1), np.where(d[:, 1] < 0, 2, 3))
        acts = torch.from_numpy(np.stack([a0, np.zeros(E, np.int64)], 1).astype(np.int32)).cuda()
        plan, ln, idx, goal = env.get_nav()
        env.step(acts)
        tgt = env.get_target_actions()
        cont = idx < ln  # no (re)plan needed before this step: the executed action is the planned one (ln == 0: first plan of an episode an
        #                  auto-reset started -- it is made by the replan wave of the next step)
        assert (tgt[cont] == plan[np.arange(E), np.minimum(idx, 1023)][cont]).all()
        replans += int((~cont).sum())
        if t % 40 == 39:
            _, ln2, idx2, _ = env.get_nav()
            check_plans(np.nonzero(idx2 < ln2)[0][:24])
    assert replans > 0, "no plan ran out: the replan kernel was not exercised"
    assert env.status() == 0
    env.close()


@pytest.mark.parametrize("env_id,E", [("Track2D-BlockFullPZR-v0", 37), ("Track2D-MazeFullRam-v0", 18), ("Track2D-EmptyFullFar-v1", 9)])
def test_full_observation_batched_matches_oracle(t2d, env_id, E):
    """obs_type 'Full' (track_1v1.py:288-290): both agents see the whole map, tracker = 2, then target = 4."""
    S = 77
    orc = _oracle_batch(env_id, E, S)
    env = t2d.Track2DVecEnv(env_id, num_envs=E, seed=S, rng="numpy", auto_reset=True, keep_f64=True)
    obs = env.reset().cpu().numpy()
    H, W = env.H, env.W
    assert obs.shape == (E, 2, 1, H, W) and env.observation_space[0].shape == (1, H, W)
    for e, o in enumerate(orc):
        assert (obs[e] == o.reset()).all(), e
    rs = np.random.RandomState(5)
    for t in range(40):
        acts = rs.randint(0, 4, size=(E, 2)).astype(np.int32)
        acts[:, 0] = 0 if t % 2 else acts[:, 0]
        obs, rew, done = env.step(torch.from_numpy(acts).cuda())
        obs, done = obs.cpu().numpy(), done.cpu().numpy()
        r64 = env.get_rewards_f64()
        for e, o in enumerate(orc):
            oo, orew, odone, _ = o.step(acts[e])
            assert r64[e].tobytes() == orew.tobytes() and bool(done[e]) == odone
            if odone:
                oo = o.reset()
            assert (obs[e] == oo).all(), (t, e)
    env.close()


def test_gym_shim_surface_matches_reference_types(t2d):
    """the drop-in surface of SURVEY 8(b): spaces are LISTS of two, reset/step return the reference's types and dtypes"""
    env = t2d.make("Track2D-BlockPartialPZR-v0", seed=3)
    assert isinstance(env.observation_space, list) and len(env.observation_space) == 2 and isinstance(env.action_space, list)
    assert env.observation_space[0].shape == (1, 13, 13) and env.action_space[1].n == 4
    assert env.seed(5) == [5]
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (2, 1, 13, 13) and obs.dtype == np.float64  # Block maps are float (generators.py:161)
    out = env.step([np.array([1]), np.int64(2)])  # 1-element arrays and numpy ints are coerced with int() (track_1v1.py:87)
    assert len(out) == 4
    obs, rew, done, info = out
    assert rew.dtype == np.float64 and rew.shape == (2,) and isinstance(done, bool) and "distance" in info and "traces" in info
    assert env.unwrapped is env and len(env.state) == 2 and env.maze.shape == (82, 82)
    with pytest.raises(NotImplementedError):
        env.render()
    env.close()
    menv = t2d.make("Track2D-MazePartialAdv-v0", seed=3)
    assert menv.reset().dtype == np.int64 and menv.maze.shape == (81, 81)  # Maze maps are int (generators.py:145)
    menv.close()
