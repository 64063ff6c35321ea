"""Property tests (hypothesis) of the env semantics on the CPU oracle -- the invariants SURVEY 8(a) lists for
_next_state / _get_partial_obs / the reward and done rules, on arbitrary maps and positions."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle

MAZE = np.zeros((82, 82), np.uint8)
MAZE[0, :] = MAZE[-1, :] = MAZE[:, 0] = MAZE[:, -1] = 1


@st.composite
def world(draw):
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rs = np.random.RandomState(seed)
    m = MAZE.copy()
    k = draw(st.integers(0, 2500))
    idx = rs.choice(6400, k, replace=False)
    m[idx // 80 + 1, idx % 80 + 1] = 1
    free = np.argwhere(m == 0)
    a = free[rs.randint(len(free))]
    b = free[rs.randint(len(free))] if draw(st.booleans()) else a
    return m, [[int(a[0]), int(a[1])], [int(b[0]), int(b[1])]], [draw(st.integers(0, 3)), draw(st.integers(0, 3))], draw(st.sampled_from(["Adv", "PZR", "Far"]))


@settings(max_examples=150, deadline=None)
@given(world())
def test_transition_reward_and_observation_invariants(wd):
    m, pos, act, mode = wd
    env = oracle.OracleEnv(map_type="Block", target_mode=mode)
    env.set_state(m, pos, c_far=3, elapsed=7)
    obs, rew, done, _ = env.step(act)
    st_, _, cfar, elapsed = env.state()
    D = {0: (-1, 0), 1: (1, 0), 2: (0, -1), 3: (0, 1)}
    for i in range(2):  # _next_state: move unless the destination is a wall; agents never stand on walls; overlap allowed
        want = (pos[i][0] + D[act[i]][0], pos[i][1] + D[act[i]][1])
        assert tuple(st_[i]) == (want if m[want] == 0 else tuple(pos[i]))
        assert m[tuple(st_[i])] == 0
    d = float(np.sqrt(float((st_[1][0] - st_[0][0]) ** 2 + (st_[1][1] - st_[0][1]) ** 2)))
    w_p = {"Adv": 0.0, "PZR": 1.0, "Far": -0.5}[mode]
    r_track = max(1 - 2 * d / 6.0, -1)
    r_target = max(-r_track - w_p * max(d - 6.0, 0) / 6.0, -1)
    assert rew[0] == r_track and rew[1] == r_target and -1 <= rew[0] <= 1 and rew[1] >= -1  # Far (w_p = -0.5) has no upper bound
    assert cfar == (0 if d <= 6 else 4) and elapsed == 8 and done is False
    # observation: centre colours, value set, walls/out-of-map, visibility rule (no line-of-sight occlusion)
    o = obs[:, 0]
    assert o[0, 6, 6] == 2 and o[1, 6, 6] == 4 and set(np.unique(o)) <= {0, 1, 2, 4}
    for i in range(2):
        r, c = st_[i]
        for wr in range(13):
            for wc in range(13):
                rr, cc = r - 6 + wr, c - 6 + wc
                v = o[i, wr, wc]
                inside = 0 <= rr < 82 and 0 <= cc < 82
                if (wr, wc) == (6, 6):
                    continue
                if inside and (rr, cc) == tuple(st_[1 - i]):
                    assert v == (4 if i == 0 else 2)
                else:
                    assert v == (m[rr, cc] if inside else 1)
    dr, dc = st_[1][0] - st_[0][0], st_[1][1] - st_[0][1]
    visible = abs(dr) <= 6 and abs(dc) <= 6 and (dr, dc) != (0, 0)
    assert (o[0] == 4).sum() == (1 if visible else 0) and (o[1] == 2).sum() == (1 if visible else 0)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 32 - 1), st.sampled_from(["Block", "Maze", "Empty"]), st.sampled_from(["Adv", "Ram", "Nav", "RPF"]))
def test_reset_invariants_for_any_seed(seed, map_type, target):
    env = oracle.OracleEnv(map_type=map_type, target_mode=target)
    env.seed(seed)
    obs = env.reset()
    m = env.maze()
    st_, goals, cfar, elapsed = env.state()
    H = 81 if map_type == "Maze" else 82
    assert m.shape == (H, H) and (m[0] == 1).all() and (m[-1] == 1).all() and (m[:, 0] == 1).all() and (m[:, -1] == 1).all()
    if map_type == "Block":
        assert m[1:-1, 1:-1].sum() <= 959
    if map_type == "Empty":
        assert m[1:-1, 1:-1].sum() == 0
    gm = env.gen_maze()
    assert gm[tuple(st_[0])] == 0 and gm[tuple(st_[1])] == 0
    d = st_[0] - st_[1]
    assert ((d >= 0) & (d <= 1)).all()  # target in {r-1, r} x {c-1, c}
    assert (cfar, elapsed) == (0, 0) and obs.shape == (2, 1, 13, 13)
    if target == "Ram":
        plan, idx = env.ram()
        assert 1 <= len(plan) <= 9 and idx == 0 and set(plan) <= {0, 1, 2, 3}
    if target in ("Nav", "RPF"):
        plan, idx, goal = env.nav()
        assert idx == 0 and len(plan) >= 1
        r, c = st_[1]
        D = {0: (-1, 0), 1: (1, 0), 2: (0, -1), 3: (0, 1)}
        ok = True
        for a in plan:
            r, c = r + D[int(a)][0], c + D[int(a)][1]
            ok = ok and gm[r, c] == 0
        if len(plan) != 10:  # a genuine A* plan is wall-free and ends on the goal (plan B = 10 random actions need not)
            assert ok and (r, c) == tuple(goal)
