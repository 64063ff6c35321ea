"""Record golden vectors from the UNMODIFIED reference env -- TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):

    python oracle/refharness/make_golden.py            # rewrites tests/golden/*.npz

The reference has no tests or golden vectors of its own, so these fixtures are
the pin for oracle/track2d_oracle.c and, through it, for the CUDA path.  Each fixture is a sequence of
episodes of one gym id driven with a scripted action stream:

  np.random.seed(seed)  ->  env.reset()  ->  step ... (done or cap)  ->  env.reset()  -> ...

with the global numpy RNG left running across the episodes of one seed, so the fixtures also pin the
ORDER and NUMBER of RNG draws (an MT19937 state checkpoint is stored after every reset and at the end
of every episode).  Actions come from a separate RandomState so they do not disturb the env's stream.
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "tests", "golden")

NAV_PAD = 400
RAM_PAD = 16

# (env id, seeds, episodes per seed, step cap per episode)
SPECS = [
    ("Track2D-BlockPartialPZR-v0", [1, 2, 3], 3, 60),
    ("Track2D-BlockPartialAdv-v0", [4], 3, 40),
    ("Track2D-BlockPartialFar-v0", [5], 3, 40),
    ("Track2D-BlockPartialRam-v0", [6, 7], 3, 80),
    ("Track2D-BlockPartialNav-v0", [8, 9], 3, 250),
    ("Track2D-BlockPartialRPF-v0", [10], 3, 200),
    ("Track2D-MazePartialAdv-v0", [11, 12], 3, 40),
    ("Track2D-MazePartialRam-v0", [13], 3, 60),
    ("Track2D-MazePartialNav-v0", [14], 3, 250),
    ("Track2D-BlockPartialPZR-v1", [15], 2, 40),
    ("Track2D-MazePartialPZR-v1", [16], 2, 40),
    ("Track2D-EmptyPartialRam-v0", [17], 2, 40),
    ("Track2D-BlockFullPZR-v0", [18], 2, 12),
    ("Track2D-MazeFullNav-v0", [19], 2, 12),
    ("Track2D-EmptyPartialAdv-v0", [20], 1, 600),  # TimeLimit case: both bump the top wall for 500 steps
]


def rng_checkpoint():
    st = np.random.get_state()
    return np.concatenate([st[1][:8].astype(np.uint32), np.asarray([st[2]], np.uint32)])


def action_for(policy, t, arng, state=None):
    if policy == 4:      # tracker chases the target (keeps Nav/RPF episodes alive past a replan)
        dr, dc = int(state[1][0]) - int(state[0][0]), int(state[1][1]) - int(state[0][1])
        if abs(dr) >= abs(dc) and dr != 0:
            a0 = 0 if dr < 0 else 1
        elif dc != 0:
            a0 = 2 if dc < 0 else 3
        else:
            a0 = int(arng.randint(4))
        if t % 7 == 6:
            a0 = int(arng.randint(4))
        return [a0, int(arng.randint(4))]
    if policy == 0:      # both uniform
        return [int(arng.randint(4)), int(arng.randint(4))]
    if policy == 1:      # tracker runs away upwards, target walks down: far-counter termination
        return [0, 1 if t % 3 else int(arng.randint(4))]
    if policy == 2:      # tracker uniform, target mostly right
        return [int(arng.randint(4)), 3 if t % 4 else int(arng.randint(4))]
    return [0, 0]        # policy 3: both push up forever (TimeLimit fixture)


def record(env_id, seeds, n_eps, cap):
    gym = ref.load_reference()
    env = gym.make(env_id)
    u = env.unwrapped
    full = 'Full' in env_id
    ep = dict(seed=[], policy=[], maze=[], gen_maze=[], init_state=[], goals=[], reset_obs=[], length=[],
              ram_plan=[], ram_len=[], nav_plan=[], nav_len=[], nav_goal=[], rng_after_reset=[], rng_after_episode=[])
    st = dict(actions=[], state=[], rewards=[], done=[], obs=[], c_far=[], tgt_i=[], tgt_len=[])
    for seed in seeds:
        arng = np.random.RandomState(1000 + seed)
        np.random.seed(seed)
        for k in range(n_eps):
            policy = 3 if cap > 500 else k % 3
            if ('Nav' in env_id or 'RPF' in env_id) and k == 0:
                policy = 4
            obs = env.reset()
            ep['seed'].append(seed if k == 0 else -1)
            ep['policy'].append(policy)
            ep['maze'].append(np.asarray(u.maze, np.uint8))
            ep['gen_maze'].append(np.asarray(u.maze_generator.get_maze(), np.uint8))
            ep['init_state'].append(np.asarray(u.state, np.int32))
            ep['goals'].append(np.asarray(u.goal_states, np.int32))
            ep['reset_obs'].append(np.asarray(obs, np.uint8).reshape(2, -1))
            rp, rl = np.full(RAM_PAD, -1, np.int32), 0
            npn, nl, ng = np.full(NAV_PAD, -1, np.int32), 0, np.zeros(2, np.int32)
            if 'Ram' in env_id:
                plan = np.asarray(u.Target[0].plan_actions, np.int32).reshape(-1)
                rl = len(plan)
                rp[:rl] = plan
            if 'Nav' in env_id or 'RPF' in env_id:
                plan = np.asarray(u.Target[0].plan_actions, np.int32).reshape(-1)
                nl = len(plan)
                assert nl <= NAV_PAD
                npn[:nl] = plan
                ng = np.asarray(u.Target[0].goal_states, np.int32)
            ep['ram_plan'].append(rp); ep['ram_len'].append(rl)
            ep['nav_plan'].append(npn); ep['nav_len'].append(nl); ep['nav_goal'].append(ng)
            ep['rng_after_reset'].append(rng_checkpoint())
            t = 0
            while True:
                a = action_for(policy, t, arng, u.state)
                obs, rew, done, info = env.step(a)
                t += 1
                st['actions'].append(a)
                st['state'].append(np.asarray(u.state, np.int32))
                st['rewards'].append(np.asarray(rew, np.float64))
                st['done'].append(bool(done))
                st['obs'].append(np.asarray(obs, np.uint8).reshape(2, -1))
                st['c_far'].append(int(u.C_far))
                if u.Target:
                    st['tgt_i'].append(int(u.Target[0].a_i)); st['tgt_len'].append(len(u.Target[0].plan_actions))
                else:
                    st['tgt_i'].append(0); st['tgt_len'].append(0)
                if done or t >= cap:
                    break
            ep['length'].append(t)
            ep['rng_after_episode'].append(rng_checkpoint())
    out = {('ep_' + k): np.asarray(v) for k, v in ep.items()}
    out.update({('st_' + k): np.asarray(v) for k, v in st.items()})
    out['st_actions'] = out['st_actions'].astype(np.int8)
    out['st_state'] = out['st_state'].astype(np.uint8)
    out['ep_init_state'] = out['ep_init_state'].astype(np.uint8)
    out['meta_full'] = np.asarray(full)
    return out


def record_astar(n_cases=60, seed=99):
    """Direct AstarSolver KATs incl. unsolvable goals and the (inverted) Frontier.replace branch."""
    ref.load_reference()
    from gym_track2d.envs import Astar_solver as A
    from gym_track2d.envs.generators import RandomBlockMazeGenerator, RandomMazeGenerator
    count = {'n': 0}
    orig = A.Frontier.replace

    def counting_replace(self, node):
        count['n'] += 1
        return orig(self, node)
    A.Frontier.replace = counting_replace
    rs = np.random.RandomState(seed)
    mazes, starts, goals, plans, lens, reps = [], [], [], [], [], []
    try:
        for i in range(n_cases):
            np.random.seed(5000 + i)
            if i % 3 == 2:
                gen = RandomMazeGenerator(width=80, height=80, complexity=0.03 * rs.rand(), density=0.03 * rs.rand())
            else:
                gen = RandomBlockMazeGenerator(maze_size=80, obstacle_ratio=[0.05, 0.15, 0.3, 0.45][i % 4] * (0.5 + 0.5 * rs.rand()))
            maze = np.asarray(gen.get_maze())
            free = np.argwhere(maze == 0)
            s = free[rs.randint(len(free))]
            g = free[rs.randint(len(free))] if i % 10 else s  # every 10th: start == goal (empty plan)
            count['n'] = 0
            solver = A.AstarSolver([int(s[0]), int(s[1])], [0, 1, 2, 3], maze, [int(g[0]), int(g[1])])
            acts = solver.get_actions() if solver.solvable() else None
            p = np.full(NAV_PAD, -1, np.int32)
            if acts is not None:
                assert len(acts) <= NAV_PAD
                p[:len(acts)] = acts
            mazes.append(np.pad(maze.astype(np.uint8), ((0, 82 - maze.shape[0]), (0, 82 - maze.shape[1]))))
            starts.append(s); goals.append(g); plans.append(p)
            lens.append(-1 if acts is None else len(acts)); reps.append(count['n'])
    finally:
        A.Frontier.replace = orig
    return dict(maze=np.asarray(mazes), dim=np.asarray([82 if i % 3 != 2 else 81 for i in range(n_cases)], np.int32),
                start=np.asarray(starts, np.int32), goal=np.asarray(goals, np.int32), plan=np.asarray(plans),
                length=np.asarray(lens, np.int32), replaces=np.asarray(reps, np.int32))


def main():
    os.makedirs(OUT, exist_ok=True)
    for env_id, seeds, n_eps, cap in SPECS:
        d = record(env_id, seeds, n_eps, cap)
        path = os.path.join(OUT, env_id.replace('Track2D-', 'episodes_') + '.npz')
        np.savez_compressed(path, **d)
        print('%-36s episodes=%d steps=%d dones=%d  %.1f KB' % (env_id, len(d['ep_length']), len(d['st_done']),
                                                              int(d['st_done'].sum()), os.path.getsize(path) / 1024))
    a = record_astar()
    path = os.path.join(OUT, 'astar_kat.npz')
    np.savez_compressed(path, **a)
    print('astar: cases=%d unsolvable=%d replaces_total=%d  %.1f KB' % (len(a['length']), int((a['length'] < 0).sum()),
                                                                      int(a['replaces'].sum()), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
