"""Task-quality parity on the SAME checkpoint: tests/golden/gym_eval_reference.npz holds what the UNMODIFIED reference
gym_eval.py measured, episode by episode, on tests/golden/tracker_b200_e4096.dat -- a tracker trained by this repo's batched learner and
loaded by the reference through its own --load-tracker path (oracle/refharness/make_golden_eval.py).  With the global numpy RNG seeded the
evaluation is a deterministic function of the checkpoint, so both restatements must reproduce it exactly:

  * CPU (not gpu): the C oracle env + the learner oracle's forward (greedy, model.py:45-46) -> same episode lengths and returns;
  * GPU: the single-env gym shim (numpy-RNG mode) + the batched model in test mode through player_util's action_test arithmetic.

Episode lengths are compared exactly and returns to 1e-9 (sums of the same float64 rewards in the same order; the policy enters only
through argmax).  The README-protocol statistics (100 unseeded episodes per env) are compared as distributions.
"""
import os

import numpy as np
import pytest
import torch

import a3c_oracle
import oracle
from conftest import GOLDEN

FIX = os.path.join(GOLDEN, "gym_eval_reference.npz")
CKPT = os.path.join(GOLDEN, "tracker_b200_e4096.dat")
CASES = ["BlockPartialRam", "BlockPartialNav", "MazePartialRam", "MazePartialNav"]

# the README "Evaluation" protocol, 100 episodes per env, measured by the reference's gym_eval.py on this checkpoint (CPU, unseeded;
# profiles/learning_r2/reference_gym_eval_on_b200_tracker.csv): (R_mean, R_std, EL_mean, S_rate)
REFERENCE_README_EVAL = {
    "Track2D-BlockPartialNav-v0": (303.50, 53.12, 492.22, 0.98),
    "Track2D-BlockPartialRam-v0": (356.37, 44.15, 495.76, 0.99),
    "Track2D-MazePartialNav-v0": (314.56, 36.97, 496.62, 0.99),
    "Track2D-MazePartialRam-v0": (353.08, 23.02, 500.00, 1.00),
}


def _state_dict():
    """the model gym_eval.py builds: tat-maze-lstm with the tracker loaded; the target net's weights do not matter for Ram / Nav ids"""
    sd = a3c_oracle.det_state_dict(tat=True, seed=5, scale=0.05)
    tracker = torch.load(CKPT, map_location="cpu")
    for k, v in tracker.items():
        sd["player0." + k] = v.clone()
    return sd


@pytest.mark.parametrize("case", CASES)
def test_oracle_replays_reference_gym_eval(case):
    g = np.load(FIX)
    env_id = "Track2D-%s-v0" % case
    lens, rets, seed = g[case + "_len"], g[case + "_ret"], int(g[case + "_seed"])
    n = min(len(lens), 4 if "Nav" in case else 6)  # CPU time: ~1 ms per policy step
    sd = _state_dict()
    torch.set_num_threads(1)  # batch-of-one CPU convolutions: intra-op threading only adds contention (main.py:3 OMP_NUM_THREADS=1)
    env = oracle.OracleEnv(env_id)
    env.seed(seed)
    env.init_maze()  # Track1v1Env.__init__ draws a map (track_1v1.py:45)
    with torch.no_grad():
        for k in range(n):
            obs = env.reset()
            hx, cx = torch.zeros(2, 128), torch.zeros(2, 128)
            ret, steps = np.zeros(2), 0
            while True:
                state = torch.from_numpy(obs.astype(np.float32)).unsqueeze(1)
                _, acts, _, _, (hx, cx), _ = a3c_oracle.forward(sd, state, hx, cx, tat=True, test=True)
                obs, rew, done, _ = env.step(acts)
                ret += rew
                steps += 1
                if done:
                    break
            assert steps == int(lens[k]), (case, k, steps, int(lens[k]))
            assert np.allclose(ret, rets[k], rtol=0, atol=1e-9), (case, k, ret, rets[k])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_replays_reference_gym_eval(case):
    from active_tracking_rl_b200 import envs
    from active_tracking_rl_b200.model import build_model
    from active_tracking_rl_b200.train import default_args
    g = np.load(FIX)
    env_id = "Track2D-%s-v0" % case
    lens, rets, seed = g[case + "_len"], g[case + "_ret"], int(g[case + "_seed"])
    env = envs.make(env_id, seed=seed, rng="numpy")
    args = default_args(network="tat-maze-lstm")
    torch.manual_seed(1)
    model = build_model(env.observation_space, env.action_space, args, "cuda:0").to("cuda:0")
    model.player0.load_state_dict(torch.load(CKPT, map_location="cuda:0"))
    model.eval()
    with torch.no_grad():
        for k in range(len(lens)):
            obs = env.reset()
            hx = torch.zeros(1, 2, 128, device="cuda:0")
            cx = torch.zeros_like(hx)
            ret, steps = np.zeros(2), 0
            while True:
                state = torch.from_numpy(obs).float().cuda().unsqueeze(0)  # (1, 2, 1, 13, 13)
                _, action, _, _, (hx, cx), _ = model((state, (hx, cx)), True)
                obs, rew, done, _ = env.step(action[0].tolist())
                ret += rew
                steps += 1
                if done:
                    break
            assert steps == int(lens[k]), (case, k, steps, int(lens[k]))
            assert np.allclose(ret, rets[k], rtol=0, atol=1e-9), (case, k, ret, rets[k])
    assert env.vec.status() == 0
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", sorted(REFERENCE_README_EVAL))
def test_batched_gym_eval_matches_reference_statistics(env_id):
    """gym_eval.evaluate (100 envs of one batch, Philox worlds) vs the reference's gym_eval.py on the same checkpoint: different
    episodes, same distribution -- within 4 standard errors of the mean on R_mean, and the success rates a few episodes apart"""
    from active_tracking_rl_b200 import gym_eval
    from active_tracking_rl_b200.envs import Track2DVecEnv
    from active_tracking_rl_b200.model import build_model
    from active_tracking_rl_b200.train import default_args
    probe = Track2DVecEnv(env_id, num_envs=1, device="cuda:0", seed=1)
    torch.manual_seed(1)
    model = build_model(probe.observation_space, probe.action_space, default_args(network="tat-maze-lstm"), "cuda:0").to("cuda:0")
    probe.close()
    gym_eval.load_weights(model, load_tracker=CKPT, device="cuda:0")
    r_mean, r_std, el_mean, s_rate = REFERENCE_README_EVAL[env_id]
    st = gym_eval.evaluate(model, env_id, num_episodes=400, seed=7, device="cuda:0")
    se = np.sqrt(r_std ** 2 / 100 + st["R_std"] ** 2 / 400)
    assert abs(st["R_mean"] - r_mean) <= 4 * se + 1.0, (st, REFERENCE_README_EVAL[env_id])
    assert abs(st["S_rate"] - s_rate) <= 0.05 and abs(st["EL_mean"] - el_mean) <= 25, (st, REFERENCE_README_EVAL[env_id])


@pytest.mark.reference
def test_reference_loads_the_b200_checkpoint_and_tracks():
    """live, where /root/reference exists: the unmodified gym_eval.py on the committed checkpoint (3 episodes of Block-Ram)"""
    import ref
    if not ref.reference_available():
        pytest.skip("reference tree not present")
    import make_golden_eval
    el, rw = make_golden_eval.run_case("Track2D-BlockPartialRam-v0", 3, 4001)
    g = np.load(FIX)
    assert (el == g["BlockPartialRam_len"][:3]).all() and np.allclose(rw, g["BlockPartialRam_ret"][:3], rtol=0, atol=1e-9)
