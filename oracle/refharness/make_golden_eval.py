"""Fixture generator for the evaluation path: runs the UNMODIFIED reference gym_eval.py (README "Evaluation" protocol) on a
committed tracker checkpoint and records what it measured, per episode.  TEST INFRASTRUCTURE; needs /root/reference.

    python oracle/refharness/make_golden_eval.py            -> tests/golden/gym_eval_reference.npz

gym_eval.py is executed as `__main__` with runpy (its own argparse, Agent.action_test loop, statistics).  It ends in os._exit(0)
(gym_eval.py:143); that call is intercepted for the duration of the run and its caller's module globals `len_lis` and `rewards_his` --
the per-episode lengths and returns behind the R_mean / EL_mean / S_rate it logs -- are read at that point.  No reference file is edited.
The global numpy RNG is seeded right before (np.random.seed(SEED)) and the reference's no-argument np.random.seed() calls are
neutralised (ref.py patch (i)), so the episodes are reproducible: `np.random.seed(SEED); gym.make(id); reset(); ...` is exactly
what Track1v1Env(id, seed=SEED, rng='numpy') and the C oracle replay.  The checkpoint was trained by THIS repo's learner
(tools/learn_experiment.sh, 4,096 envs) and is loaded by the reference through its own --load-tracker path: checkpoint
interchange and task-quality parity in one fixture.
"""
import os
import runpy
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import ref  # noqa: E402

CKPT = os.path.join(ROOT, "tests", "golden", "tracker_b200_e4096.dat")
CASES = [("Track2D-BlockPartialRam-v0", 20, 4001), ("Track2D-BlockPartialNav-v0", 12, 4002), ("Track2D-MazePartialRam-v0", 12, 4003),
         ("Track2D-MazePartialNav-v0", 8, 4004)]


def run_case(env_id, episodes, seed):
    ref.load_reference(neutralise_reseed=True)
    work = "/tmp/track2d_ref_eval_work"
    os.makedirs(os.path.join(work, "logs"), exist_ok=True)
    os.chdir(work)
    argv = sys.argv
    sys.argv = ["gym_eval.py", "--env", env_id, "--network", "tat-maze-lstm", "--load-tracker", CKPT, "--num-episodes", str(episodes),
                "--log-dir", os.path.join(work, "logs") + "/"]
    captured = {}

    class _Finished(BaseException):
        pass

    def _exit_hook(code=0):
        g = sys._getframe(1).f_globals
        captured["len"] = np.asarray(g["len_lis"], np.int64)
        captured["ret"] = np.asarray(g["rewards_his"], np.float64)
        raise _Finished()

    real_exit = os._exit
    os._exit = _exit_hook
    np.random.seed(seed)
    try:
        runpy.run_path(os.path.join(ref.REFERENCE_ROOT, "gym_eval.py"), run_name="__main__")
    except _Finished:
        pass
    finally:
        os._exit = real_exit
        sys.argv = argv
    return captured["len"], captured["ret"]


def main():
    out = {}
    for env_id, episodes, seed in CASES:
        el, rw = run_case(env_id, episodes, seed)
        key = env_id[len("Track2D-"):-3]
        out[key + "_seed"], out[key + "_len"], out[key + "_ret"] = np.int64(seed), el, rw
        print("%s: EL_mean %.2f  R_mean %.3f  S_rate %.2f" % (env_id, el.mean(), rw[:, 0].mean(), float((el >= 500).mean())), flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gym_eval_reference.npz"), **out)
    # the README-protocol numbers (100 unseeded episodes per env) measured by the same script are kept in profiles/learning_r2/


if __name__ == "__main__":
    main()
