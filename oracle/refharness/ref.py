"""Loader for the UNMODIFIED reference (zfw1226/active_tracking_rl) -- TEST INFRASTRUCTURE ONLY.

Only usable where /root/reference exists (the build container).  It puts the import stubs in
oracle/refharness/stubs/ (gym 0.12.5, matplotlib, skimage, gym_unrealcv, tensorboardX -- none are
installed here) and the reference tree on sys.path and applies the two oracle patches SURVEY.md
section 8(c) lists:

  (i)  np.random.seed() with NO argument becomes a no-op.  The reference re-seeds the global numpy RNG
       from OS entropy on every goal/spawn sample (generators.py:41,56), which makes it
       irreproducible; with the patch, np.random.seed(s) before reset() pins the whole episode.
  (ii) numpy DeprecationWarning for int(array([a])) (track_1v1.py:87 fed by navigator.py:82) is
       silenced.

No reference file is copied or edited.  Nothing in the product path imports this module; it is used
by oracle/refharness/make_golden.py (fixture generator) and by tests that validate the oracle
restatement when the reference is present.
"""
import os
import sys
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("TRACK2D_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")

_loaded = False
_orig_seed = np.random.seed


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "envs", "gym-track2d", "gym_track2d", "envs", "track_1v1.py"))


def _patched_seed(seed=None):
    if seed is None:
        return None  # patch (i): neutralise entropy re-seeding
    return _orig_seed(seed)


def load_reference(neutralise_reseed=True):
    """Make `import gym, gym_track2d, environment, model, player_util, ...` resolve to the stubs +
    the reference tree.  Returns the gym stub module."""
    global _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not _loaded:
        for p in (os.path.join(REFERENCE_ROOT, "envs", "gym-track2d"), REFERENCE_ROOT, _STUBS):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        warnings.filterwarnings("ignore", category=DeprecationWarning)
        _loaded = True
    np.random.seed = _patched_seed if neutralise_reseed else _orig_seed
    import gym  # noqa: E402  (the stub)
    import gym_track2d  # noqa: F401,E402  (reference: registers the 72 ids)
    return gym


def make_env(env_id, seed=None):
    """gym.make(env_id) on the real reference.  NOTE: Track1v1Env.__init__ already draws a map
    (track_1v1.py:45), so seed BEFORE make_env if the constructor's draws matter; `seed` here
    re-seeds the global numpy RNG after construction, i.e. pins the first reset()."""
    gym = load_reference()
    env = gym.make(env_id)
    if seed is not None:
        np.random.seed(int(seed))
    return env


class RefArgs(object):
    """The argparse namespace main.py:16-50 builds, with its defaults."""

    def __init__(self, **kw):
        self.lr = 0.001
        self.gamma = 0.9
        self.tau = 1.00
        self.entropy = 0.01
        self.entropy_target = 0.2
        self.seed = 1
        self.workers = 1
        self.num_steps = 20
        self.test_eps = 100
        self.env = 'Track2D-BlockPartialPZR-v0'
        self.env_base = 'Track2D-BlockPartialNav-v0'
        self.optimizer = 'Adam'
        self.amsgrad = True
        self.load_model_dir = None
        self.log_dir = '/tmp/track2d_ref_logs/'
        self.network = 'tat-maze-lstm'
        self.aux = 'reward'
        self.gpu_ids = [-1]
        self.obs = 'img'
        self.single = False
        self.gray = False
        self.crop = False
        self.inv = False
        self.rescale = False
        self.render = False
        self.shared_optimizer = True
        self.split = False
        self.train_mode = -1
        self.stack_frames = 1
        self.input_size = 80
        self.rnn_out = 128
        self.sleep_time = 0
        self.max_step = 150000
        self.init_step = -1
        for k, v in kw.items():
            setattr(self, k, v)
