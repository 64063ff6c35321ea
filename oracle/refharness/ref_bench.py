"""Times the UNMODIFIED reference's A3C path on the host cores -- TEST / BENCH INFRASTRUCTURE ONLY.

main.py:86-116 is reproduced as a launcher: one shared model + SharedAdam in shared memory, W processes running the reference's own
`train(rank, args, shared_model, optimizer, train_modes, n_iters, env)` (train.py:15-113) Hogwild-style, OMP_NUM_THREADS=1.  Nothing
of the reference is edited: train() takes the env as an argument, so each worker gets the reference's own
`create_env(env_id, args)` wrapped in a counter that adds 1 to a shared integer per env.step(); the workers are stopped the way
test.py:131-133 stops them (train_modes[rank] = -100).  env-steps/s = counted steps of all workers over a fixed window after a warm-up.

The reference tree is oracle/_ref (populate_ref.py) or /root/reference, whichever exists.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
LOCAL = os.path.join(os.path.dirname(HERE), "_ref")


def reference_root():
    for cand in (os.environ.get("TRACK2D_REFERENCE_ROOT"), LOCAL, "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "train.py")):
            return cand
    return None


class _CountingEnv(object):
    """delegates everything to the reference env; counts steps"""

    def __init__(self, env, counter):
        self._env, self._counter = env, counter

    def step(self, action):
        out = self._env.step(action)
        with self._counter.get_lock():
            self._counter.value += 1
        return out

    def __getattr__(self, name):
        return getattr(self._env, name)


def _worker(rank, args, shared_model, optimizer, train_modes, n_iters, counter):
    import torch
    torch.set_num_threads(1)
    from environment import create_env
    from train import train
    env = _CountingEnv(create_env(args.env, args), counter)
    train(rank, args, shared_model, optimizer, train_modes, n_iters, env)


def run(seconds, workers, warmup_s=3.0, env_id="Track2D-BlockPartialPZR-v0", seed=1, samples=1, **arg_overrides):
    """-> (env_steps_per_s, workers) of the reference's own training path; with samples > 1 the first element is the list of the
    rates of `samples` consecutive windows of `seconds` each (one set of workers, started once)"""
    root = reference_root()
    if root is None:
        raise RuntimeError("no reference tree (oracle/_ref or /root/reference)")
    os.environ["TRACK2D_REFERENCE_ROOT"] = root
    os.environ["OMP_NUM_THREADS"] = "1"  # main.py:3
    sys.path.insert(0, HERE)
    import ref
    ref.REFERENCE_ROOT = root
    ref.load_reference(neutralise_reseed=False)
    import warnings
    warnings.filterwarnings("ignore")
    import torch
    import torch.multiprocessing as mp
    from environment import create_env
    from model import build_model
    from shared_optim import SharedAdam
    args = ref.RefArgs(env=env_id, workers=workers, seed=seed, log_dir="/tmp/track2d_ref_bench/%d/" % os.getpid(), **arg_overrides)
    torch.manual_seed(seed)
    env = create_env(args.env, args)
    shared_model = build_model(env.observation_space, env.action_space, args, torch.device("cpu"))
    if args.train_mode == 0:  # main.py:72-77
        params = shared_model.player0.parameters()
    elif args.train_mode == 1:
        params = shared_model.player1.parameters()
    else:
        params = shared_model.parameters()
    shared_model.share_memory()
    optimizer = SharedAdam(params, lr=args.lr, amsgrad=args.amsgrad)
    optimizer.share_memory()
    env.close()
    ctx = mp.get_context("fork")
    manager = ctx.Manager()
    train_modes, n_iters = manager.list(), manager.list()
    counter = ctx.Value("l", 0)
    procs = [ctx.Process(target=_worker, args=(r, args, shared_model, optimizer, train_modes, n_iters, counter), daemon=True) for r in range(workers)]
    for p in procs:
        p.start()
    t_dead = time.time() + 120
    while len(train_modes) < workers and time.time() < t_dead:  # every worker has built its env and model
        time.sleep(0.05)
    time.sleep(warmup_s)
    rates = []
    for _ in range(max(1, int(samples))):
        c0, t0 = counter.value, time.time()
        time.sleep(seconds)
        c1, t1 = counter.value, time.time()
        rates.append((c1 - c0) / (t1 - t0))
    for r in range(len(train_modes)):
        train_modes[r] = -100
    for p in procs:
        p.join(timeout=20)
        if p.is_alive():
            p.terminate()
    manager.shutdown()
    return (rates[0] if int(samples) <= 1 else rates), workers


if __name__ == "__main__":
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    W = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    v, W = run(secs, W)
    print("reference A3C (%s): %.1f env-steps/s with %d workers" % (reference_root(), v, W))
