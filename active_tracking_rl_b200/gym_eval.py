"""Greedy evaluation and checkpointing -- the batched form of gym_eval.py:94-141 and of the evaluator process
test.py:56-134.

The reference plays `num_episodes` greedy episodes one after another; here they are `num_episodes` envs of one
batch, each playing exactly one episode (auto-reset off, finished envs masked).  Statistics and the CSV layout are
gym_eval.py's: R_mean / R_std (tracker return), EL_mean / EL_std (episode length), S_rate = share of episodes that
reach the 500-step limit (gym_eval.py:114).  Checkpoints are plain reference-format `state_dict`s, so `.dat`
files interchange with the reference both ways (test.py:112-127: all-best-{n}.dat / all-new.dat and, with --split,
tracker-/target-{best,new}.dat).

    python -m active_tracking_rl_b200.gym_eval --env Track2D-BlockPartialNav-v0 --network tat-maze-lstm \
        --load-tracker logs/tracker-best.dat --num-episodes 100 --csv results.csv
"""
import argparse
import csv
import os

import torch

from .envs import Track2DVecEnv
from .model import build_model
from .train import default_args


def load_weights(model, load_model_dir=None, load_tracker=None, load_target=None, device="cpu"):
    """gym_eval.py:74-92"""
    if load_model_dir is not None:
        model.load_state_dict(torch.load(load_model_dir, map_location=device), strict=False)
    if load_tracker is not None:
        model.player0.load_state_dict(torch.load(load_tracker, map_location=device))
    if load_target is not None:
        model.player1.load_state_dict(torch.load(load_target, map_location=device))
    return model


@torch.no_grad()
def evaluate(model, env_id, num_episodes=100, seed=1, device="cuda:0", rng="philox", max_steps=500):
    """Play `num_episodes` greedy episodes (player.action_test, player_util.py:69-82) and return gym_eval.py's
    statistics.  The model's target head is ignored by the env for Ram / Nav / RPF ids, as in the reference."""
    device = torch.device(device)
    env = Track2DVecEnv(env_id, num_envs=num_episodes, device=device, seed=seed, rng=rng, auto_reset=False)
    was_training = model.training
    model.eval()
    E = num_episodes
    obs = env.reset()
    hx = torch.zeros(E, 2, model.player0.lstm.hidden_size, device=device)
    cx = torch.zeros_like(hx)
    alive = torch.ones(E, dtype=torch.bool, device=device)
    ret = torch.zeros(E, 2, dtype=torch.float64, device=device)
    length = torch.zeros(E, dtype=torch.int32, device=device)
    for _ in range(max_steps):
        _, action, _, _, (hx, cx), _ = model((obs, (hx, cx)), True)
        obs, reward, done = env.step(action.to(torch.int32).contiguous())
        ret += reward.double() * alive.unsqueeze(1)
        length += alive.to(torch.int32)
        alive &= ~done.bool()
        if not bool(alive.any()):
            break
    env.close()
    if was_training:
        model.train()
    r0 = ret[:, 0]
    el = length.double()
    return {"Env": env_id, "Seed": seed, "R_mean": float(r0.mean()), "R_std": float(r0.std(unbiased=False)),
            "EL_mean": float(el.mean()), "EL_std": float(el.std(unbiased=False)), "S_rate": float((length >= 500).double().mean()),
            "R_target_mean": float(ret[:, 1].mean()), "episodes": E}


HEADER = ['Env', 'Seed', 'R_mean', 'R_std', 'EL_mean', 'EL_std', 'S_rate']


def write_csv(path, row):
    """gym_eval.py:128-141: append, header on first write"""
    new = not os.path.exists(path)
    with open(path, 'w' if new else 'a') as f:
        w = csv.DictWriter(f, HEADER, extrasaction='ignore')
        if new:
            w.writeheader()
        w.writerows([row])


def _compact(sd):
    """the parameters are views into SharedAdam's flat buffer; torch.save would write the WHOLE underlying storage for each of them"""
    return {k: v.detach().clone() for k, v in sd.items()}


class Evaluator(object):
    """test.py:56-134 without the process: evaluate on --env-base, keep the best tracker return, write checkpoints,
    and say which training mode comes next (init_step schedule, test.py:84-91)."""

    def __init__(self, args, log_dir, device="cuda:0"):
        self.args, self.log_dir, self.device = args, log_dir, device
        self.max_score = -100
        os.makedirs(log_dir, exist_ok=True)

    def run(self, model, n_iter):
        env_id = self.args.env_base if self.args.env_base is not None else self.args.env
        stats = evaluate(model, env_id, self.args.test_eps, self.args.seed, self.device)
        if stats["R_mean"] >= self.max_score:  # test.py:112-122
            self.max_score = stats["R_mean"]
            paths = ('all-best-{0}.dat'.format(n_iter), 'tracker-best.dat', 'target-best.dat')
        else:
            paths = ('all-new.dat', 'tracker-new.dat', 'target-new.dat')
        torch.save(_compact(model.state_dict()), os.path.join(self.log_dir, paths[0]))
        if getattr(self.args, 'split', False):
            torch.save(_compact(model.player0.state_dict()), os.path.join(self.log_dir, paths[1]))
            torch.save(_compact(model.player1.state_dict()), os.path.join(self.log_dir, paths[2]))
        stats["checkpoint"] = paths[0]
        stats["next_train_mode"] = 0 if n_iter < self.args.init_step else self.args.train_mode
        stats["stop"] = n_iter > self.args.max_step
        return stats


def main():
    p = argparse.ArgumentParser(description='A3C_EVAL (batched, on device)')
    p.add_argument('--env', default='Track2D-BlockPartialNav-v0')
    p.add_argument('--num-episodes', type=int, default=100)
    p.add_argument('--load-model-dir', default=None)
    p.add_argument('--load-tracker', default=None)
    p.add_argument('--load-target', default=None)
    p.add_argument('--csv', default=None)
    p.add_argument('--network', default='tat-maze-lstm')
    p.add_argument('--stack-frames', type=int, default=1)
    p.add_argument('--seed', type=int, default=1)
    p.add_argument('--gpu-id', type=int, default=0)
    p.add_argument('--rnn-out', type=int, default=128)
    a = p.parse_args()
    device = torch.device('cuda:%d' % max(a.gpu_id, 0))
    torch.cuda.set_device(device)
    args = default_args(network=a.network, stack_frames=a.stack_frames, rnn_out=a.rnn_out, seed=a.seed)
    torch.manual_seed(a.seed)
    probe = Track2DVecEnv(a.env, num_envs=1, device=device, seed=a.seed)
    model = build_model(probe.observation_space, probe.action_space, args, device).to(device)
    probe.close()
    load_weights(model, a.load_model_dir, a.load_tracker, a.load_target, device)
    stats = evaluate(model, a.env, a.num_episodes, a.seed, device)
    print("R_mean: {R_mean:.3f}, R_std: {R_std:.3f}, EL_mean: {EL_mean:.2f}, EL_std {EL_std:.2f}, S_rate: {S_rate:.3f}".format(**stats))
    if a.csv is not None:
        write_csv(a.csv, stats)


if __name__ == '__main__':
    main()
