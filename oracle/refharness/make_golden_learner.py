"""Record learner golden vectors from the UNMODIFIED reference (model.py, player_util.py, shared_optim.py,
utils.py, environment.py) -- TEST INFRASTRUCTURE ONLY.  Needs /root/reference.

    python oracle/refharness/make_golden_learner.py     # rewrites tests/golden/learner_*.npz

Per case: the reference's own train() loop body (train.py:69-95) is replayed for a few iterations on a
seeded env with scripted actions (torch.Tensor.multinomial is intercepted in the harness; no reference file is
edited), deterministic weights (oracle/a3c_oracle.det_state_dict, so fixtures stay small), the reference's
SharedAdam, and the one oracle patch SURVEY 5 lists for learner parity: zero_grad(set_to_none=False).
Recorded: obs / actions / rewards / dones, per-step values, log-probs, entropies, reward predictions,
the three loss tensors, per-tensor gradient norms and sums, and per-tensor sums / norms of the shared
weights after every SharedAdam step.
"""
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
sys.path.insert(0, os.path.dirname(_HERE))
import ref  # noqa: E402
import a3c_oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "tests", "golden")

CASES = [
    # name, env id, network, aux, train_mode, entropy_target, env seed, iterations, weight scale
    ("advat", "Track2D-BlockPartialPZR-v0", "tat-maze-lstm", "reward", -1, 0.2, 77, 4, 0.08),
    ("naive", "Track2D-MazePartialAdv-v0", "maze-lstm", "none", -1, 0.01, 78, 3, 0.08),
    ("tracker_only", "Track2D-BlockPartialRam-v0", "tat-maze-lstm", "reward", 0, 0.2, 79, 3, 0.08),
    ("target_only_bigw", "Track2D-BlockPartialPZR-v0", "tat-maze-lstm", "reward", 1, 0.2, 80, 3, 0.5),
]


def record(name, env_id, network, aux, train_mode, entropy_target, env_seed, iters, scale):
    ref.load_reference()
    from environment import create_env
    from model import build_model
    from player_util import Agent
    from shared_optim import SharedAdam

    args = ref.RefArgs(env=env_id, network=network, aux=aux, train_mode=train_mode, entropy_target=entropy_target)
    tat = 'tat' in network
    device = torch.device('cpu')
    torch.manual_seed(5)
    np.random.seed(env_seed)
    env = create_env(env_id, args)
    shared_model = build_model(env.observation_space, env.action_space, args, device)
    sd0 = a3c_oracle.det_state_dict(tat=tat, seed=1234, scale=scale)
    shared_model.load_state_dict(sd0)
    opt_params = (shared_model.player0.parameters() if train_mode == 0 else
                  shared_model.player1.parameters() if train_mode == 1 else shared_model.parameters())
    optimizer = SharedAdam(opt_params, lr=args.lr, amsgrad=args.amsgrad)
    # train.py:39-44: ONE generator, handed to every optimize() call
    params = (shared_model.player0.parameters() if train_mode == 0 else
              shared_model.player1.parameters() if train_mode == 1 else shared_model.parameters())
    player = Agent(None, env, args, None, device)
    player.w_entropy_target = args.entropy_target
    player.gpu_id = -1
    player.model = build_model(env.observation_space, env.action_space, args, device)
    player.model.train()
    orig_zero = player.model.zero_grad
    player.model.zero_grad = lambda: orig_zero(set_to_none=False)  # oracle patch (iii)

    arng = np.random.RandomState(900 + env_seed)
    orig_multinomial = torch.Tensor.multinomial
    taken = []

    def scripted(self, n, *a, **k):
        act = int(arng.randint(4))
        taken.append(act)
        return torch.tensor([[act]])
    torch.Tensor.multinomial = scripted

    rec = dict(obs=[], actions=[], rewards=[], dones=[], values=[], log_probs=[], entropies=[], preds=[], iter_len=[], iter_done=[],
               policy_loss=[], value_loss=[], pred_loss=[], boot_actions=[], grad_norm=[], grad_sum=[], param_sum=[], param_norm=[], reset_before=[])
    names = list(shared_model.state_dict().keys())
    try:
        np.random.seed(env_seed)
        player.reset()
        for it in range(iters):
            player.model.load_state_dict(shared_model.state_dict())
            did_reset = False
            if player.done and it > 0:
                player.reset()
                did_reset = True
            rec['reset_before'].append(did_reset or it == 0)
            player.update_rnn_hiden()
            n = 0
            rec['obs'].append(player.state.numpy().astype(np.uint8).reshape(2, 169))
            cap = 20 if it != 1 else 7  # one short rollout that does not end the episode (bootstrapped)
            for i in range(cap):
                k0 = len(taken)
                player.action_train()
                n += 1
                acts = taken[k0:k0 + 2]
                rec['actions'].append(acts)
                rec['rewards'].append(player.reward.numpy().copy())
                rec['dones'].append(bool(player.done))
                rec['values'].append(player.values[-1].detach().numpy().reshape(2))
                rec['log_probs'].append(player.log_probs[-1].detach().numpy().reshape(2))
                rec['entropies'].append(player.entropies[-1].detach().numpy().reshape(2))
                rp = player.preds[-1]
                rec['preds'].append(float(rp.item()) if torch.is_tensor(rp) else 0.0)
                rec['obs'].append(player.state.numpy().astype(np.uint8).reshape(2, 169))
                if player.done:
                    break
            rec['iter_len'].append(n)
            rec['iter_done'].append(bool(player.done))
            k0 = len(taken)
            pl, vl, ent, prl = player.optimize(params, optimizer, shared_model, train_mode, device)
            # the bootstrap forward inside optimize() SAMPLES actions; the tracker's one feeds the TAT target's value
            rec['boot_actions'].append(taken[k0:k0 + 2] if len(taken) >= k0 + 2 else [-1, -1])
            rec['policy_loss'].append(pl.detach().numpy().reshape(2))
            rec['value_loss'].append(vl.detach().numpy().reshape(2))
            rec['pred_loss'].append(float(prl.sum().item()))
            gn, gs = [], []
            for nme, p_ in player.model.named_parameters():
                g = p_.grad
                gn.append(float(g.norm()) if g is not None else 0.0)
                gs.append(float(g.double().sum()) if g is not None else 0.0)
            rec['grad_norm'].append(gn)
            rec['grad_sum'].append(gs)
            ssd = shared_model.state_dict()
            rec['param_sum'].append([float(ssd[k].double().sum()) for k in names])
            rec['param_norm'].append([float(ssd[k].double().norm()) for k in names])
    finally:
        torch.Tensor.multinomial = orig_multinomial
    out = {k: np.asarray(v) for k, v in rec.items()}
    out['param_names'] = np.asarray(names)
    out['grad_names'] = np.asarray([n for n, _ in player.model.named_parameters()])
    out['meta'] = np.asarray([env_id, network, aux, str(train_mode), str(entropy_target), str(env_seed), str(scale)])
    total_gn = [float(np.sqrt((np.asarray(g) ** 2).sum())) for g in rec['grad_norm']]
    out['total_grad_norm'] = np.asarray(total_gn)
    path = os.path.join(OUT, 'learner_%s.npz' % name)
    np.savez_compressed(path, **out)
    print('%-18s iters=%d steps=%d dones=%s total_grad_norm=%s  %.1f KB' % (name, iters, len(rec['dones']), rec['iter_done'],
                                                                          ['%.1f' % g for g in total_gn], os.path.getsize(path) / 1024))


if __name__ == '__main__':
    for c in CASES:
        record(*c)
