"""The synchronous on-device actor-learner that replaces the reference's W Hogwild worker processes
(main.py:102-116, train.py:15-113): E envs per GPU advance in lock step -- policy.step and env.step
alternate on one CUDA stream with no host round trip -- and every num_steps steps ONE update is applied.
Across GPUs (one process each) the env shards are independent; the only exchange is one all-reduce of
the flat gradient per rollout (over NVLink peer memory, peer.py; NCCL as the fallback), after which every rank applies the identical fused
SharedAdam update (so the replicas stay bit-identical without ever broadcasting weights again).

    python -m active_tracking_rl_b200.train --env Track2D-BlockPartialPZR-v0 --num-envs 65536 --iters 100
    torchrun --nproc-per-node 8 -m active_tracking_rl_b200.train ...
"""
import argparse
import os
import sys
import time

import torch

from . import blas
from .environment import create_env
from .model import build_model
from .player_util import Agent
from .shared_optim import SharedAdam


def make_parser():
    """the flags of main.py:16-50 (same names and defaults) + the batched-run additions"""
    p = argparse.ArgumentParser(description='A3C (synchronous, batched, on-device)')
    p.add_argument('--lr', type=float, default=0.001)
    p.add_argument('--gamma', type=float, default=0.9)
    p.add_argument('--tau', type=float, default=1.00)
    p.add_argument('--entropy', type=float, default=0.01)
    p.add_argument('--entropy-target', type=float, default=0.2)
    p.add_argument('--seed', type=int, default=1)
    p.add_argument('--workers', type=int, default=1, help='kept for CLI compatibility; parallelism is --num-envs')
    p.add_argument('--num-steps', type=int, default=20)
    p.add_argument('--test-eps', type=int, default=100)
    p.add_argument('--env', default='Track2D-BlockPartialPZR-v0')
    p.add_argument('--env-base', default='Track2D-BlockPartialNav-v0')
    p.add_argument('--optimizer', default='Adam')
    p.add_argument('--amsgrad', default=True)
    p.add_argument('--load-model-dir', default=None)
    p.add_argument('--log-dir', default='logs/')
    p.add_argument('--network', default='tat-maze-lstm')
    p.add_argument('--aux', default='reward')
    p.add_argument('--shared-optimizer', dest='shared_optimizer', action='store_true')
    p.add_argument('--split', dest='split', action='store_true')
    p.add_argument('--train-mode', type=int, default=-1)
    p.add_argument('--stack-frames', type=int, default=1)
    p.add_argument('--rnn-out', type=int, default=128)
    p.add_argument('--max-step', type=int, default=150000)
    p.add_argument('--init-step', type=int, default=-1)
    # batched-run additions
    p.add_argument('--num-envs', type=int, default=4096, help='envs per GPU')
    p.add_argument('--iters', type=int, default=100, help='rollout+update iterations to run (the run also stops once --max-step updates are done)')
    p.add_argument('--eval-every', type=int, default=0,
                   help='evaluate + checkpoint every this many updates on rank 0 (test.py:56-134); 0 = only once, at the end')
    p.add_argument('--graph', action='store_true', help='replay each iteration from a CUDA graph (Trainer.capture)')
    p.add_argument('--plan-ahead', dest='plan_ahead', action='store_true',
                   help='Maze / Nav ids: next-episode worlds (and Nav plans) are prepared on a side stream while the policy computes, '
                        'so an auto-reset is a copy (same episodes: the RNG is counter-based)')
    p.add_argument('--max-grad-norm', type=float, default=0.0,
                   help='0 = the reference\'s effective behaviour (its clip_grad_norm_(params, 50) is inert); 50 = its intent')
    p.add_argument('--fp32-emulation', action='store_true', help='fp32 GEMMs through cuBLAS 12.9 BF16x9 emulation (see blas.py)')
    p.add_argument('--no-fused', dest='fused', action='store_false',
                   help='run the policy through model.py\'s autograd path instead of the hand-written forward / backward (learner.FusedA3C)')
    p.add_argument('--tf32', action='store_true', help='allow TF32 tensor-core math in the policy (off: float32 like the reference)')
    return p


def default_args(**kw):
    args = make_parser().parse_args([])
    args.single = False
    args.rescale = False
    for k, v in kw.items():
        setattr(args, k, v)
    return args


class Trainer(object):
    def __init__(self, args, device, rank=0, world_size=1, rng="philox"):
        self.args, self.device, self.rank, self.world_size = args, torch.device(device), rank, world_size
        # the reference computes in float32; keep cuDNN / cuBLAS from silently dropping to TF32
        torch.backends.cudnn.allow_tf32 = bool(getattr(args, 'tf32', False))
        torch.backends.cuda.matmul.allow_tf32 = bool(getattr(args, 'tf32', False))
        if not hasattr(args, 'single'):
            args.single = False
        self.env = create_env(args.env, args, num_envs=args.num_envs, device=self.device, seed=args.seed + rank, rng=rng)
        torch.manual_seed(args.seed)  # identical initial weights on every rank (the reference shares ONE model)
        self.model = build_model(self.env.observation_space, self.env.action_space, args, self.device).to(self.device)
        if args.load_model_dir is not None:
            self.model.load_state_dict(torch.load(args.load_model_dir, map_location=self.device))
        torch.manual_seed(args.seed + rank)  # train.py:20: per-worker sampling stream
        torch.cuda.manual_seed(args.seed + rank)
        if args.train_mode == 0:
            params = self.model.player0.parameters()
        elif args.train_mode == 1:
            params = self.model.player1.parameters()
        else:
            params = self.model.parameters()
        self.optimizer = SharedAdam(params, lr=args.lr, amsgrad=args.amsgrad)
        self.player = Agent(self.model, self.env, args, None, self.device)
        self.player.w_entropy_target = args.entropy_target
        self.player.max_grad_norm = float(getattr(args, 'max_grad_norm', 0.0))
        if self.player.engine is not None:  # train.py:20: every worker samples from its own stream
            self.player.engine.seed = (int(args.seed) + 1000003 * rank) & 0xFFFFFFFFFFFFFFFF
        self.player.reset()
        self.n_iter = 0
        self.allreduce = None
        self.peer = None  # all-reduce over NVLink peer memory (peer.PeerAllReduce): plain kernels, capturable with the iteration
        if world_size > 1:
            import torch.distributed as dist
            self.allreduce = lambda g: dist.all_reduce(g, op=dist.ReduceOp.SUM)
            if self.device.type == 'cuda' and world_size <= 8 and os.environ.get("T2D_PEER_ALLREDUCE", "1") != "0":
                ok = 1
                try:
                    from .peer import PeerAllReduce
                    self.peer = PeerAllReduce(self.optimizer.fp.grad.numel(), self.device, rank, world_size)
                except Exception as ex:  # noqa: BLE001  (no peer access / IPC on this box: the library collective does it)
                    ok = 0
                    sys.stderr.write("rank %d: peer-memory all-reduce unavailable (%s); using torch.distributed\n" % (rank, str(ex).splitlines()[0] if str(ex) else repr(ex)))
                flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # every rank takes the same path
                if int(flag.item()) == 1:
                    self.allreduce = self.peer
                elif self.peer is not None:
                    self.peer.close()
                    self.peer = None

    def iteration(self, training_mode=None, host=None):
        """train.py:69-95 for all envs: detach LSTM state, num_steps x (policy.step, env.step), optimize.
        host: pinned host buffers -> every env.step goes through the host-buffer C ABI (see Agent.action_train)."""
        mode = self.args.train_mode if training_mode is None else training_mode
        p = self.player
        p.update_rnn_hiden()
        for _ in range(self.args.num_steps):
            p.action_train(host=host)
        out = p.optimize(None, self.optimizer, self.model, mode, None, world_size=self.world_size, allreduce=self.allreduce)
        self.n_iter += 1
        return out

    # ---- CUDA-graph replay of the whole iteration ------------------------------------------------------------
    def capture(self, training_mode=None, warmup=3):
        """Capture one full iteration -- 20 x (policy forward, env.step, auto-reset), bootstrap, GAE, losses, backward,
        [gradient all-reduce over peer memory], fused SharedAdam -- into ONE CUDA graph.  A rollout is ~4,300 kernel launches; below
        ~16k envs per GPU the Python/driver launch path, not the GPU, bounds the step, and a graph replay removes it.
        Everything inside is already stream-ordered with static shapes and no host synchronisation (the env kernels
        are plain launches on the capturing stream), which is what makes the capture legal."""
        mode = self.args.train_mode if training_mode is None else training_mode
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):  # warm up allocator, cuBLAS workspaces and autograd on the capture stream
            for _ in range(warmup):
                self.iteration(mode)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        # (the optimizer's update count lives in device memory and is advanced by the captured kernels themselves;
        # capture only records, so the counters below are untouched until the first replay)
        step0, iter0, nsteps0 = self.optimizer.step_count, self.n_iter, self.player.n_steps
        l0 = self.env.lib.track2d_launch_count()
        self._graph_apply = None
        if self.world_size == 1 or self.peer is not None:  # the peer-memory all-reduce is ordinary kernels: one graph on any number of GPUs
            with torch.cuda.graph(self._graph, stream=side):
                self._graph_out = self.iteration(mode)
        elif os.environ.get("T2D_NCCL_IN_GRAPH", "0") == "1":
            # the all-reduce captured INSIDE the one graph.  The NCCL watchdog thread polls CUDA events, which a capture in the default
            # "global" error mode forbids process-wide (that was the dead-lock); "thread_local" confines the restriction to this thread.
            with torch.cuda.graph(self._graph, stream=side, capture_error_mode="thread_local"):
                self._graph_out = self.iteration(mode)
        else:
            # multi-GPU: two graphs with the NCCL all-reduce launched eagerly between them (capturing the collective itself
            # dead-locked on this image): [rollout, losses, backward] | all-reduce | [SharedAdam, state hand-over]
            p = self.player
            with torch.cuda.graph(self._graph, stream=side):
                p.update_rnn_hiden()
                for _ in range(self.args.num_steps):
                    p.action_train()
                self._graph_out = p.optimize(None, self.optimizer, self.model, mode, None, world_size=self.world_size, allreduce=None, apply=False)
            self._graph_apply = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_apply, stream=side, pool=self._graph.pool()):
                p.apply_update(self.optimizer, self.world_size, None, skip_allreduce=True)
        self.launches_per_replay = int(self.env.lib.track2d_launch_count() - l0)  # libtrack2d kernels inside one replay
        self.optimizer.step_count, self.n_iter, self.player.n_steps = step0, iter0, nsteps0
        self._graph_mode = mode
        return self

    def replay(self):
        """one captured iteration; returns the same tensors as iteration() (overwritten by the next replay)"""
        self.optimizer.advance_for_replay()
        self._graph.replay()
        if self._graph_apply is not None:
            self.allreduce(self.optimizer.fp.grad)
            self._graph_apply.replay()
        self.n_iter += 1
        self.player.n_steps += self.args.num_steps * self.args.num_envs
        return self._graph_out

    def env_steps_per_iteration(self):
        return self.args.num_steps * self.args.num_envs

    def save(self, path):
        torch.save({k: v.detach().clone() for k, v in self.model.state_dict().items()}, path)  # (views of one flat buffer: clone)


def scheduled_mode(args, n_iter):
    """test.py:84-91: until `init_step` updates only the tracker learns (mode 0), afterwards --train-mode"""
    return 0 if n_iter < args.init_step else args.train_mode


def main():
    """main.py:54-116 + test.py:56-134 for the synchronous learner: train, evaluate on --env-base every --eval-every updates, keep
    the best tracker (all-best-{n}.dat / all-new.dat and, with --split, tracker-/target-{best,new}.dat under --log-dir/<env>/<date>),
    follow the init_step training-mode schedule, stop after --max-step updates or --iters iterations.  An "update" here is one
    optimizer step on the mean gradient of num_envs x num_steps env-steps (n_iter of the reference counts optimize() calls too)."""
    from datetime import datetime
    from .gym_eval import Evaluator
    args = make_parser().parse_args()
    if args.fp32_emulation and not blas.status()["enabled"]:
        print("note: --fp32-emulation needs `blas.enable_fp32_emulation()` before torch is imported; run through "
              "`python -c 'from active_tracking_rl_b200 import blas; blas.enable_fp32_emulation(); "
              "from active_tracking_rl_b200 import train; train.main()' ...`")
    args.single = False
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dev = 'cuda:%d' % local_rank
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(dev))
    tr = Trainer(args, dev, rank, world)
    ev = None
    if rank == 0:
        log_dir = os.path.join(args.log_dir, args.env, datetime.now().strftime('%b%d_%H-%M'))  # main.py:97-98
        ev = Evaluator(args, log_dir, dev)

    def evaluate():
        if ev is None:
            return None
        st = ev.run(tr.model, tr.n_iter)
        print('eval @ %d updates (%s, %d episodes): R_mean %.3f  EL_mean %.1f  S_rate %.3f  -> %s' % (
            tr.n_iter, st['Env'], st['episodes'], st['R_mean'], st['EL_mean'], st['S_rate'], os.path.join(ev.log_dir, st['checkpoint'])), flush=True)
        return st

    t0 = time.time()
    graph_mode = None
    for it in range(args.iters):
        mode = scheduled_mode(args, tr.n_iter)
        if args.graph:
            if graph_mode != mode:
                tr.capture(training_mode=mode, warmup=1)
                graph_mode = mode
            pl, vl, ent, prl = tr.replay()
        else:
            pl, vl, ent, prl = tr.iteration(training_mode=mode)
        if rank == 0 and (it % 10 == 0 or it == args.iters - 1):
            torch.cuda.synchronize()
            sps = (it + 1) * tr.env_steps_per_iteration() * world / (time.time() - t0)
            print('iter %d  mode %d  policy_loss %.4f %.4f  value_loss %.4f %.4f  pred_loss %.4f  reward %.4f  env-steps/s %.3e' % (
                it, mode, pl[:, 0].mean(), pl[:, 1].mean(), vl[:, 0].mean(), vl[:, 1].mean(), prl.mean(), tr.player.rew_buf[:, :, 0].mean(), sps), flush=True)
        if args.eval_every > 0 and tr.n_iter % args.eval_every == 0:
            evaluate()
        if tr.n_iter > args.max_step:  # test.py:130-134
            break
    evaluate()
    if tr.peer is not None and tr.peer.status():  # a peer's signal did not arrive within the bound: the gradient sums cannot be trusted
        raise RuntimeError("rank %d: peer-memory all-reduce timed out waiting for rank %d" % (rank, tr.peer.status() - 1))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
