#!/usr/bin/env python
"""bench.py -- env-steps/sec of the north-star path on Track2D-BlockPartialPZR-v0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the whole hot path over the batch: a 20-step rollout of E envs per GPU
(policy.step <-> env.step on one stream, no host round trip) followed by one update (bootstrap, GAE,
A3C losses of tracker + TAT target + aux reward head, backward, [NCCL all-reduce], fused SharedAdam):
E * 20 env-steps per GPU.  The step is replayed from CUDA graphs (one graph on a single GPU; two graphs around an
eagerly launched all-reduce with several).  The policy computes in float32: fc / LSTM / head GEMMs on the tcgen05
tensor cores with the fp32-accurate 3xTF32 split, conv stack / LSTM cell / GAE / SharedAdam as hand-written kernels.
The JSON line carries:

  value / ms_per_step   whole-job env-steps/s (all GPUs), device-timed with CUDA events, max over ranks
  env_only              the environment alone (step kernel + auto-reset, resident random actions)
  roofline              the step kernel: algorithmic 1,755 B per env-step / its launch duration, against the
                        measured HBM copy bandwidth (MEASURED_PEAKS.json); observations go to a ring of
                        buffers larger than L2 so every store reaches HBM
  e2e                   the same loop with every env.step going through the host-buffer C ABI
                        (track2d_step_host: H2D actions, kernels, D2H obs/reward/done) and the observation copied
                        back for the policy -- the reference's own Agent.action_train data flow
  cpu_baseline          the oracle port of the reference's Hogwild A3C workers on this box's host cores
                        (rank 0, N = 1 only, bounded sample)
  --impl reference      times that CPU implementation as the reference arm
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_ID = "Track2D-BlockPartialPZR-v0"
METRIC = "env-steps/sec Track2D-BlockPartialPZR-v0 @64k envs"
BYTES_PER_ENV_STEP = 1755  # 370 B read + 1,385 B written, fp32 observations
NUM_STEPS = 20


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's A3C workers (train.py / main.py:102-116), Hogwild over host cores
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(rank, sd, adam_state, seconds, seed, out_q):
    import torch
    torch.set_num_threads(1)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import a3c_oracle
    import oracle
    env = oracle.OracleEnv(ENV_ID)
    env.seed(seed + rank)
    w = a3c_oracle.Worker(env, sd, adam_state, tat=True, seed=seed + rank)
    w.iteration()  # warm-up (first reset + first update)
    n0, t0 = w.n_steps, time.time()
    while time.time() - t0 < seconds:
        w.iteration()
    out_q.put((w.n_steps - n0, time.time() - t0))


def cpu_a3c_sample(seconds, workers=None, seed=1):
    """W Hogwild workers sharing weights and SharedAdam state through shared memory, like main.py:86-116.
    Returns (env_steps_per_s, workers)."""
    import torch
    import torch.multiprocessing as mp
    os.environ["OMP_NUM_THREADS"] = "1"  # main.py:3
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import a3c_oracle
    import oracle
    oracle.build()
    W = workers or max(1, min(os.cpu_count() or 1, 64))
    torch.manual_seed(seed)
    sd = a3c_oracle.det_state_dict(tat=True, seed=seed, scale=0.05)
    for v in sd.values():
        v.share_memory_()
    adam_state = {}
    for k, v in sd.items():  # SharedAdam.share_memory (shared_optim.py:113-120); the step counter stays per-process
        adam_state[k] = [0, torch.zeros_like(v).share_memory_(), torch.zeros_like(v).share_memory_(), torch.zeros_like(v).share_memory_()]
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(r, sd, adam_state, seconds, seed, q)) for r in range(W)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    return sum(n / t for n, t in res), W


def run_reference_arm(a):
    """--impl reference: the reference's CPU implementation of the path (oracle port; /root/reference is a Python
    tree that cannot travel to the GPU box) on all host cores.  Each step = one bounded sample."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sample_s = 6.0
    for _ in range(a.warmup):
        cpu_a3c_sample(1.0)
    vals, W = [], 0
    t0 = time.time()
    for _ in range(a.steps):
        v, W = cpu_a3c_sample(sample_s)
        vals.append(v)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": (time.time() - t0) / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s, tat-maze-lstm + aux reward, 20-step rollouts, Hogwild A3C, 1 env per worker" % ENV_ID,
                   "parallelism": "%d CPU worker processes" % W},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": W, "kind": "port",
                         "sample": "%d x %.0f s of %d Hogwild A3C workers (oracle port of train.py + gym-track2d)" % (a.steps, sample_s, W)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_gpu_arm(a):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    assert world == a.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        # before CUDA is initialised in this process (the workers are forked)
        v, W = cpu_a3c_sample(a.cpu_seconds)
        cpu_base = {"value": v, "unit": "env-steps/s", "cores": W, "kind": "port",
                    "sample": "%.0f s of %d Hogwild A3C workers (oracle port of train.py + gym-track2d), %s, tat-maze-lstm" % (a.cpu_seconds, W, ENV_ID)}

    emu = None
    if a.fp32_emulation:  # must happen before torch is imported (see active_tracking_rl_b200/blas.py)
        from active_tracking_rl_b200 import blas
        blas.enable_fp32_emulation()
        emu = blas.status()
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version banner there)
    import torch
    import torch.distributed as dist
    from active_tracking_rl_b200 import _lib
    from active_tracking_rl_b200.envs import Track2DVecEnv
    from active_tracking_rl_b200.train import Trainer, default_args

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    lib = _lib.load()
    E, T = a.envs, NUM_STEPS

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- env-only + step-kernel roofline (rank-local; reported by rank 0) -------------------------------
    env_only, roofline = None, None
    if rank == 0:
        peak, peak_src = measured_peak()
        ring = 8  # 8 x 88.6 MB of observations >> 126 MB of L2: every store goes to HBM
        env = Track2DVecEnv(ENV_ID, num_envs=E, device=dev, seed=1, rng="philox", auto_reset=False)
        env.reset()
        g = torch.Generator(device=dev).manual_seed(0)
        acts = [torch.randint(0, 4, (E, 2), generator=g, device=dev, dtype=torch.int32) for _ in range(ring)]
        obs_ring = [torch.empty_like(env.obs) for _ in range(ring)]
        rew, done = torch.empty_like(env.reward), torch.empty_like(env.done)
        for i in range(20):
            env.step_into(acts[i % ring], obs_ring[i % ring], rew, done)
        torch.cuda.synchronize()
        n_l = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_l):
            env.step_into(acts[i % ring], obs_ring[i % ring], rew, done)  # exactly one launch: the step kernel
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / n_l
        achieved = BYTES_PER_ENV_STEP * E / (k_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "step_kernel<learned target, f32 obs, 16 envs/CTA>", "achieved": round(achieved, 1), "peak": peak,
                    "unit": "GB/s", "frac": round(achieved / peak, 4), "peak_source": peak_src, "us_per_launch": round(k_ms * 1e3, 2),
                    "bytes_per_launch": BYTES_PER_ENV_STEP * E, "traffic": a.traffic_bytes,
                    "timing": "CUDA events round %d back-to-back launches on the launching stream (programmatic dependent launch between them), obs ring of %d buffers (> L2)" % (n_l, ring)}
        env.close()
        env = Track2DVecEnv(ENV_ID, num_envs=E, device=dev, seed=1, rng="philox", auto_reset=True)
        env.reset()
        for i in range(100):
            env.step_into(acts[i % ring], obs_ring[i % ring], rew, done)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n_l):
            env.step_into(acts[i % ring], obs_ring[i % ring], rew, done)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_l
        env_only = {"value": E / (ms * 1e-3), "unit": "env-steps/s", "us_per_step": round(ms * 1e3, 2),
                    "what": "step kernel + auto-reset of finished envs, resident uniform-random actions, 1 GPU"}
        env.close()
        del obs_ring, acts
        torch.cuda.empty_cache()

    # ---- the full path --------------------------------------------------------------------------------
    args = default_args(env=ENV_ID, num_envs=E, num_steps=T, seed=1)
    tr = Trainer(args, dev, rank, world)
    # The whole iteration (~3,700 launches) is replayed from CUDA graphs unless --no-graph: ONE graph on a single GPU; with
    # several GPUs two graphs with the NCCL all-reduce launched eagerly between them (capturing the collective itself dead-locked
    # on this image, torch 2.11 / NCCL 2.28).
    step_fn, mode = tr.iteration, "eager launches"
    if not a.no_graph:
        ok = 1
        try:
            tr.capture(warmup=2)
        except Exception as ex:  # noqa: BLE001
            ok = 0
            sys.stderr.write("rank %d: CUDA-graph capture failed (%s)\n" % (rank, str(ex).splitlines()[0] if str(ex) else repr(ex)))
            torch.cuda.synchronize()
        if world > 1:  # every rank must take the same path (a re-created Trainer restarts from the initial weights)
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if ok:
            step_fn = tr.replay
            mode = "one CUDA graph replay per step" if world == 1 else "two CUDA graph replays per step around an eager NCCL all-reduce"
        else:
            if rank == 0:
                sys.stderr.write("running eagerly\n")
            tr = Trainer(args, dev, rank, world)
            step_fn = tr.iteration
    for _ in range(a.warmup):
        step_fn()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = lib.track2d_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step_fn()
    e1.record()
    barrier()
    launches = lib.track2d_launch_count() - l0
    if step_fn == getattr(tr, "replay", None) and hasattr(tr, "launches_per_replay"):
        launches = tr.launches_per_replay * a.steps  # kernels of ours inside the replayed graphs
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / a.steps
    value = E * T * world / (ms_per_step * 1e-3)

    # ---- e2e: every env.step through the host-buffer C ABI ------------------------------------------------
    host = tr.env.alloc_host_buffers(obs_dtype=torch.uint8 if a.e2e_obs == "u8" else torch.float32)
    k2 = max(1, min(a.steps, a.e2e_steps))
    tr.iteration(host=host)
    barrier()
    e0.record()
    for _ in range(k2):
        tr.iteration(host=host)
    e1.record()
    barrier()
    ms2 = max_over_ranks(e0.elapsed_time(e1)) / k2
    obs_b = E * 2 * 169 * 4
    obs_pcie = E * 2 * 169 * (1 if a.e2e_obs == "u8" else 4)
    h2d = T * (E * 2 * 4 + obs_pcie + E * 2 * 4 + E)      # actions into the env; obs, reward, done back to the policy's device
    d2h = T * (E * 2 * 4 + obs_pcie + E * 2 * 4 + E) + 16  # actions out of the policy; obs, reward, done out of the env; loss scalars
    e2e = {"value": E * T * world / (ms2 * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": ms2, "steps": k2,
           "what": "Agent.action_train with env.step via track2d_step_host%s (pinned host buffers, %s observations on the host side) "
                   "and the observation re-uploaded, per GPU" % ("_u8" if a.e2e_obs == "u8" else "", "uint8" if a.e2e_obs == "u8" else "float32")}
    status = tr.env.status()
    replicas_identical = None
    if world > 1:  # every rank applied the same fused update to the same all-reduced gradient: the weights must be bit-identical
        chk = tr.optimizer.fp.flat.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        replicas_identical = bool((lo == hi).item())
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s, %d envs per GPU, tat-maze-lstm tracker+target + aux reward (full AD-VAT), rollout %d steps + 1 update per step"
                               % (ENV_ID, E, T), "envs_per_gpu": E, "rollout_steps": T, "parallelism": "dp%d" % world,
                   "rng": "philox", "l2": "rollout observation buffers (21 x %.0f MB) and the roofline ring exceed the 126 MB L2" % (obs_b / 1e6),
                   "policy_math": "float32; fc / LSTM GEMMs (forward, dgrad, wgrad) on tcgen05 with the 3xTF32 split (track2d_gemm_tf32x3, "
                                  "fp32-accurate: |err| <= 2e-6 |A||B|, tests/test_gpu_gemm.py); conv stack fp32 FFMA2; heads cuBLAS SIMT"
                                  + ("; remaining cuBLAS GEMMs via " + emu["reason"] if emu and emu["enabled"] else ""),
                   "max_grad_norm": args.max_grad_norm, "launch_mode": mode},
        "roofline": roofline, "env_only": env_only, "cpu_baseline": cpu_base, "e2e": e2e, "clocks": clocks,
        "gpu_launches": int(launches), "device_status": status, "replicas_identical": replicas_identical,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU (weak scaling)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-obs", default="u8", choices=["u8", "f32"], help="dtype of the observations in the host buffers of the e2e path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fp32-emulation", action="store_true",
                    help="route the policy's fp32 GEMMs through cuBLAS 12.9 BF16x9 emulation (fp32-accurate, tensor cores); off = SIMT SGEMM")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying one CUDA graph per step")
    ap.add_argument("--traffic-bytes", type=float, default=56198400.0,
                    help="dram__bytes_read.sum + dram__bytes_write.sum per step-kernel launch from the committed ncu capture "
                         "(profiles/ncu_r1_step_kernel_pdl.txt: 24.46 MB read + 31.74 MB written INSIDE the kernel; the 126 MB write-back L2 "
                         "retires the rest of the 88.6 MB of observations after the kernel ends)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_gpu_arm(a)


if __name__ == "__main__":
    main()
