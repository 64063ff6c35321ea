"""DRAM traffic of the step kernel in steady state -- the figure bench.py reports as roofline.traffic.

A single-launch ncu window misses the write-back of the observation stores (the 126 MB L2 retires them after the kernel
ends), so the counters are summed over MANY consecutive launches on the same > L2 observation ring bench.py times, with
ncu told not to flush caches between kernels (--cache-control none): the write-back of launch i then lands inside the
window of launch i+1, and the mean over the steady-state launches is the per-launch traffic.

  run    (on the GPU box):  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none \
                                --clock-control none -k regex:step_kernel --csv --log-file gpurun_out/step_traffic.csv \
                                python tools/measure_step_traffic.py run [E]
  parse  (anywhere):        python tools/measure_step_traffic.py parse gpurun_out/step_traffic.csv profiles/step_traffic_r2.json
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LAUNCHES, RING = 40, 8


def run(E):
    import torch
    from active_tracking_rl_b200.envs import Track2DVecEnv
    dev = "cuda:0"
    env = Track2DVecEnv("Track2D-BlockPartialPZR-v0", num_envs=E, device=dev, seed=1, rng="philox", auto_reset=False)
    env.reset()
    g = torch.Generator(device=dev).manual_seed(0)
    acts = [torch.randint(0, 4, (E, 2), generator=g, device=dev, dtype=torch.int32) for _ in range(RING)]
    obs_ring = [torch.empty_like(env.obs) for _ in range(RING)]
    rew, done = torch.empty_like(env.reward), torch.empty_like(env.done)
    for i in range(LAUNCHES):
        env.step_into(acts[i % RING], obs_ring[i % RING], rew, done)
    torch.cuda.synchronize()
    assert env.status() == 0


def parse(path, out_path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Metric Name" in r and "Kernel Name" in r)
    hdr = rows[hdr_i]
    ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    per = {}
    for r in rows[hdr_i + 1:]:
        if len(r) <= iv or "step_kernel" not in r[ik]:
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "nsecond": 1.0, "msecond": 1e6}.get(unit, 1.0)
        per.setdefault(int(r[iid]), {})[r[im]] = v * scale
    ids = sorted(per)
    steady = ids[len(ids) // 4:]  # drop the first quarter (cold L2, ring not yet cycled)
    rd = [per[i]["dram__bytes_read.sum"] for i in steady]
    wr = [per[i]["dram__bytes_write.sum"] for i in steady]
    out = {
        "kernel": "step_kernel<learned target, f32 obs, 16 envs/CTA>", "launches_captured": len(ids), "launches_averaged": len(steady),
        "dram_bytes_read_per_launch": sum(rd) / len(rd), "dram_bytes_write_per_launch": sum(wr) / len(wr),
        "traffic_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd),
        "first_launch_traffic": per[ids[0]]["dram__bytes_read.sum"] + per[ids[0]]["dram__bytes_write.sum"],
        "ncu_duration_ns_mean": sum(per[i].get("gpu__time_duration.sum", 0.0) for i in steady) / len(steady),
        "how": "ncu --cache-control none --clock-control none, %d consecutive launches on a ring of %d observation buffers (> L2); "
               "mean over the last %d" % (len(ids), RING, len(steady)),
        "source": os.path.basename(path),
    }
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 65536)
    else:
        parse(sys.argv[2], sys.argv[3])
