"""Recipe: make the reference's CPU implementation of the path available to bench.py's reference arm on the GPU box.

    python oracle/refharness/populate_ref.py          (run by __graft_entry__.build() where /root/reference exists)

/root/reference is a Python tree, there is nothing to compile: the files of the 2D path are copied VERBATIM from where they lie
into oracle/_ref/ (git-ignored: never part of the repo's history or of the product; NOT gpurun-ignored, so it travels to the
GPU box like a built .so).  bench.py --impl reference and its cpu_baseline leg then time the UNMODIFIED train.py / player_util.py
/ model.py / gym_track2d through oracle/refharness/ref_bench.py (import stubs for the packages this image lacks).  Without
oracle/_ref those legs fall back to the oracle port and say so (kind: "port").
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(os.path.dirname(HERE), "_ref")
FILES = ["environment.py", "main.py", "model.py", "perception.py", "player_util.py", "shared_optim.py", "test.py", "train.py", "utils.py",
         "gym_eval.py", "LICENSE"]


def populate(src="/root/reference", dst=DST):
    if not os.path.isfile(os.path.join(src, "train.py")):
        return None
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copy2(os.path.join(src, f), os.path.join(dst, f))
    pkg_src = os.path.join(src, "envs", "gym-track2d", "gym_track2d")
    pkg_dst = os.path.join(dst, "envs", "gym-track2d", "gym_track2d")
    if os.path.isdir(pkg_dst):
        shutil.rmtree(pkg_dst)
    shutil.copytree(pkg_src, pkg_dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return dst


if __name__ == "__main__":
    out = populate(*(sys.argv[1:2] or ["/root/reference"]))
    print(out or "reference tree not found: nothing copied")
