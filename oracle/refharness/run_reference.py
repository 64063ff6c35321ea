"""Run a command-line entry point of the UNMODIFIED reference (main.py, gym_eval.py) under the import stubs.
TEST INFRASTRUCTURE ONLY -- usable only where /root/reference exists (the build container).

    python oracle/refharness/run_reference.py main.py --shared-optimizer --workers 6 --split --train-mode -1 \
        --env Track2D-BlockPartialPZR-v0 --log-dir /tmp/ref_train/ --max-step 150000
    python oracle/refharness/run_reference.py gym_eval.py --env Track2D-BlockPartialNav-v0 --network tat-maze-lstm \
        --load-tracker /tmp/ref_train/.../tracker-best.dat --num-episodes 100 --csv /tmp/ref_eval.csv

The script is executed with runpy as `__main__`, from a scratch working directory (the reference writes logs/ and
CSV files relative to cwd; /root/reference is read-only).  np.random.seed() keeps its real behaviour here unless
T2D_REF_NEUTRALISE_RESEED=1: this runner is for training / evaluating the reference as shipped, not for fixtures.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref  # noqa: E402


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = sys.argv[1]
    neutralise = os.environ.get("T2D_REF_NEUTRALISE_RESEED", "0") == "1"
    ref.load_reference(neutralise_reseed=neutralise)
    path = os.path.join(ref.REFERENCE_ROOT, script)
    if not os.path.isfile(path):
        raise SystemExit("no such reference script: %s" % path)
    work = os.environ.get("T2D_REF_WORKDIR", "/tmp/track2d_ref_work")
    os.makedirs(work, exist_ok=True)
    os.chdir(work)
    sys.argv = [path] + sys.argv[2:]
    runpy.run_path(path, run_name="__main__")


if __name__ == "__main__":
    main()
