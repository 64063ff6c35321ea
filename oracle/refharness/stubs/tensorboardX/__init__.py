"""Stub: train.py:10 / test.py:10 log scalars through tensorboardX; timing/parity runs drop them.
Like the real SummaryWriter it creates its log directory (test.py:22 opens `<log_dir>/logger` right after)."""
import os


class SummaryWriter(object):
    def __init__(self, logdir=None, *a, **k):
        if logdir:
            os.makedirs(logdir, exist_ok=True)

    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass
