"""Stand-in for tensorboardX.SummaryWriter (R/algorithms/gd.py:18,111): records scalars in memory.  Test infrastructure."""


class SummaryWriter:
    def __init__(self, logdir=None, **_):
        self.logdir = logdir
        self.scalars = []

    def add_scalar(self, tag, value, step=None, *a, **k):
        self.scalars.append((tag, float(value), step))

    def flush(self):
        pass

    def close(self):
        pass
