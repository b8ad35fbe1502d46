"""Stub for `tensorboardX` (reference main.py/trainers import it; not installed here)."""


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def add_figure(self, *a, **k):
        pass

    def close(self):
        pass
