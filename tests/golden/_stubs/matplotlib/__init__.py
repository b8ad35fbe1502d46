"""Stub for `matplotlib` (reference utils.py imports it for plots that are out of scope)."""


def use(*a, **k):
    pass
