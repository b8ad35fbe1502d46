"""Stub for the `ipdb` import the reference does at module top (not installed here)."""


def set_trace(*a, **k):
    raise RuntimeError("ipdb.set_trace() reached in the reference")
