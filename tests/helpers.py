"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance for fp32 results: 1e-5 relative.  Measured against the tensor's
# scale (max |ref|) for near-zero entries, as usual for reassociated fp32 sums.
RTOL = 1e-5


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def t(a, dtype=None):
    x = torch.from_numpy(np.asarray(a))
    return x.to(dtype) if dtype is not None else x


def _np(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def rel_err(got, ref, floor=0.0):
    got, ref = _np(got), _np(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    scale = max(np.abs(ref).max(), floor, 1e-30)
    return float(np.abs(got - ref).max() / scale)


def assert_close(got, ref, rtol=RTOL, what="", floor=0.0):
    """max |got - ref| <= rtol * max(max |ref|, floor).

    `floor` is used for gradient groups: a tensor whose gradient is orders of magnitude below
    the group's largest one (cancelling sums; the reference's own fp32 result is noise there)
    is measured against GRAD_FLOOR x the group's scale instead of its own."""
    err = rel_err(got, ref, floor)
    assert err <= rtol, "%s: rel err %.3e > %.1e" % (what, err, rtol)


GRAD_FLOOR = 1e-3


def group_floor(refs):
    """GRAD_FLOOR x the largest |value| over a group of reference tensors."""
    m = 0.0
    for r in refs:
        r = _np(r)
        if r.size:
            m = max(m, float(np.abs(r).max()))
    return GRAD_FLOOR * m


def params_from(g, prefix):
    """Collect `prefix + name` entries of a golden dict as float32 torch tensors."""
    return {k[len(prefix):]: t(v) for k, v in g.items() if k.startswith(prefix)}
