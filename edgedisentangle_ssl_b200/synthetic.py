"""Synthetic power-law graphs of the BASELINE shapes: Chung-Lu endpoints with
degree exponent ~2.1, then the reference's own processing (self loops, symmetrise, dedup)
through the CSR builder.  Deterministic for a seed."""
import numpy as np

from .graph import build_adjacency


def power_law_graph(n, m_raw, seed=0, gamma=2.1, max_degree=None):
    """Processed adjacency indices [2, E] (row-major sorted) of a power-law graph.

    m_raw directed draws (src, dst) ~ w x w with w_i ~ (i + i0)^(-1/(gamma-1)); i0 caps the
    expected maximum degree at `max_degree` (default 8*sqrt(n) * m_raw/(13n), ogbn-products
    like: 2.4M nodes / 62M edges -> hubs of ~1e4).  E ~= 2*m_raw + n minus duplicates.
    """
    rng = np.random.RandomState(seed)
    expo = 1.0 / (gamma - 1.0)
    if max_degree is None:
        max_degree = 8.0 * np.sqrt(n) * max(1.0, m_raw / (13.0 * n))
    ranks = np.arange(n, dtype=np.float64)

    def top_share(i0):
        w = (ranks + i0) ** (-expo)
        return w[0] / w.sum()

    target = min(0.5, max_degree / (2.0 * m_raw))
    lo, hi = 1.0, float(n)
    for _ in range(60):
        mid = np.sqrt(lo * hi)
        if top_share(mid) > target:
            lo = mid
        else:
            hi = mid
    w = (ranks + hi) ** (-expo)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    relabel = rng.permutation(n)
    src = relabel[np.searchsorted(cdf, rng.random_sample(m_raw))]
    dst = relabel[np.searchsorted(cdf, rng.random_sample(m_raw))]
    idx, _ = build_adjacency(n, dst, src)
    return idx
