"""Streaming, bit-exact SSL pair sampler (host side).

Replaces `SupEdgeTrainer.sample_train` (pretrainer.py:683-707) and the per-label-set body of
`GeneratedEdgeTrainer.sample_train` (pretrainer.py:552-574) without any N x N tensor: the
Bernoulli mask is replayed from the SAME torch CPU generator in row chunks (torch.rand fills
serially, so chunked draws consume the stream exactly like one `torch.rand(N, N)`), the
positives come from the CSR edge list, and the merge / dedup / labelling is integer work in
libedis.so (`edis_merge_pairs_host`).  Memory O(chunk_rows * N); RNG work stays O(N^2), which
is what bit-exactness with the reference's definition costs.
"""
from ctypes import c_float, c_int64

import numpy as np
import torch

from ._lib import check, lib, np_ptr


def homo_hetero_split(indices, labels):
    """DisEdge label sets as edge lists (replaces the dense masks of pretrainer.py:440-456)."""
    indices = np.asarray(indices)
    labels = np.asarray(labels)
    same = labels[indices[0]] == labels[indices[1]]
    return np.ascontiguousarray(indices[:, same]), np.ascontiguousarray(indices[:, ~same])


def bernoulli_hits(n_draws, thr):
    """Sorted linear indices k with u_k < thr, where u = torch.rand(n_draws) on torch's CPU default
    generator -- same hits and same generator state afterwards as drawing the uniforms (bit-exact),
    but the MT19937 stream is replayed in libedis.so and only the hits are stored
    (`edis_rand_hits_host`): ~4x faster than torch.rand + compare + nonzero and O(hits) memory."""
    st = torch.get_rng_state()
    if st.numel() != 5056:                       # unknown generator layout: do it the slow way
        return None
    thr32 = float(np.float32(thr))               # `tensor(float32) < python float` compares in float32
    thr24 = int(min(max(np.ceil(thr32 * 16777216.0), 0), 16777216))
    buf = st.numpy().copy()
    cap = int(n_draws * thr32 * 1.05) + 4096
    while True:
        out = np.empty(cap, dtype=np.int64)
        work = buf.copy()
        m = lib.edis_rand_hits_host(work.ctypes.data, work.nbytes, int(n_draws), thr24, np_ptr(out, c_int64), cap)
        if m >= 0:
            break
        if m > -16:                              # EDIS_ERR_* codes, not a (negative) hit count
            check(m, "edis_rand_hits_host")
        cap = -int(m) + 16
    torch.set_rng_state(torch.from_numpy(work))
    return out[:m]


def sample_pairs(n, pos_indices, chunk_rows=None):
    """One `sample_train` draw for one label set.

    pos_indices: [2, E_L] int64, row-major sorted positive entries (the label matrix's nonzeros).
    Consumes torch's CPU default generator (N*N uniforms) and numpy's global RNG (one shuffle),
    in the reference's order.  Returns (indices[2, M] int64 row-major sorted, label[M] float32).
    """
    pos_indices = np.ascontiguousarray(pos_indices, dtype=np.int64)
    e_l = pos_indices.shape[1]
    # pretrainer.py:690-692: float32 tensor division, then python-double * 3
    thr = (torch.tensor(float(e_l), dtype=torch.float32) / (n * n)).item() * 3
    hit = bernoulli_hits(n * n, thr) if chunk_rows is None else None
    if hit is None:                              # reference path: torch.rand in row chunks
        if chunk_rows is None:
            chunk_rows = max(1, min(n, (1 << 24) // max(n, 1)))
        hits = []
        for r0 in range(0, n, chunk_rows):
            r1 = min(n, r0 + chunk_rows)
            nz = (torch.rand(size=(r1 - r0, n)) < thr).nonzero()
            if nz.numel():
                nz = nz.numpy()
                hits.append((nz[:, 0].astype(np.int64) + r0) * n + nz[:, 1])
        hit = np.concatenate(hits) if hits else np.empty(0, dtype=np.int64)
    pos = np.ascontiguousarray(pos_indices.T)          # [E_L, 2] == adj.nonzero() order (697)
    np.random.shuffle(pos)                             # row shuffle (699)
    forced = pos[: e_l // 3]
    forced_key = np.ascontiguousarray(forced[:, 0] * n + forced[:, 1])
    pos_key = np.ascontiguousarray(pos_indices[0] * n + pos_indices[1])
    if e_l > 1 and not np.all(pos_key[1:] > pos_key[:-1]):
        pos_key = np.unique(pos_key)
    cap = hit.shape[0] + forced_key.shape[0]
    out_key = np.empty(max(cap, 1), dtype=np.int64)
    out_lab = np.empty(max(cap, 1), dtype=np.float32)
    m = lib.edis_merge_pairs_host(hit.shape[0], np_ptr(hit, c_int64), forced_key.shape[0],
                                  np_ptr(forced_key, c_int64), pos_key.shape[0], np_ptr(pos_key, c_int64),
                                  np_ptr(out_key, c_int64), np_ptr(out_lab, c_float))
    check(m, "edis_merge_pairs_host")
    key = out_key[:m]
    return np.stack([key // n, key % n]), out_lab[:m].copy()


def sample_pairs_device(n, pos_key, generator=None, row_range=None, e_total=None):
    """Distribution-equivalent sampler for graphs where N x N uniforms are not an option
    ('infeasible at N = 2.4 M').  Same law as `sample_pairs` -- an iid
    Bernoulli(3*rho) mask over the N x N cells, united with a uniformly random third of the
    positives -- but drawn in O(M): K ~ Binomial(N^2, 3*rho) (normal approximation), K distinct
    cells uniformly at random (draw, dedup, top up), union, sort, label.  Not bit-exact with the
    reference's RNG stream; runs entirely on the device of `pos_key` (sorted int64 keys i*n+j).
    Returns (indices[2, M] int64, label[M] float32) on that device.

    row_range=(lo, hi), e_total: the slice of that sampler owned by one rank of a destination-range
    partition -- cells restricted to rows [lo, hi) (pos_key = this rank's positives, global keys),
    hit probability from the GLOBAL edge count; the union over ranks has the single-process law."""
    dev = pos_key.device
    e_l = pos_key.numel()
    lo, hi = (0, n) if row_range is None else row_range
    cells = float(hi - lo) * float(n)
    thr = (torch.tensor(float(e_l if e_total is None else e_total), dtype=torch.float32) / (n * n)).item() * 3
    mean, var = cells * thr, cells * thr * (1.0 - thr)
    z = torch.randn((), generator=generator, device=dev).item() if generator is not None else torch.randn(()).item()
    k = int(max(0, round(mean + (var ** 0.5) * z)))
    key = torch.empty(0, dtype=torch.int64, device=dev)
    need = k
    while need > 0:
        draw = int(need * 1.05) + 16
        ri = torch.randint(lo, hi, (draw,), device=dev, generator=generator)
        ci = torch.randint(0, n, (draw,), device=dev, generator=generator)
        key = torch.unique(torch.cat([key, ri * n + ci]))
        if key.numel() > k:        # drop a random surplus so that exactly k distinct cells remain
            keep = torch.randperm(key.numel(), device=dev, generator=generator)[:k]
            key = key[keep]
        need = k - key.numel()
    forced = pos_key[torch.randperm(e_l, device=dev, generator=generator)[: e_l // 3]]
    key = torch.unique(torch.cat([key, forced]))
    pos = torch.searchsorted(pos_key, key).clamp_(max=max(e_l - 1, 0))
    label = (pos_key[pos] == key).to(torch.float32) if e_l else torch.zeros_like(key, dtype=torch.float32)
    return torch.stack([torch.div(key, n, rounding_mode="floor"), key % n]), label
