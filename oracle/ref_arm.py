"""The reference's own CPU implementation of the hot path, timed (bench.py `--impl reference` and
the `cpu_baseline` leg).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Drives the UNMODIFIED reference from oracle/_ref/ (oracle/make_ref.py): `models.DISGAT.get_em`
(models.py:217-252) -> `layers.DisGALayer.forward_sparse` (layers.py:340-416) -> `utils.sp_softmax`
/ `sp_matmul` (utils.py:192-207) + `layers.FuseLayer` (layers.py:876-921), forward + backward of a
scalar loss, train mode, on a torch sparse COO adjacency built WITHOUT the reference's dense N x N
`load_data` (data_load.py:44-77 cannot load a synthetic graph).  The graph
comes from numpy + oracle.graph, so this module never imports edgedisentangle_ssl_b200 and no
libedis.so is mapped into the process.  If oracle/_ref/ is absent the oracle port
(oracle/disgat.py) is timed instead and the line says kind "port".
"""
import os
import sys
import time
import types

import numpy as np
import torch

from . import disgat as od
from . import graph as og

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def have_ref():
    return os.path.exists(os.path.join(REF, "MANIFEST.json"))


def load_ref():
    """Import the reference's `layers`, `models`, `utils` from oracle/_ref (unmodified files)."""
    for p in (os.path.join(REF, "_stubs"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.dont_write_bytecode = True
    import layers as ref_layers    # noqa: E402  (the reference's flat modules)
    import models as ref_models    # noqa: E402
    import utils as ref_utils      # noqa: E402
    for m in (ref_layers, ref_models, ref_utils):
        assert os.path.dirname(os.path.abspath(m.__file__)) == REF, "reference module shadowed: %s" % m.__file__
    return ref_layers, ref_models, ref_utils


def power_law_coo(n, m_raw, seed=0, gamma=2.1):
    """Processed adjacency [2, E] of a Chung-Lu power-law graph: the same law as the
    GPU arm's generator (degree exponent 2.1, expected max degree 8 sqrt(n) m/(13 n)), restated in
    numpy; self loops + symmetrise + dedup through oracle.graph.build_adjacency."""
    rng = np.random.RandomState(seed)
    expo = 1.0 / (gamma - 1.0)
    max_degree = 8.0 * np.sqrt(n) * max(1.0, m_raw / (13.0 * n))
    ranks = np.arange(n, dtype=np.float64)
    target = min(0.5, max_degree / (2.0 * m_raw))
    lo, hi = 1.0, float(n)
    for _ in range(60):
        mid = np.sqrt(lo * hi)
        w = (ranks + mid) ** (-expo)
        if w[0] / w.sum() > target:
            lo = mid
        else:
            hi = mid
    w = (ranks + hi) ** (-expo)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    relabel = rng.permutation(n)
    src = relabel[np.searchsorted(cdf, rng.random_sample(m_raw))]
    dst = relabel[np.searchsorted(cdf, rng.random_sample(m_raw))]
    idx, val = og.build_adjacency(n, dst, src)
    return idx, val


def _ref_args(ref_utils, att, gnn, C, D, dropout):
    a = ref_utils.get_parser().parse_args(["--model=DISGAT", "--sparse", "--att=%d" % att, "--gnn_type=" + gnn,
                                           "--nhead=%d" % C, "--nhid=%d" % D, "--dropout=%g" % dropout, "--no-cuda"])
    a.cuda = False
    return a


class Workload:
    """One bounded sample of the bench workload: graph + model + inputs, reused across steps."""

    def __init__(self, n, m_raw, feat, C, D, att, gnn, dropout, threads, seed=1):
        torch.set_num_threads(threads)
        self.threads = threads
        idx, val = power_law_coo(n, m_raw, seed=seed)
        self.n, self.e = n, int(idx.shape[1])
        gen = torch.Generator().manual_seed(0)
        self.x = torch.randn(n, feat, generator=gen)
        self.R = torch.randn(n, D, generator=gen)
        self.kind = "reference" if have_ref() else "port"
        if self.kind == "reference":
            L, M, U = load_ref()
            args = _ref_args(U, att, gnn, C, D, dropout)
            torch.manual_seed(4)
            self.enc = M.DISGAT(args, nfeat=feat, nhid=D, nclass=D, nheads=C, dropout=dropout)
            self.fus = [L.FuseLayer(args, C, nfeat=D), L.FuseLayer(args, C, nfeat=D)]
            self.enc.train()
            # data_load.py:158-165: COO in CSR order, float32 values; coalesced once like adj.cuda()
            self.adj = torch.sparse_coo_tensor(torch.from_numpy(idx), torch.from_numpy(val), (n, n)).coalesce()
            self.params = [p for m in [self.enc] + self.fus for p in m.parameters()]
        else:
            self.idx = torch.from_numpy(idx)
            self.att, self.gnn, self.C, self.dropout = att, gnn, C, dropout
            p = {}
            for layer, fin in ((1, feat), (2, D)):
                for c in range(C):
                    pre = "attention%d_%d." % (layer, c)
                    p[pre + "W"] = (torch.randn((2 * fin if att == 3 else fin), D, generator=gen) * 0.1).requires_grad_(True)
                    p[pre + "a"] = (torch.randn((D if att == 3 else 2 * D), 1, generator=gen) * 0.1).requires_grad_(True)
                    if gnn == "AT":
                        p[pre + "W_em"] = (torch.randn(fin, D, generator=gen) * 0.1).requires_grad_(True)
                    elif gnn == "GCN":
                        p[pre + "ag_layer.weight"] = (torch.randn(fin, D, generator=gen) * 0.1).requires_grad_(True)
                        p[pre + "ag_layer.bias"] = torch.zeros(D, requires_grad=True)
                    else:
                        p[pre + "ag_layer.proj.weight"] = (torch.randn(D, 2 * fin, generator=gen) * 0.1).requires_grad_(True)
            self.p = p
            self.fusp = [{"fuse.weight": (torch.randn(D, C * D, generator=gen) * 0.05).requires_grad_(True),
                          "fuse.bias": torch.zeros(D, requires_grad=True)} for _ in range(2)]
            self.params = list(p.values()) + [w for f in self.fusp for w in f.values()]

    def step(self):
        """One DISGAT.get_em forward + backward; returns seconds."""
        t0 = time.perf_counter()
        if self.kind == "reference":
            feats = self.enc.get_em(self.x, self.adj, self.fus)
        else:
            feats = od.disgat_traverse(self.p, self.fusp, self.x, self.idx, self.C, self.att, self.gnn,
                                       dropout=self.dropout, training=True)["feats"]
        loss = (feats[-1] * self.R).sum()
        loss.backward()
        for p in self.params:
            p.grad = None
        return time.perf_counter() - t0

    def describe(self):
        what = ("unmodified reference (oracle/_ref: models.DISGAT.get_em -> layers.DisGALayer.forward_sparse), "
                if self.kind == "reference" else "oracle port (oracle/disgat.py), ")
        return what + "power-law sample n=%d E=%d" % (self.n, self.e)


def sized_workload(feat, C, D, att, gnn, dropout, threads, budget_s, start_edges=2_000_000, max_edges=64_000_000,
                   deg=26.3, probe_edges=250_000):
    """Bounded sample of the bench workload for one CPU step of about `budget_s` seconds.

    BASELINE.md section 4: E = 2 M, doubling while a step is predicted to fit the budget and the
    [E, 2F] temporaries fit host RAM.  The prediction comes from one timed step on a `probe_edges` graph
    (the rate per edge is flat in E: the path is a chain of streaming index / scatter ops).  When even
    2 M edges do not fit the budget (few host cores, or the driver's 20 + 5 steps in a few minutes) the
    sample shrinks in steps of `probe_edges` instead -- the line states the E it ran.
    Returns (Workload, seconds of the probe-predicted step)."""
    probe = Workload(int(probe_edges / deg), int(probe_edges * 0.4853), feat, C, D, att, gnn, dropout, threads)
    probe.step()                                   # first call: allocator / thread-pool warm-up
    per_edge = probe.step() / probe.e
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 8 << 30
    # att=3 gathers [E, 2F] fp32 per channel and autograd keeps ~3 such temporaries per channel alive
    e_mem = int(0.5 * avail / ((2 * feat + 3 * D) * 4 * 3 * max(C // 4, 1)))
    e = start_edges
    if per_edge * e > budget_s:
        e = max(probe_edges, int(budget_s / per_edge / probe_edges) * probe_edges)
    else:
        while per_edge * e * 2 <= budget_s and e * 2 <= max_edges:
            e *= 2
    e = max(probe_edges, min(e, e_mem))
    if e == probe.e or abs(e - probe.e) < probe_edges // 2:
        return probe, per_edge * probe.e
    del probe
    return Workload(int(e / deg), int(e * 0.4853), feat, C, D, att, gnn, dropout, threads), per_edge * e
