"""CPU tests of the multi-GPU path's host logic with the gloo backend (world_size 2):
destination-range partitioning, compact halo indexing, the halo exchange (forward and its
reverse in backward) and the weight-gradient all-reduce.  The per-rank layer compute is done by
the CPU oracle here (tests may use it); on the GPU box the same code path runs libedis kernels
(tests/test_gpu_parity.py::test_partitioned_matches_full_graph)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import disgat as od
from oracle import graph as og


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_problem():
    import edgedisentangle_ssl_b200 as edis
    from edgedisentangle_ssl_b200.utils import get_parser
    rng = np.random.RandomState(3)
    n, fin = 90, 12
    idx, _ = og.build_adjacency(n, rng.randint(0, n, 500), rng.randint(0, n, 500))
    args = get_parser().parse_args(["--model=DISGAT", "--sparse", "--att=3", "--gnn_type=AT", "--nhead=2",
                                    "--nhid=8", "--dropout=0.0"])
    torch.manual_seed(0)
    enc = edis.DISGAT(args, nfeat=fin, nhid=8, nclass=8, nheads=2, dropout=0.0)
    fus = [edis.FuseLayer(args, 2, nfeat=8), edis.FuseLayer(args, 2, nfeat=8)]
    x = torch.randn(n, fin)
    R = torch.randn(n, 8)
    return n, idx, args, enc, fus, x, R


def oracle_layer(chs, x_need, row, col, n_rows):
    p = {}
    for c, l in enumerate(chs):
        for name, prm in l.named_parameters():
            p["c%d.%s" % (c, name)] = prm
    idx = torch.from_numpy(np.stack([row, col]))
    outs = [od.disga_layer(p, "c%d." % c, x_need, idx, chs[0].att_type, chs[0].gnn_type)[0][:n_rows]
            for c in range(len(chs))]
    return torch.cat(outs, 1)


def worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from edgedisentangle_ssl_b200 import parallel as par
        n, idx, args, enc, fus, x, R = make_problem()
        rowptr = np.concatenate([[0], np.cumsum(np.bincount(idx[0], minlength=n))])
        bounds = par.row_ranges(rowptr, world, balance="edges")
        lo, hi = bounds[rank], bounds[rank + 1]
        sel = (idx[0] >= lo) & (idx[0] < hi)
        part = par.Partition(rank, world, bounds, idx[0][sel], idx[1][sel])
        assert part.n_src == part.n_local + len(part.halo_ids)
        assert part.send_counts.sum() == len(part.send_idx)

        def layer_fn(chs, x_need, graph):
            return oracle_layer(chs, x_need, part.row_local, part.col_local, part.n_local)

        feats = par.get_em_partitioned(enc, fus, x[lo:hi], part, layer_fn)
        loss = (feats[-1] * R[lo:hi]).sum()
        loss.backward()
        params = [p for m in [enc] + fus for p in m.parameters()]
        par.allreduce_grads(params)
        got = {"feat": feats[-1].detach().numpy().copy(), "lo": int(lo), "hi": int(hi),
               "grads": {k: v.grad.numpy().copy() for k, v in enc.named_parameters() if v.grad is not None},
               "halo": len(part.halo_ids)}   # numpy: plain pickling, no shared-memory handles
        out_q.put((rank, got))
    finally:
        dist.destroy_process_group()


def test_partitioned_get_em_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the full graph
    n, idx, args, enc, fus, x, R = make_problem()
    feats = enc_get_em_oracle(enc, fus, x, idx, n)
    (feats[-1] * R).sum().backward()
    for r in range(world):
        got = results[r]
        ref = feats[-1][got["lo"]:got["hi"]]
        assert torch.allclose(torch.from_numpy(got["feat"]), ref.detach(), rtol=1e-5, atol=1e-6)
        assert got["halo"] > 0
        for k, v in enc.named_parameters():
            if v.grad is not None:
                assert torch.allclose(torch.from_numpy(got["grads"][k]), v.grad, rtol=2e-4, atol=1e-6), k


def enc_get_em_oracle(enc, fus, x, idx, n):
    from edgedisentangle_ssl_b200 import parallel as par
    bounds = np.array([0, n])
    part = par.Partition(0, 1, bounds, idx[0], idx[1])

    def layer_fn(chs, x_need, graph):
        return oracle_layer(chs, x_need, part.row_local, part.col_local, part.n_local)

    return par.get_em_partitioned(enc, fus, x, part, layer_fn)


def test_row_ranges_and_compact_indexing():
    from edgedisentangle_ssl_b200 import parallel as par
    rng = np.random.RandomState(0)
    n = 200
    idx, _ = og.build_adjacency(n, rng.randint(0, n, 1500), rng.randint(0, n, 1500))
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(idx[0], minlength=n))])
    for world in (1, 2, 4, 8):
        b = par.row_ranges(rowptr, world)
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
        loads = np.diff(rowptr[b])
        assert loads.max() <= rowptr[-1] / world + np.diff(rowptr).max()     # within one row of balance
    b = par.row_ranges(rowptr, 1)
    part = par.Partition(0, 1, b, idx[0], idx[1])
    assert part.n_src == n and len(part.halo_ids) == 0
    assert np.array_equal(part.col_local, idx[1]) and np.array_equal(part.row_local, idx[0])
