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


def make_problem(gnn="AT"):
    import edgedisentangle_ssl_b200 as edis
    from edgedisentangle_ssl_b200.utils import get_parser
    rng = np.random.RandomState(3)
    n, fin = 90, 12
    idx, _ = og.build_adjacency(n, rng.randint(0, n, 500), rng.randint(0, n, 500))
    args = get_parser().parse_args(["--model=DISGAT", "--sparse", "--att=3", "--gnn_type=" + gnn, "--nhead=2",
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


def oracle_kernels(part):
    """(fwd, bwd) on raw operands for parallel.PartitionedLayer, computed by plain torch on the CPU
    (layers.py:375-379, 392-399 on P_i + Q_j): stands in for libedis in the gloo tests."""
    row = torch.from_numpy(part.row_local)
    col = torch.from_numpy(part.col_local)

    def forward(P, QV, a, bias, C, D):
        CD = C * D
        z = P[row] + QV[col, :CD]
        e = (torch.nn.functional.leaky_relu(z, 0.01).reshape(-1, C, D) * a.reshape(1, C, D)).sum(-1)       # [E, C]
        w = torch.exp(torch.sigmoid(e))
        den = torch.zeros(part.n_local, C).index_add_(0, row, w)
        alpha = w / den[row]
        msg = alpha.unsqueeze(-1) * QV[col, CD:].reshape(-1, C, D)
        agg = torch.zeros(part.n_local, C, D).index_add_(0, row, msg).reshape(-1, CD)
        if bias is not None:
            agg = agg + bias
        return torch.nn.functional.elu(agg), e

    def fwd(graph, d, P, QV, a, bias, want_sign):
        with torch.no_grad():
            out, e = forward(P, QV, a, bias, d.C, d.D)
        return out, None, e, None, None

    def bwd(graph, d, P, QV, a, bias, saved, g_out, g_edge_e):
        leaves = [t.detach().requires_grad_(True) for t in (P, QV, a)]
        b = bias.detach().requires_grad_(True) if bias is not None else None
        with torch.enable_grad():
            out, e = forward(leaves[0], leaves[1], leaves[2], b, d.C, d.D)
            tot = (out * g_out).sum() + (0 if g_edge_e is None else (e * g_edge_e).sum())
        gs = torch.autograd.grad(tot, leaves + ([b] if b is not None else []))
        return gs[0], gs[1], gs[2], (gs[3] if b is not None else None)

    return fwd, bwd


def oracle_agg_kernels(part):
    """(fwd, bwd) of the shared-operand (aggregate-then-project) kernels for parallel.PartitionedAggLayer, in
    plain torch: agg[i, c, :] = sum_j alpha^c_ij x_j with alpha from a . lrelu(P_i + Q_j)."""
    row = torch.from_numpy(part.row_local)
    col = torch.from_numpy(part.col_local)

    def forward(P, Q, X, a, C, D):
        z = P[row] + Q[col]
        e = (torch.nn.functional.leaky_relu(z, 0.01).reshape(-1, C, D) * a.reshape(1, C, D)).sum(-1)
        w = torch.exp(torch.sigmoid(e))
        alpha = w / torch.zeros(part.n_local, C).index_add_(0, row, w)[row]
        msg = alpha.unsqueeze(-1) * X[col].unsqueeze(1)                               # [E, C, F]
        agg = torch.zeros(part.n_local, C, X.shape[1]).index_add_(0, row, msg)
        return agg.reshape(part.n_local, -1), e

    def fwd(graph, d, P, Q, X, a, want_sign):
        with torch.no_grad():
            agg, e = forward(P, Q, X, a, d.C, d.D)
        return agg, e, None, None

    def bwd(graph, d, P, Q, X, a, saved, g_agg, g_edge_e, need_gx):
        leaves = [t.detach().requires_grad_(True) for t in (P, Q, a, X)]
        with torch.enable_grad():
            agg, e = forward(leaves[0], leaves[1], leaves[3], leaves[2], d.C, d.D)
            tot = (agg * g_agg).sum() + (0 if g_edge_e is None else (e * g_edge_e).sum())
        gP, gQ, ga, gX = torch.autograd.grad(tot, leaves)
        return gP, gQ, ga, (gX if need_gx else None)

    return fwd, bwd


def layer_worker(rank, world, port, out_q, mode, gnn, agg=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), EDIS_EXCHANGE=mode)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from edgedisentangle_ssl_b200 import parallel as par
        n, idx, args, enc, fus, x, R = make_problem(gnn)
        part = par.partition_of_global_graph(idx, n, rank, world)
        assert part.mode == mode
        x_loc = x[part.lo:part.hi].clone().requires_grad_(True)
        feats = par.get_em_partitioned(enc, fus, x_loc, part, kernels=oracle_kernels(part),
                                       agg_kernels=oracle_agg_kernels(part) if agg else None)
        (feats[-1] * R[part.lo:part.hi]).sum().backward()
        params = [p for m in [enc] + fus for p in m.parameters()]
        par.allreduce_grads(params)
        out_q.put((rank, {"feat": feats[-1].detach().numpy().copy(), "lo": part.lo, "hi": part.hi,
                          "gx": x_loc.grad.numpy().copy(),
                          "grads": {k: v.grad.numpy().copy() for k, v in enc.named_parameters() if v.grad is not None}}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,gnn,world,agg", [("allgather", "AT", 2, False), ("alltoall", "AT", 2, False),
                                                ("allgather", "GCN", 3, False), ("allgather", "AT", 2, True),
                                                ("alltoall", "GCN", 2, True)])
def test_partitioned_layer_node_matches_single_process(mode, gnn, world, agg):
    """parallel.PartitionedLayer (own-row P, own+halo Q|V, overlapped exchange, hand-ordered backward)
    through both exchange collectives: features, INPUT gradient and weight gradients equal the
    single-process run of the same oracle kernels.  agg=True: layer 2 (F == D) runs as
    parallel.PartitionedAggLayer (aggregate-then-project: only Q projected for halo rows, W_em on own rows);
    the single-process comparison always uses the project-then-aggregate node, so the two plans are also
    checked against each other."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=layer_worker, args=(r, world, port, q, mode, gnn, agg)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from edgedisentangle_ssl_b200 import parallel as par
    n, idx, args, enc, fus, x, R = make_problem(gnn)
    part = par.Partition(0, 1, np.array([0, n]), idx[0], idx[1])
    xg = x.clone().requires_grad_(True)
    feats = par.get_em_partitioned(enc, fus, xg, part, kernels=oracle_kernels(part))
    (feats[-1] * R).sum().backward()
    for r in range(world):
        got = results[r]
        assert torch.allclose(torch.from_numpy(got["feat"]), feats[-1][got["lo"]:got["hi"]].detach(), rtol=1e-5, atol=1e-6)
        assert torch.allclose(torch.from_numpy(got["gx"]), xg.grad[got["lo"]:got["hi"]], rtol=2e-4, atol=1e-6)
        for k, v in enc.named_parameters():
            if v.grad is not None:
                assert torch.allclose(torch.from_numpy(got["grads"][k]), v.grad, rtol=2e-4, atol=1e-6), k


def test_row_ranges_never_empty_with_a_dominant_hub():
    from edgedisentangle_ssl_b200 import parallel as par
    deg = np.array([1, 1, 1000, 1, 1, 1, 1, 1])
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    for world in (2, 4, 8):
        b = par.row_ranges(rowptr, world)
        assert b[0] == 0 and b[-1] == 8 and np.all(np.diff(b) >= 1), b


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


# ------------------------------------------------------------------ partitioned SSL pair loss
def torch_pair_fn(att, C, D, pi, pj, c_lo, c_hi, P, Q, a, plist=None):
    """layers.py:349-389 on pair lists, plain torch (stands in for the CUDA PairScore on CPU)."""
    if att == 1:
        e = P[pi] + Q[pj]
    elif att == 2:
        e = (P[pi].reshape(-1, C, D) * Q[pj].reshape(-1, C, D)).sum(-1)
    else:
        e = (torch.nn.functional.leaky_relu(P[pi] + Q[pj], 0.01).reshape(-1, C, D) * a.reshape(1, C, D)).sum(-1)
    return e[:, c_lo:c_hi]


def torch_wmse(scores, target, n_pos, m_total):
    """utils.adj_mse_loss (utils.py:287-298) on one slice of a pair set of m_total pairs."""
    p = torch.sigmoid(scores.sum(1))
    w_neg = n_pos / (float(m_total) ** 2 - n_pos)
    w = torch.where(target != 0, torch.ones_like(p), torch.full_like(p, w_neg))
    return (w * (p - target) ** 2).sum() / m_total


def ssl_problem():
    n, idx, args, enc, fus, x, R = make_problem()
    rng = np.random.RandomState(11)
    key = np.unique(np.concatenate([rng.randint(0, n * n, 900), idx[0][::3] * n + idx[1][::3]]))
    pos = set((idx[0] * n + idx[1]).tolist())
    lab = np.array([1.0 if k in pos else 0.0 for k in key], dtype=np.float32)
    return n, idx, enc, fus, x, key, lab


def ssl_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from edgedisentangle_ssl_b200 import parallel as par
        n, idx, enc, fus, x, key, lab = ssl_problem()
        rowptr = np.concatenate([[0], np.cumsum(np.bincount(idx[0], minlength=n))])
        bounds = par.row_ranges(rowptr, world, balance="edges")
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        sel = (idx[0] >= lo) & (idx[0] < hi)
        part = par.Partition(rank, world, bounds, idx[0][sel], idx[1][sel])
        mine = (key // n >= lo) & (key // n < hi)
        pairs = torch.from_numpy(np.stack([key[mine] // n - lo, key[mine] % n]))
        labels = torch.from_numpy(lab[mine])

        def layer_fn(chs, x_need, graph):
            return oracle_layer(chs, x_need, part.row_local, part.col_local, part.n_local)

        loss = par.ssl_pair_loss_partitioned(enc, fus, x[lo:hi], part, [pairs], [labels], [(0, 2)],
                                             [int(lab.sum())], [len(key)], 0, layer_fn, torch_pair_fn, torch_wmse)
        loss.backward()
        par.allreduce_grads([p for m in [enc] + fus for p in m.parameters()])
        out_q.put((rank, {"loss": float(loss.detach()), "m": int(mine.sum()),
                          "grads": {k: v.grad.numpy().copy() for k, v in enc.named_parameters() if v.grad is not None}}))
    finally:
        dist.destroy_process_group()


def test_partitioned_ssl_pair_loss_matches_single_process():
    """Pairs partitioned by row, Q from the all-gathered layer input, reduce-scatter in backward:
    the rank losses add up to the single-process loss and the all-reduced gradients match."""
    from edgedisentangle_ssl_b200 import parallel as par
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=ssl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, idx, enc, fus, x, key, lab = ssl_problem()
    part = par.Partition(0, 1, np.array([0, n]), idx[0], idx[1])

    def layer_fn(chs, x_need, graph):
        return oracle_layer(chs, x_need, part.row_local, part.col_local, part.n_local)

    pairs = torch.from_numpy(np.stack([key // n, key % n]))
    loss = par.ssl_pair_loss_partitioned(enc, fus, x, part, [pairs], [torch.from_numpy(lab)], [(0, 2)],
                                         [int(lab.sum())], [len(key)], 0, layer_fn, torch_pair_fn, torch_wmse)
    loss.backward()
    assert sum(r["m"] for r in results.values()) == len(key) and min(r["m"] for r in results.values()) > 0
    ref_loss = float(loss.detach())
    assert abs(sum(r["loss"] for r in results.values()) - ref_loss) <= 1e-6 * abs(ref_loss) + 1e-9
    seen = 0
    for k, v in enc.named_parameters():
        if v.grad is not None:
            for r in range(world):
                assert torch.allclose(torch.from_numpy(results[r]["grads"][k]), v.grad, rtol=2e-4, atol=1e-7), k
            seen += 1
    assert seen > 0


def test_partition_sampler_rows_and_labels():
    """sample_pairs_partitioned (world 1 here): rows local, columns global, labels = membership in
    the positive set, a third of the positives forced in (pretrainer.py:696-700)."""
    from edgedisentangle_ssl_b200 import parallel as par      # the sampler is device-agnostic torch code
    rng = np.random.RandomState(5)
    n = 300
    idx, _ = og.build_adjacency(n, rng.randint(0, n, 2500), rng.randint(0, n, 2500))
    part = par.Partition(0, 1, np.array([0, n]), idx[0], idx[1])
    pos_key = torch.from_numpy(idx[0] * n + idx[1])
    gen = torch.Generator().manual_seed(3)
    pairs, lab, n_pos, m = par.sample_pairs_partitioned(part, pos_key, gen)
    assert pairs.shape[0] == 2 and pairs.shape[1] == lab.numel() == m
    key = pairs[0] * n + pairs[1]
    assert torch.all(key[1:] > key[:-1])                                   # row-major sorted, deduplicated
    member = torch.isin(key, pos_key)
    assert torch.equal(member, lab != 0) and n_pos == int(member.sum())
    assert n_pos >= pos_key.numel() // 3                                    # the forced third is inside
    expect = 3.0 * pos_key.numel() + pos_key.numel() / 3.0                  # E[M] ~ 3E + E/3 (minus overlap)
    assert 0.8 * expect < m < 1.1 * expect


# ------------------------------------------------------------------ bench.py: sharing the generated graph between ranks
def graph_share_worker(rank, world, port, out_q, cache_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import argparse
        import importlib.util
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
        a = argparse.Namespace(nodes=3000, raw_edges=20000, max_chunk=0, cache_dir=cache_dir)
        idx, hit = bench.global_graph_indices(a, rank, world, "cpu")
        idx2, hit2 = bench.global_graph_indices(a, rank, world, "cpu")          # second call: cache hit if writable
        out_q.put((rank, (np.asarray(idx).copy(), bool(hit), np.asarray(idx2).copy(), bool(hit2))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("writable", [True, False])
def test_bench_graph_is_shared_consistently_between_ranks(tmp_path, writable):
    """bench.global_graph_indices: rank 0 generates and publishes the graph, the other ranks wait on a broadcast
    and map the file -- or, when the cache directory cannot be written, every rank generates it from the same
    seed.  Both ranks must end up with identical arrays and take identical collective branches (no hang)."""
    cache = str(tmp_path / "cache") if writable else "/proc/edis-not-writable"
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=graph_share_worker, args=(r, world, port, q, cache)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][0], res[0][2])
    assert res[0][1] is False and res[1][1] is False
    assert res[0][3] == res[1][3] == writable
