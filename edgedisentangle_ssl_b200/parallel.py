"""Multi-GPU execution of the DISGAT path: destination-range graph partitioning, per-layer halo
exchange of source features and weight-gradient all-reduce (the reference itself is
single-process, so there is no reference behaviour to mirror here).

One process per GPU (`torch.distributed`, NCCL over NVLink; gloo on CPU for the tests).
  * rank r owns the contiguous destination rows [lo_r, hi_r) and their in-edges (CSR slice);
  * the sources its edges read are its own nodes plus a HALO of remote nodes; columns are
    re-indexed compactly (own first, halo after) and the slice becomes a rectangular edis graph;
  * per layer the layer INPUT rows (F or D floats per node, not the 2*C*D projected ones) of the
    halo nodes are fetched from their owners (point-to-point, volume = halo size); projections
    are recomputed locally for own + halo rows;
  * backward returns the halo rows' input gradients to their owners (reverse exchange) and
    all-reduces the weight gradients.
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F


def row_ranges(rowptr, world, balance="edges"):
    """world+1 boundaries of contiguous destination ranges, balanced by in-edge (or node) count."""
    n = len(rowptr) - 1
    if balance == "nodes":
        return np.linspace(0, n, world + 1).round().astype(np.int64)
    target = rowptr[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(rowptr, target, side="left")
    return np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)


def compact_columns(lo, hi, rows_global, cols_global):
    """Local indexing of one destination range: rows -> row - lo; columns -> own nodes first
    (col - lo), then halo nodes in ascending global id.  Returns (row_local, col_local, halo_ids)."""
    rows_global = np.asarray(rows_global, dtype=np.int64)
    cols_global = np.asarray(cols_global, dtype=np.int64)
    own = (cols_global >= lo) & (cols_global < hi)
    halo_ids = np.unique(cols_global[~own])
    col_local = np.where(own, cols_global - lo, (hi - lo) + np.searchsorted(halo_ids, cols_global))
    return rows_global - lo, col_local.astype(np.int64), halo_ids


class Partition:
    """One rank's slice: compact local indexing + the halo exchange plan."""

    def __init__(self, rank, world, bounds, rows_global, cols_global):
        """rows_global/cols_global: COO (row-major sorted) of the edges whose destination this rank
        owns.  Collective: all ranks must construct their Partition together (plan exchange)."""
        self.rank, self.world = rank, world
        self.bounds = np.asarray(bounds, dtype=np.int64)
        self.lo, self.hi = int(bounds[rank]), int(bounds[rank + 1])
        self.n_local = self.hi - self.lo
        self.n_total = int(bounds[-1])
        self.row_local, self.col_local, self.halo_ids = compact_columns(self.lo, self.hi, rows_global, cols_global)
        self.n_src = self.n_local + len(self.halo_ids)
        # ---- exchange plan: who owns my halo nodes, and which of my nodes others need
        owner = np.searchsorted(self.bounds, self.halo_ids, side="right") - 1
        self.recv_counts = np.bincount(owner, minlength=world).astype(np.int64)   # halo sorted => grouped
        want = [self.halo_ids[owner == r] for r in range(world)]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, want)
            asked = [gathered[src][rank] for src in range(world)]
        else:
            asked = [want[0]]
        self.send_counts = np.array([len(a) for a in asked], dtype=np.int64)
        self.send_idx = (np.concatenate(asked) - self.lo).astype(np.int64) if sum(self.send_counts) else \
            np.zeros(0, dtype=np.int64)
        self.graph = None          # set by attach_graph
        self._send_idx_dev = None

    def attach_graph(self, device, max_chunk=0):
        from .graph import Graph
        self.graph = Graph(self.n_local, self.row_local, self.col_local, device=device, max_chunk=max_chunk,
                           n_cols=self.n_src)
        return self

    def send_index(self, device):
        if self._send_idx_dev is None or self._send_idx_dev.device != device:
            self._send_idx_dev = torch.from_numpy(self.send_idx).to(device)
        return self._send_idx_dev


def _exchange(send, send_counts, recv_counts, width):
    """Variable all-to-all of row blocks with point-to-point ops (works on NCCL and gloo)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    recv = send.new_empty(int(recv_counts.sum()), width)
    s_off = np.concatenate([[0], np.cumsum(send_counts)])
    r_off = np.concatenate([[0], np.cumsum(recv_counts)])
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        if recv_counts[peer]:
            ops.append(dist.P2POp(dist.irecv, recv[r_off[peer]:r_off[peer + 1]], peer))
        if send_counts[peer]:
            ops.append(dist.P2POp(dist.isend, send[s_off[peer]:s_off[peer + 1]], peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv


class HaloExchange(torch.autograd.Function):
    """x_local[n_local, F] -> x_needed[n_local + n_halo, F] (own rows first, then halo rows in
    ascending global id).  Backward routes the halo rows' gradients back to their owners."""

    @staticmethod
    def forward(ctx, x_local, part):
        ctx.part = part
        if part.world == 1:
            return x_local
        send = x_local.index_select(0, part.send_index(x_local.device)).contiguous()
        recv = _exchange(send, part.send_counts, part.recv_counts, x_local.shape[1])
        return torch.cat([x_local, recv], 0)

    @staticmethod
    def backward(ctx, g):
        part = ctx.part
        if part.world == 1:
            return g, None
        g_local = g[:part.n_local].clone()
        back = _exchange(g[part.n_local:].contiguous(), part.recv_counts, part.send_counts, g.shape[1])
        g_local.index_add_(0, part.send_index(g.device), back)
        return g_local, None


def get_em_partitioned(enc, fusers, x_local, part, layer_fn=None):
    """`DISGAT.get_em` (models.py:217-252) over a destination-range partition: returns this
    rank's rows of [feature_1, feature_2]."""
    if layer_fn is None:
        from .layers import run_channels

        def layer_fn(chs, x_need, graph):
            return run_channels(chs, x_need, graph)[0]
    x = F.dropout(x_local, enc.dropout, training=enc.training)
    feats = []
    for layer, chs in enumerate((enc.attentions1, enc.attentions2)):
        x_need = HaloExchange.apply(x, part)
        out = layer_fn(chs, x_need, part.graph)
        fused = enc._fuse(layer, fusers, out, x)
        x = F.dropout(fused, enc.dropout, training=enc.training)
        feats.append(x)
    return feats


class AllGatherRows(torch.autograd.Function):
    """x_local[n_local, F] -> x_all[n_total, F] in global node order (the ranges are contiguous and
    rank-ordered).  Backward: every rank holds a partial gradient for ALL rows -> reduce-scatter to
    the owners.  Ranges may differ in size: rows are padded to the largest one."""

    @staticmethod
    def forward(ctx, x_local, part):
        ctx.part = part
        if part.world == 1:
            return x_local
        sizes = np.diff(part.bounds)
        m = int(sizes.max())
        pad = x_local.new_zeros(m, x_local.shape[1])
        pad[:part.n_local] = x_local
        bufs = [torch.empty_like(pad) for _ in range(part.world)]
        dist.all_gather(bufs, pad)
        return torch.cat([b[:int(s)] for b, s in zip(bufs, sizes)], 0)

    @staticmethod
    def backward(ctx, g_all):
        part = ctx.part
        if part.world == 1:
            return g_all, None
        if dist.get_backend() == "nccl":
            sizes = np.diff(part.bounds)
            m = int(sizes.max())
            chunks = []
            for r in range(part.world):
                c = g_all.new_zeros(m, g_all.shape[1])
                c[:int(sizes[r])] = g_all[int(part.bounds[r]):int(part.bounds[r + 1])]
                chunks.append(c)
            out = torch.empty_like(chunks[0])
            dist.reduce_scatter(out, chunks)
            return out[:part.n_local].contiguous(), None
        g = g_all.contiguous().clone()            # gloo has no reduce_scatter: all-reduce, keep own rows
        dist.all_reduce(g)
        return g[part.lo:part.hi].contiguous(), None


def ssl_pair_loss_partitioned(enc, fusers, x_local, part, pairs, labels, ranges, n_pos_total, m_total,
                              constrain_layer=0, layer_fn=None, pair_fn=None, loss_fn=None):
    """SupEdge / DisEdge loss (pretrainer.py:709-763, 578-641) over a destination-range partition.

    Pair sets are partitioned by the pair's ROW: pairs[k] = [2, M_k] int64 with LOCAL row ids
    (0..n_local) and GLOBAL column ids; labels[k][M_k].  The row operand P comes from this rank's
    nodes, the column operand Q from the all-gathered layer input (projections recomputed locally:
    F or D floats per node travel, not C*D).  n_pos_total[k] / m_total[k] are the positive count and
    size of the WHOLE set k (all ranks), so the per-rank return values SUM to the reference's loss
    and `allreduce_grads` yields its gradient.  ranges[k] = channel range consumed by set k."""
    if layer_fn is None or pair_fn is None or loss_fn is None:
        from . import functional as Fn
        from .layers import run_channels
        layer_fn = layer_fn or (lambda chs, x_need, graph: run_channels(chs, x_need, graph)[0])
        pair_fn = pair_fn or Fn.PairScore.apply
        loss_fn = loss_fn or Fn.SslWmse.apply
    from .functional import PairList
    from .layers import pair_operands
    pairs = [PairList.wrap(p) for p in pairs]
    x = F.dropout(x_local, enc.dropout, training=enc.training)
    loss = None
    for layer, chs in enumerate((enc.attentions1, enc.attentions2)):
        if constrain_layer == 0 or constrain_layer == layer:
            x_all = AllGatherRows.apply(x, part)
            att, C, D, P, Q, a = pair_operands(chs, x, x_all)
            for k, pr in enumerate(pairs):
                lo, hi = ranges[k]
                scores = pair_fn(att, C, D, pr.pi, pr.pj, lo, hi, P, Q, a, pr)
                term = loss_fn(scores, labels[k], n_pos_total[k], m_total[k])
                loss = term if loss is None else loss + term
        if layer == 0:       # layer-2 aggregation is dead compute for the pair losses (models.py:311-330)
            out = layer_fn(chs, HaloExchange.apply(x, part), part.graph)
            x = F.dropout(enc._fuse(0, fusers, out, x), enc.dropout, training=enc.training)
    if loss is None:
        raise ValueError("--constrain_layer=%d selects no layer" % constrain_layer)
    return loss


def sample_pairs_partitioned(part, pos_key_local, generator=None):
    """This rank's share of `sample_train` (pretrainer.py:683-707) with the O(M) device sampler:
    Bernoulli(3 rho) cells in the rows it owns (rho from the global edge count) united with a random
    third of its own positives.  pos_key_local: sorted int64 keys i_global * n_total + j of its
    positives.  Returns (pairs[2, M] with LOCAL rows / GLOBAL columns, label[M], n_pos_total, m_total)."""
    from .sampler import sample_pairs_device
    e_tot = torch.tensor([pos_key_local.numel()], dtype=torch.int64, device=pos_key_local.device)
    if part.world > 1:
        dist.all_reduce(e_tot)
    idx, lab = sample_pairs_device(part.n_total, pos_key_local, generator, row_range=(part.lo, part.hi),
                                   e_total=int(e_tot.item()))
    cnt = torch.stack([(lab != 0).sum(), torch.tensor(lab.numel(), device=lab.device)]).to(torch.int64)
    if part.world > 1:
        dist.all_reduce(cnt)
    pairs = torch.stack([idx[0] - part.lo, idx[1]])
    return pairs, lab, int(cnt[0].item()), int(cnt[1].item())


def allreduce_grads(params):
    """Sum the weight gradients over ranks (one flat bucket; the per-rank losses add up)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


# --------------------------------------------------------------------------- synthetic workload
def _community_endpoints(rng, n_loc, m, base, gamma=2.1, max_degree=None):
    """m power-law endpoints inside the community [base, base + n_loc) (see synthetic.py)."""
    expo = 1.0 / (gamma - 1.0)
    ranks = np.arange(n_loc, dtype=np.float64)
    if max_degree is None:
        max_degree = 8.0 * np.sqrt(n_loc) * max(1.0, m / (13.0 * n_loc))
    target = min(0.5, max_degree / (2.0 * max(m, 1)))
    lo, hi = 1.0, float(n_loc)
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
    relabel = np.random.RandomState(1000 + base % 100003).permutation(n_loc)
    return base + relabel[np.searchsorted(cdf, rng.random_sample(m))]


def build_partitioned_power_law(n_total, m_raw_total, seed, rank, world, device, locality=0.9, max_chunk=0):
    """Weak-scaling workload: `world` communities of n_total/world nodes (one per rank), power-law
    degrees inside each, a fraction (1 - locality) of the draws across communities.  Every rank
    generates only the blocks it touches; both ends of a cross block use the same seed, so the
    global graph is symmetric and consistent without any exchange of edges."""
    n_loc = n_total // world
    m_raw = m_raw_total // world
    bounds = np.arange(world + 1, dtype=np.int64) * n_loc
    m_in = int(m_raw * (locality if world > 1 else 1.0))
    m_pair = int((m_raw - m_in) / max(world - 1, 1))
    base = rank * n_loc
    rng = np.random.RandomState(seed * 7919 + rank)
    a = _community_endpoints(rng, n_loc, m_in, base)
    b = _community_endpoints(rng, n_loc, m_in, base)
    rows = [a, b, np.arange(base, base + n_loc)]
    cols = [b, a, np.arange(base, base + n_loc)]
    for peer in range(world):
        if peer == rank or m_pair == 0:
            continue
        lo_r, hi_r = min(rank, peer), max(rank, peer)
        prng = np.random.RandomState(seed * 104729 + lo_r * 131 + hi_r)
        u = _community_endpoints(prng, n_loc, m_pair, lo_r * n_loc)     # endpoint in the lower community
        v = _community_endpoints(prng, n_loc, m_pair, hi_r * n_loc)     # endpoint in the higher one
        mine, other = (u, v) if rank == lo_r else (v, u)
        rows.append(mine)
        cols.append(other)
    key = np.unique(np.concatenate(rows) * n_total + np.concatenate(cols))      # sort + dedup, row-major
    part = Partition(rank, world, bounds, key // n_total, key % n_total)
    return part.attach_graph(device, max_chunk)
