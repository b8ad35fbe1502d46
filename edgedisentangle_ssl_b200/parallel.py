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
