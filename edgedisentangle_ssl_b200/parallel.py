"""Multi-GPU execution of the DISGAT path: destination-range graph partitioning, per-layer halo
exchange of source features and weight-gradient all-reduce (the reference itself is
single-process, so there is no reference behaviour to mirror here).

One process per GPU (`torch.distributed`, NCCL over NVLink; gloo on CPU for the tests).
  * rank r owns the contiguous destination rows [lo_r, hi_r) and their in-edges (CSR slice);
  * the sources its edges read are its own nodes plus a HALO of remote nodes; columns are
    re-indexed compactly (own first, halo after) and the slice becomes a rectangular edis graph;
  * per layer the layer INPUT rows (F or D floats per node, not the 2*C*D projected ones) of the
    halo nodes reach the rank by ONE collective (`SourceExchange`): a padded all-gather when the halo
    is most of the graph (a power-law graph cut into ranges: 71 % of all nodes are sources of every
    rank at 8 ranks), a variable all-to-all of only the needed rows when it is not (locality-ordered
    graphs); the reverse collective (reduce-scatter / all-to-all) returns input gradients to owners;
  * projections are recomputed locally: the destination operand P for the rank's OWN rows only, the
    source operands Q | V for own + halo rows (`PartitionedLayer`, one autograd node per layer with a
    hand-ordered backward so that the exchange runs under the own-row GEMMs);
  * weight gradients are all-reduced in one rank-invariant flat bucket.
"""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F


def row_ranges(rowptr, world, balance="edges"):
    """world+1 boundaries of contiguous destination ranges, balanced by in-edge (or node) count."""
    n = len(rowptr) - 1
    if balance == "nodes":
        return np.linspace(0, n, world + 1).round().astype(np.int64)
    if n < world:
        raise ValueError("cannot cut %d rows into %d non-empty ranges" % (n, world))
    target = rowptr[-1] * np.arange(1, world) / world
    cuts = np.clip(np.searchsorted(rowptr, target, side="left"), 0, n)
    b = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    # a hub row heavier than E/world would leave a neighbour range empty: every rank keeps >= 1 row
    for r in range(1, world):
        b[r] = max(b[r], b[r - 1] + 1)
    for r in range(world - 1, 0, -1):
        b[r] = min(b[r], b[r + 1] - 1)
    return b


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
    """One rank's slice: compact local indexing + the exchange plans."""

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
        # ---- all-gather plan: ranges are padded to the largest one; source row s of this rank sits at
        # row pad_index[s] of the gathered [world * max_rows, F] buffer
        self.max_rows = int(np.diff(self.bounds).max())
        owner = np.searchsorted(self.bounds, self.halo_ids, side="right") - 1
        self.pad_index = np.concatenate([rank * self.max_rows + np.arange(self.n_local, dtype=np.int64),
                                         owner * self.max_rows + (self.halo_ids - self.bounds[owner])])
        # ---- all-to-all plan: who owns my halo nodes, and which of my nodes others need
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
        # exchange mode: all-gather moves n_total rows per rank, the all-to-all only the halo (but needs a
        # gather of the rows to send and an index_add on the way back): all-gather once the halo is
        # more than half of the other ranks' nodes.  All ranks must agree -> decided on the global sum.
        mode = os.environ.get("EDIS_EXCHANGE", "auto")
        if mode == "auto":
            tot = np.array([len(self.halo_ids), self.n_total - self.n_local], dtype=np.int64)
            if world > 1:
                g = [None] * world
                dist.all_gather_object(g, tot.tolist())
                tot = np.sum(np.array(g, dtype=np.int64), axis=0)
            mode = "allgather" if 2 * tot[0] > tot[1] else "alltoall"
        self.mode = mode
        self.graph = None          # set by attach_graph
        self._dev = {}

    def attach_graph(self, device, max_chunk=0):
        from .graph import Graph
        self.graph = Graph(self.n_local, self.row_local, self.col_local, device=device, max_chunk=max_chunk,
                           n_cols=self.n_src)
        return self

    def _on(self, name, arr, device):
        key = (name, str(device))
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
        return self._dev[key]

    def send_index(self, device):
        return self._on("send", self.send_idx, device)

    def pad_index_dev(self, device):
        return self._on("pad", self.pad_index, device)

    def halo_pad_index_dev(self, device):
        return self._on("halo_pad", self.pad_index[self.n_local:], device)


def partition_of_global_graph(idx, n, rank, world, device=None, max_chunk=0, balance="edges"):
    """This rank's Partition of ONE global graph (strong scaling): `idx` [2, E] is the processed,
    row-major sorted adjacency every rank holds (or generates from the same seed); destination rows are
    cut into `world` contiguous ranges balanced by in-edge count.  Collective."""
    deg = np.bincount(idx[0], minlength=n)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    bounds = row_ranges(rowptr, world, balance)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    e0, e1 = int(rowptr[lo]), int(rowptr[hi])
    part = Partition(rank, world, bounds, idx[0][e0:e1], idx[1][e0:e1])
    return part.attach_graph(device, max_chunk) if device is not None else part


# ------------------------------------------------------------------------------- collectives
class _Pending:
    """A collective in flight (async_op=True: it runs on the backend's own stream; `wait` makes the
    CURRENT stream wait for it, the host does not block) plus what turns its buffer into the result."""

    def __init__(self, work, finish):
        self.work, self.finish = work, finish

    def wait(self):
        if self.work is not None:
            self.work.wait()
        return self.finish()


def start_source_gather(x_local, part):
    """Start fetching the source rows this rank's edges read: -> _Pending whose wait() returns
    x_src[n_src, F] = own rows followed by the halo rows (ascending global id)."""
    if part.world == 1:
        return _Pending(None, lambda: x_local)
    F_ = x_local.shape[1]
    dev = x_local.device
    if part.mode == "allgather":
        if part.n_local == part.max_rows:
            pad = x_local.contiguous()
        else:
            pad = x_local.new_zeros(part.max_rows, F_)
            pad[:part.n_local] = x_local
        buf = x_local.new_empty(part.world * part.max_rows, F_)
        work = dist.all_gather_into_tensor(buf, pad, async_op=True)

        def finish():
            halo = buf.index_select(0, part.halo_pad_index_dev(dev))
            return torch.cat([x_local, halo], 0)
        return _Pending(work, finish)
    send = x_local.index_select(0, part.send_index(dev))
    recv = x_local.new_empty(int(part.recv_counts.sum()), F_)
    work = dist.all_to_all_single(recv, send, part.recv_counts.tolist(), part.send_counts.tolist(), async_op=True)
    return _Pending(work, lambda: torch.cat([x_local, recv], 0))


def start_source_scatter(g_src, part):
    """Reverse of start_source_gather for gradients: g_src[n_src, F] (partial gradient of every source
    row this rank read) -> _Pending whose wait() returns the part of g_local[n_local, F] that comes
    from OTHER ranks' reads plus this rank's own rows of g_src (sum over ranks of the rows it owns)."""
    if part.world == 1:
        return _Pending(None, lambda: g_src)
    F_ = g_src.shape[1]
    dev = g_src.device
    if part.mode == "allgather":
        buf = g_src.new_zeros(part.world * part.max_rows, F_)
        buf.index_copy_(0, part.pad_index_dev(dev), g_src)
        out = g_src.new_empty(part.max_rows, F_)
        work = dist.reduce_scatter_tensor(out, buf, async_op=True)
        return _Pending(work, lambda: out[:part.n_local])
    send = g_src[part.n_local:].contiguous()
    back = g_src.new_empty(int(part.send_counts.sum()), F_)
    work = dist.all_to_all_single(back, send, part.send_counts.tolist(), part.recv_counts.tolist(), async_op=True)

    def finish():
        g_local = g_src[:part.n_local].clone()
        g_local.index_add_(0, part.send_index(dev), back)
        return g_local
    return _Pending(work, finish)


class SourceExchange(torch.autograd.Function):
    """x_local[n_local, F] -> x_src[n_local + n_halo, F] (own rows first, then halo rows in
    ascending global id).  Backward routes the source rows' gradients back to their owners."""

    @staticmethod
    def forward(ctx, x_local, part):
        from .functional import phase
        ctx.part = part
        with phase("exchange_exposed"):
            return start_source_gather(x_local, part).wait()

    @staticmethod
    def backward(ctx, g):
        from .functional import phase
        with phase("exchange_exposed"):
            return start_source_scatter(g.contiguous(), ctx.part).wait(), None


HaloExchange = SourceExchange      # round-1 name


# ------------------------------------------------------------------------------- one layer
def _edis_kernels():
    """(forward, backward) of the fused layer on raw operands: libedis through functional.py."""
    from . import functional as Fn

    def fwd(graph, d, P, QV, a, bias, want_sign):
        CD = d.C * d.D
        return Fn.disga_forward_raw(graph, d, Fn._ptr(P), P.stride(0), Fn._ptr(QV), QV.stride(0),
                                    Fn._off(QV, CD), QV.stride(0), a, bias, P.device, want_sign)

    def bwd(graph, d, P, QV, a, bias, saved, g_out, g_edge_e):
        CD = d.C * d.D
        hpre, edge_e, stats, esign = saved
        gP = torch.empty_like(P)
        gQV = torch.empty_like(QV)
        ga = torch.zeros(d.C, d.D, dtype=torch.float32, device=P.device) if d.att == 3 else None
        gh = Fn.disga_backward_raw(graph, d, Fn._ptr(P), P.stride(0), Fn._ptr(QV), QV.stride(0), Fn._off(QV, CD),
                                   QV.stride(0), a, bias, hpre, edge_e, stats, esign, g_out, g_edge_e,
                                   Fn._ptr(gP), gP.stride(0), Fn._ptr(gQV), gQV.stride(0), ga, Fn._off(gQV, CD),
                                   gQV.stride(0), P.device)
        return gP, gQV, ga, (gh.sum(0) if bias is not None else None)

    return fwd, bwd


def _project(x, w):
    """Node projection at fp32 accuracy: 3xTF32 on the tensor cores for big node counts (see
    functional.Proj3xTF32), plain fp32 otherwise (and always on CPU)."""
    from .functional import _mm_3xtf32, use_proj3x
    if x.is_cuda and use_proj3x(x.shape[0]):
        return _mm_3xtf32(x, w)
    return x @ w


class PartitionedLayer(torch.autograd.Function):
    """One DISGAT layer (gnn_type AT / GCN, att 3) on a destination-range partition, as ONE autograd node:

      forward   start the source exchange of x_local -> P = x_local W_top (own rows only, runs under the
                exchange) -> x_src = own + halo rows -> Q|V = x_src [W_bot | W_val] -> fused kernel
      backward  fused backward kernels -> gP[n_local], gQ|gV[n_src] -> (layer 2 only) input gradient of
                the source rows, sent back to the owners by the reverse collective, which runs under the
                weight-gradient GEMMs and the own-row part -> sum.

    Layer 1 never needs an input gradient (x = features), which removes the largest of the per-rank
    GEMMs over own + halo rows.  Saved for the backward: x_src (F floats per source row), P, Q|V, and the
    kernel's own records.  `kernels` = (fwd, bwd) callables on raw operands (libedis by default; the CPU
    gloo tests plug in the oracle)."""

    @staticmethod
    def forward(ctx, x_local, w_top, w_qv, a, bias, part, desc, kernels):
        from .functional import phase
        pend = start_source_gather(x_local, part)
        with phase("gemm_fwd"):
            P = _project(x_local, w_top)                   # overlaps with the exchange
        with phase("exchange_exposed"):
            x_src = pend.wait()
        with phase("gemm_fwd"):
            QV = _project(x_src, w_qv)
        kfwd, _ = kernels
        a_c = a.contiguous()
        bias_c = bias.contiguous() if bias is not None else None
        out, hpre, edge_e, stats, esign = kfwd(part.graph, desc, P, QV, a_c, bias_c, True)
        ctx.part, ctx.desc, ctx.kernels = part, desc, kernels
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x_local, x_src, w_top, w_qv, a_c, bias_c, P, QV, hpre, edge_e, stats, esign)
        ctx.set_materialize_grads(False)
        return out, edge_e

    @staticmethod
    def backward(ctx, g_out, g_edge_e):
        from .functional import _xt_g
        x_local, x_src, w_top, w_qv, a, bias, P, QV, hpre, edge_e, stats, esign = ctx.saved_tensors
        part, d = ctx.part, ctx.desc
        if g_out is None:
            g_out = torch.zeros(part.n_local, d.C * d.D, dtype=torch.float32, device=P.device)
        g_out = g_out.contiguous()
        if g_edge_e is not None:
            g_edge_e = g_edge_e.contiguous()
        gP, gQV, ga, gbias = ctx.kernels[1](part.graph, d, P, QV, a, bias, (hpre, edge_e, stats, esign), g_out,
                                            g_edge_e)
        del P, QV, hpre, edge_e, stats, esign
        from .functional import phase
        need_gx = ctx.needs_input_grad[0]
        pend = None
        with phase("gemm_bwd"):
            if need_gx:
                pend = start_source_scatter(gQV @ w_qv.t(), part)     # [n_src, F] -> owners (async)
            g_wqv = _xt_g(x_src, gQV) if ctx.needs_input_grad[2] else None      # these run under the collective
            g_wtop = _xt_g(x_local, gP) if ctx.needs_input_grad[1] else None
            gx_own = gP @ w_top.t() if need_gx else None
        gx = None
        if need_gx:
            with phase("exchange_exposed"):
                gx = pend.wait()
            gx = gx + gx_own
        return gx, g_wtop, g_wqv, (ga if ctx.needs_input_grad[3] else None), gbias, None, None, None


def _edis_agg_kernels():
    """(forward, backward) of the shared-operand (aggregate-then-project) layer on raw operands."""
    from . import functional as Fn

    def fwd(graph, d, P, Q, X, a, want_sign):
        return Fn.sage_forward_raw(graph, d, Fn._ptr(P), P.stride(0), Fn._ptr(Q), Q.stride(0), X, a, want_sign)

    def bwd(graph, d, P, Q, X, a, saved, g_agg, g_edge_e, need_gx):
        agg, edge_e, stats, esign = saved
        gP = torch.empty_like(P)
        gQ = torch.empty_like(Q)
        ga = torch.zeros(d.C, d.D, dtype=torch.float32, device=P.device)
        gX = Fn.sage_backward_raw(graph, d, Fn._ptr(P), P.stride(0), Fn._ptr(Q), Q.stride(0), X, a, agg, edge_e, stats,
                                  esign, g_agg, g_edge_e, Fn._ptr(gP), gP.stride(0), Fn._ptr(gQ), gQ.stride(0), ga,
                                  need_gx)
        return gP, gQ, ga, gX

    return fwd, bwd


class PartitionedAggLayer(torch.autograd.Function):
    """Layer 2 of DISGAT (F == D == 64) on a partition as AGGREGATE-THEN-PROJECT (att 3, gnn_type AT / GCN):
    (sum_j a_ij x_j) W_em == sum_j a_ij (x_j W_em), so only the SCORE operand Q is projected for the own + halo
    source rows (C*D columns instead of 2*C*D) and the value projection W_em runs on the rank's OWN rows
    (functional.ChannelLinear, outside this node).  The projection GEMMs over source rows are what does not
    shrink with the rank count on a graph without locality; this halves them for the layer.

      forward   exchange x_local  ||  P = x_local W_top  ->  Q = x_src W_bot  ->  shared-operand kernel
                -> agg[n_local, C*F] = sum_j alpha_ij x_j per channel
      backward  kernels -> gP[n_local], gQ[n_src], gX_src[n_src, F] (aggregation part, from the source pass)
                -> dX_src = gX_src + gQ W_bot^T  -> reverse exchange  ||  dW GEMMs + own-row part."""

    @staticmethod
    def forward(ctx, x_local, w_top, w_bot, a, part, desc, kernels):
        from .functional import phase
        pend = start_source_gather(x_local, part)
        with phase("gemm_fwd"):
            P = _project(x_local, w_top)
        with phase("exchange_exposed"):
            x_src = pend.wait().contiguous()
        with phase("gemm_fwd"):
            Q = _project(x_src, w_bot)
        a_c = a.contiguous()
        agg, edge_e, stats, esign = kernels[0](part.graph, desc, P, Q, x_src, a_c, True)
        ctx.part, ctx.desc, ctx.kernels = part, desc, kernels
        ctx.save_for_backward(x_local, x_src, w_top, w_bot, a_c, P, Q, agg, edge_e, stats, esign)
        ctx.set_materialize_grads(False)
        return agg, edge_e

    @staticmethod
    def backward(ctx, g_agg, g_edge_e):
        from .functional import _xt_g, phase
        x_local, x_src, w_top, w_bot, a, P, Q, agg, edge_e, stats, esign = ctx.saved_tensors
        part, d = ctx.part, ctx.desc
        if g_agg is None:
            g_agg = torch.zeros_like(agg)
        g_agg = g_agg.contiguous()
        if g_edge_e is not None:
            g_edge_e = g_edge_e.contiguous()
        need_gx = ctx.needs_input_grad[0]
        gP, gQ, ga, gX_src = ctx.kernels[1](part.graph, d, P, Q, x_src, a, (agg, edge_e, stats, esign), g_agg, g_edge_e,
                                            need_gx)
        del P, Q, agg, edge_e, stats, esign
        pend = None
        with phase("gemm_bwd"):
            if need_gx:
                pend = start_source_scatter(torch.addmm(gX_src, gQ, w_bot.t()), part)
            g_wbot = _xt_g(x_src, gQ) if ctx.needs_input_grad[2] else None
            g_wtop = _xt_g(x_local, gP) if ctx.needs_input_grad[1] else None
            gx_own = gP @ w_top.t() if need_gx else None
        gx = None
        if need_gx:
            with phase("exchange_exposed"):
                gx = pend.wait()
            gx = gx + gx_own
        return gx, g_wtop, g_wbot, (ga if ctx.needs_input_grad[3] else None), None, None, None


def layer_partitioned(chs, x_local, part, training=None, kernels=None, agg_kernels=None):
    """C DisGALayer channels on this rank's rows: -> out[n_local, C*D] (= cat_c elu(h'_c)), edge_e."""
    from . import _lib
    from .layers import _next_seed
    l0 = chs[0]
    C, D, Fin, att, gnn = len(chs), l0.out_features, l0.in_features, l0.att_type, l0.gnn_type
    if att != 3 or gnn not in ("AT", "GCN"):
        # generic route: exchange the inputs, then the single-GPU layer on own + halo rows
        from .layers import run_channels
        out, edge_e, _ = run_channels(chs, SourceExchange.apply(x_local, part), part.graph)
        return out, edge_e
    w_top = torch.cat([l.W[:Fin] for l in chs], 1)
    w_qv = torch.cat([l.W[Fin:] for l in chs] + [(l.W_em if gnn == "AT" else l.ag_layer.weight) for l in chs], 1)
    a = torch.cat([l.a.reshape(1, D) for l in chs], 0)
    bias = None
    if gnn == "GCN" and l0.ag_layer.bias is not None:
        bias = torch.cat([l.ag_layer.bias for l in chs], 0)
    training = l0.training if training is None else training
    p = l0.dropout
    # the rank is folded into the seed: ranks hash LOCAL edge ids and must not draw identical masks
    seed = (_next_seed() + 0x632BE59BD9B4E019 * (part.rank + 1)) & (2 ** 64 - 1) if (training and p > 0) else 0
    desc = _lib.LayerDesc(att=att, C=C, D=D, Dv=D, training=1 if (training and p > 0) else 0, p=float(p), seed=seed)
    # EDIS_PART_PLAN=agg: aggregate-then-project (PartitionedAggLayer) -- halves the source-row projection GEMMs of
    # a layer with F == D == 64.  Measured on config A (profiles/r2_bench_n2_layer2_agg_plan.json, N = 2): GEMMs
    # 42.4 -> 33.4 ms, but the shared-operand kernels have no ring variant (+7 ms) and the per-channel W_em GEMMs +
    # ELU run as separate passes over the own rows (+12 ms): 142.6 -> 153.0 ms.  At 8 ranks the own rows are 4x
    # fewer and the estimate is break-even, so project-then-aggregate stays the default.
    plan = os.environ.get("EDIS_PART_PLAN", "proj")
    if agg_kernels is not None or plan == "agg":
        from .functional import ChannelLinear
        desc.Dv = Fin
        desc.flags = _lib.FLAG_PLAIN_MEAN | (0 if x_local.requires_grad else _lib.FLAG_NO_GX)
        w_bot = torch.cat([l.W[Fin:] for l in chs], 1)
        agg, edge_e = PartitionedAggLayer.apply(x_local, w_top, w_bot, a, part, desc, agg_kernels or _edis_agg_kernels())
        w_em = torch.stack([(l.W_em if gnn == "AT" else l.ag_layer.weight) for l in chs], 0)      # [C, F, D]
        h = ChannelLinear.apply(agg, w_em)
        return F.elu(h + bias if bias is not None else h), edge_e
    return PartitionedLayer.apply(x_local, w_top, w_qv, a, bias, part, desc, kernels or _edis_kernels())


def get_em_partitioned(enc, fusers, x_local, part, layer_fn=None, kernels=None, agg_kernels=None):
    """`DISGAT.get_em` (models.py:217-252) over a destination-range partition: returns this
    rank's rows of [feature_1, feature_2].  layer_fn(chs, x_src, graph) -> out replaces the whole layer
    after a plain SourceExchange (round-1 test hook); kernels = (fwd, bwd) replaces only the fused
    kernels inside PartitionedLayer."""
    x = F.dropout(x_local, enc.dropout, training=enc.training)
    feats = []
    for layer, chs in enumerate((enc.attentions1, enc.attentions2)):
        if layer_fn is not None:
            out = layer_fn(chs, SourceExchange.apply(x, part), part.graph)
        else:
            # agg_kernels (test hook) selects the aggregate-then-project node for the layers it applies to
            ak = agg_kernels if (agg_kernels is not None and chs[0].in_features == chs[0].out_features) else None
            out = layer_partitioned(chs, x, part, kernels=kernels, agg_kernels=ak)[0]
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
        m = part.max_rows
        pad = x_local.new_zeros(m, x_local.shape[1])
        pad[:part.n_local] = x_local
        buf = x_local.new_empty(part.world * m, x_local.shape[1])
        dist.all_gather_into_tensor(buf, pad)
        sizes = np.diff(part.bounds)
        if int(sizes.min()) == m:
            return buf
        return torch.cat([buf[r * m:r * m + int(sizes[r])] for r in range(part.world)], 0)

    @staticmethod
    def backward(ctx, g_all):
        part = ctx.part
        if part.world == 1:
            return g_all, None
        m = part.max_rows
        sizes = np.diff(part.bounds)
        if int(sizes.min()) == m:
            buf = g_all.contiguous()
        else:
            buf = g_all.new_zeros(part.world * m, g_all.shape[1])
            for r in range(part.world):
                buf[r * m:r * m + int(sizes[r])] = g_all[int(part.bounds[r]):int(part.bounds[r + 1])]
        out = g_all.new_empty(m, g_all.shape[1])
        dist.reduce_scatter_tensor(out, buf)
        return out[:part.n_local].contiguous(), None


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


def allreduce_grads(params, async_op=False):
    """Sum the weight gradients over ranks in ONE flat bucket (the per-rank losses add up).  The bucket
    covers every parameter -- a missing gradient (a rank whose slice never touched the parameter) is sent
    as zeros -- so its size is the same on all ranks whatever each rank's autograd graph looked like.
    async_op=True returns a callable that waits and writes the sums back (overlap with later work)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return (lambda: None) if async_op else None
    params = [p for p in params if p.requires_grad]
    if not params:
        return (lambda: None) if async_op else None
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    seen = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], device=flat.device)
    flat = torch.cat([flat, seen])
    work = dist.all_reduce(flat, async_op=True)

    def finish():
        work.wait()
        off = 0
        got = flat[-len(params):].tolist() if any(p.grad is None for p in params) else None
        for i, p in enumerate(params):
            k = p.numel()
            if p.grad is not None:
                p.grad.copy_(flat[off:off + k].view_as(p))
            elif got is not None and got[i] > 0:          # some other rank produced a gradient for it
                p.grad = flat[off:off + k].view_as(p).clone()
            off += k
    if async_op:
        return finish
    finish()


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
