"""Drop-in DISGAT layers on the B200 path (mirrors /root/reference/layers.py for this path).

Same class names, constructor signatures, parameter names / shapes / init order as the
reference, so a reference `state_dict` loads unchanged and same-seed initialisation is equal:
  DisGALayer      layers.py:303-511   (sparse branch only; the dense branch is out of scope)
  FuseLayer       layers.py:876-921
  SageConv        layers.py:63-112    (parameters + projection; aggregation is fused upstream)
  GraphConvolution layers.py:16-59    (parameters; aggregation is fused upstream)
All channels of a DISGAT layer run in ONE fused kernel launch (`run_channels`).
"""
import math
import os

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.parameter import Parameter

from . import _lib
from .functional import (ChannelLinear, DisGAFused, PairList, PairScore, Proj3xTF32, SageFused, node_linear,
                         use_proj3x)
from .graph import as_graph

_seed_state = {"count": 0}


def _next_seed():
    # counter-based dropout stream derived from torch's seed: no device->host sync
    _seed_state["count"] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seed_state["count"] * 0xD1B54A32D192ED03) & (2 ** 64 - 1)


def dropout_stream_state():
    """Position of the attention-dropout stream (saved with a checkpoint so that a resumed run does not
    replay the masks of its first epochs)."""
    return int(_seed_state["count"])


def set_dropout_stream_state(count):
    _seed_state["count"] = int(count)


class GraphConvolution(nn.Module):
    """Parameters of the reference GCN layer (layers.py:16-59): weight[F, D], bias[D]."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = Parameter(torch.FloatTensor(in_features, out_features))
        if bias:
            self.bias = Parameter(torch.FloatTensor(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)


class SageConv(nn.Module):
    """Parameters of the reference GraphSage layer (layers.py:63-112): proj = Linear(2F, D)."""

    def __init__(self, in_features, out_features, bias=False):
        super().__init__()
        self.proj = nn.Linear(in_features * 2, out_features, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.proj.weight)
        if self.proj.bias is not None:
            nn.init.constant_(self.proj.bias, 0.0)


def run_channels(chs, x, graph, aux=None, aggregate=True, aux_ranges=None):
    """Run C DisGALayer channels that share an input as one fused launch.

    chs: list of DisGALayer with identical (in, out, att, gnn).  Returns
      out    [N, C*D] = cat_c elu(h'_c)  (None if aggregate=False)
      edge_e [E, C]   raw logits           (None if aggregate=False)
      auxs   list over pair sets of [M_k, c_hi-c_lo] logits (None if aux is None)
    The dense projections are ordinary fp32 torch GEMMs (one per layer, all channels and all
    operands concatenated); everything per-edge happens in libedis.so.
    """
    l0 = chs[0]
    C, D, Fin, att, gnn = len(chs), l0.out_features, l0.in_features, l0.att_type, l0.gnn_type
    if x.dtype != torch.float32 or not x.is_cuda:
        raise _lib.EdisError("DisGALayer input must be a float32 CUDA tensor; there is no CPU fallback")
    CD = C * D
    # Execution plan for gnn_type AT / GCN.  "proj" (default): project first, gather V_j = (x W)_j per
    # edge with 128-bit loads (C*D floats per edge).  "agg" (EDIS_AT_PLAN=agg, F <= 256): aggregate the
    # raw input per channel with the shared-operand kernels and project afterwards,
    # (sum_j a_ij x_j) W == sum_j a_ij (x_j W): only F floats of the aggregated operand are gathered per
    # edge, in the forward and in the destination pass of the backward.  Where the shared operand fits
    # the 128-bit layout (F == D == 64, C in {2, 4, 8}: DISGAT's second layer) its kernels beat the
    # "proj" ones (B200, config A: fwd 38.7 vs 44 ms, dst pass 27.4 vs 34.6 ms), but they are issue-bound,
    # not HBM-bound, and the extra per-channel GEMM + ELU passes eat the gain (291 vs 286 ms per step,
    # profiles/r1_sweep10.log); for other F the operand runs lane-strided and is slower still
    # (profiles/r1_sweep4_plans.log).  SAGE always uses the shared-operand kernels.
    plan = os.environ.get("EDIS_AT_PLAN") or "proj"
    use_agg = aggregate and (gnn == "SAGE" or plan == "agg")
    a = None
    if att == 3:
        a = torch.cat([l.a.reshape(1, D) for l in chs], 0)
        w_score = [torch.cat([l.W[:Fin] for l in chs], 1), torch.cat([l.W[Fin:] for l in chs], 1)]
    else:
        w_score = [torch.cat([l.W for l in chs], 1)]
    bias = None
    w_val = []
    if aggregate and gnn == "GCN" and l0.ag_layer.bias is not None:
        bias = torch.cat([l.ag_layer.bias for l in chs], 0)
    if aggregate and not use_agg:
        w_val = [torch.cat([(l.W_em if gnn == "AT" else l.ag_layer.weight) for l in chs], 1)]
    w_all = torch.cat(w_score + w_val, 1)
    if use_proj3x(x.shape[0]):
        proj = Proj3xTF32.apply(x, w_all)             # fp32-accurate 3xTF32 split on the tensor cores
    else:
        proj = x @ w_all                              # one fp32 GEMM (TF32 off) for all channels / operands
    sdst = ssrc = None
    if att == 3:
        off_p, off_q, off_v = 0, CD, 2 * CD
        P, Q = proj[:, :CD], proj[:, CD:2 * CD]
    else:
        off_p, off_q, off_v = 0, 0, CD
        P = Q = proj[:, :CD]
        if att == 1:
            h3 = P.reshape(-1, C, D)
            a_top = torch.cat([l.a[:D].reshape(1, D) for l in chs], 0)
            a_bot = torch.cat([l.a[D:].reshape(1, D) for l in chs], 0)
            P = sdst = (h3 * a_top).sum(-1)           # [N, C] destination-side scalar
            Q = ssrc = (h3 * a_bot).sum(-1)           # [N, C] source-side scalar
    auxs = None
    if aux is not None:
        auxs = []
        for k, pairs in enumerate(aux):
            lo, hi = (0, C) if aux_ranges is None else aux_ranges[k]
            pl = PairList.wrap(pairs)
            auxs.append(PairScore.apply(att, C, D, pl.pi, pl.pj, lo, hi, P, Q, a, pl))
    if not aggregate:
        return None, None, auxs
    training, p = l0.training, l0.dropout
    seed = _next_seed() if (training and p > 0) else 0
    if gnn == "SAGE":
        neigh, edge_e = SageFused.apply(graph, att, C, D, proj, off_p, off_q, sdst, ssrc, a, x, training, p,
                                        seed, False)                                       # [N, C*F]
        wp = torch.stack([l.ag_layer.proj.weight for l in chs], 0)                         # [C, D, 2F]
        self_part = torch.einsum("nf,cdf->ncd", x[:graph.n], wp[:, :, :Fin])
        neigh_part = torch.einsum("ncf,cdf->ncd", neigh.reshape(-1, C, Fin), wp[:, :, Fin:])
        h = self_part + neigh_part
        if l0.ag_layer.proj.bias is not None:
            h = h + torch.stack([l.ag_layer.proj.bias for l in chs], 0)
        out = F.elu(h).reshape(-1, CD)
    elif use_agg:
        agg, edge_e = SageFused.apply(graph, att, C, D, proj, off_p, off_q, sdst, ssrc, a, x, training, p,
                                      seed, True)                                          # [N, C*F]
        w_em = torch.stack([(l.W_em if gnn == "AT" else l.ag_layer.weight) for l in chs], 0)  # [C, F, D]
        h = ChannelLinear.apply(agg, w_em)
        out = F.elu(h + bias if bias is not None else h)
    else:
        with _maybe_recompute(proj, x, w_all, lean=(aux is None and att != 1)):
            out, edge_e = DisGAFused.apply(graph, att, C, D, proj, off_p, off_q, off_v, sdst, ssrc, a, bias,
                                           training, p, seed)
    return out, edge_e, auxs


class _Recompute:
    """Marker packed in place of the saved projection (see _maybe_recompute)."""


def _maybe_recompute(proj, x, w_all, lean):
    """Memory-lean mode: `proj[N, 3*C*D]` is 6 KB per source row and layer and is the largest tensor
    the fused layer saves for its backward.  With EDIS_RECOMPUTE_PROJ=1 -- or automatically when one
    projection exceeds an eighth of the device memory -- the autograd graph keeps a marker instead
    and the backward recomputes it from (x, W) with the same GEMM (bit-identical; ~3 ms per layer
    at config A), so it is resident for one layer at a time instead of for every layer."""
    import contextlib
    mode = os.environ.get("EDIS_RECOMPUTE_PROJ", "auto")
    if not lean or mode == "0" or not proj.is_cuda or not proj.requires_grad:
        return contextlib.nullcontext()
    if mode != "1":
        total = torch.cuda.get_device_properties(proj.device).total_memory
        if proj.numel() * 4 * 8 < total:
            return contextlib.nullcontext()
    key, shape = proj.data_ptr(), proj.shape
    xd, wd = x.detach(), w_all.detach()
    use3x = use_proj3x(x.shape[0])

    def pack(t):
        return _Recompute if (t.data_ptr() == key and t.shape == shape) else t

    def unpack(h):
        if h is _Recompute:
            from .functional import _mm_3xtf32
            return _mm_3xtf32(xd, wd) if use3x else xd @ wd
        return h

    return torch.autograd.graph.saved_tensors_hooks(pack, unpack)


def pair_operands(chs, x_dst, x_src):
    """Score-side operands of C channels for pairs whose row side and column side live in two
    different node tensors (multi-GPU SSL: rows = this rank's nodes, columns = all nodes):
    P from x_dst, Q from x_src, as `run_channels` builds them from one tensor (layers.py:349-389).
    Pure torch.  Returns (att, C, D, P, Q, a)."""
    l0 = chs[0]
    C, D, Fin, att = len(chs), l0.out_features, l0.in_features, l0.att_type
    if att == 3:
        a = torch.cat([l.a.reshape(1, D) for l in chs], 0)
        P = x_dst @ torch.cat([l.W[:Fin] for l in chs], 1)
        Q = x_src @ torch.cat([l.W[Fin:] for l in chs], 1)
        return att, C, D, P, Q, a
    w = torch.cat([l.W for l in chs], 1)
    P, Q = x_dst @ w, x_src @ w
    if att == 1:
        a_top = torch.cat([l.a[:D].reshape(1, D) for l in chs], 0)
        a_bot = torch.cat([l.a[D:].reshape(1, D) for l in chs], 0)
        P = (P.reshape(-1, C, D) * a_top).sum(-1)
        Q = (Q.reshape(-1, C, D) * a_bot).sum(-1)
    return att, C, D, P, Q, None


class DisGALayer(nn.Module):
    """One disentangled attention channel; drop-in for layers.py:303-511 (sparse branch)."""

    def __init__(self, in_features, out_features, dropout, alpha, concat=True, att_type=1, gnn_type="AT"):
        super().__init__()
        self.dropout = dropout
        self.in_features = in_features
        self.out_features = out_features
        self.alpha = alpha          # stored but unused, like the reference (leaky slope is 0.01)
        self.concat = concat
        self.att_type = att_type
        self.gnn_type = gnn_type
        if att_type == 3:
            self.W = nn.Parameter(torch.zeros(size=(in_features * 2, out_features)))
            nn.init.xavier_uniform_(self.W.data, gain=1.414)
            self.a = nn.Parameter(torch.zeros(size=(out_features, 1)))
            nn.init.xavier_uniform_(self.a.data, gain=1.414)
        else:
            self.W = nn.Parameter(torch.zeros(size=(in_features, out_features)))
            nn.init.xavier_uniform_(self.W.data, gain=1.414)
            self.a = nn.Parameter(torch.zeros(size=(2 * out_features, 1)))
            nn.init.xavier_uniform_(self.a.data, gain=1.414)
        if gnn_type == "AT":
            self.W_em = nn.Parameter(torch.zeros(size=(in_features, out_features)))
            nn.init.xavier_uniform_(self.W_em.data, gain=1.414)
        elif gnn_type == "SAGE":
            self.ag_layer = SageConv(in_features, out_features)
        elif gnn_type == "GCN":
            self.ag_layer = GraphConvolution(in_features, out_features)
        else:
            raise ValueError("not implemented for gnn_type {} in DISGAT".format(gnn_type))

    def forward(self, input, adj, aux_indices=None):
        """-> (elu(h')[N, D], edge_e[E, 1]) or (..., [aux_k[M_k, 1]]) like layers.py:493-511."""
        graph = as_graph(adj)
        if aux_indices is not None and not isinstance(aux_indices, (list, tuple)):
            aux_indices = [aux_indices]
        out, edge_e, auxs = run_channels([self], input, graph, aux_indices)
        if not self.concat:
            raise _lib.EdisError("concat=False is not used by DISGAT (models.py:163,168) and not built")
        if aux_indices is not None:
            return out, edge_e, auxs
        return out, edge_e


class FuseLayer(nn.Module):
    """Fuses the C channel outputs; drop-in for layers.py:876-921.

    `forward` also accepts the channel-fused [N, C*D] tensor directly, which makes the
    reference's `torch.cat(feature_list, -1)` (layers.py:900) free.
    """

    def __init__(self, args, nheads, nfeat=64, residue=0):
        super().__init__()
        self.args = args
        self.nheads = nheads
        self.nfeat = nfeat
        self.residue_dim = residue
        if self.args.residue_type == 0:
            self.fuse = nn.Linear(self.nfeat * nheads + self.residue_dim, self.nfeat)
        if self.args.residue_type == 1:
            self.fuse = nn.Linear(self.nfeat * nheads + self.residue_dim, self.nfeat * 2)
            self.fuse2 = nn.Linear(self.nfeat * 2, self.nfeat)
        if self.args.residue_type == 2:
            self.fuse = nn.Linear(self.nfeat * nheads, self.nfeat)
            if self.residue_dim != 0:
                self.fuse2 = nn.Linear(self.residue_dim, self.nfeat)

    def forward(self, feature_list, residue=None):
        features = feature_list if torch.is_tensor(feature_list) else torch.cat(feature_list, dim=-1)
        use_res = self.residue_dim != 0 and residue is not None
        rt = self.args.residue_type
        if rt == 0:
            if use_res:
                features = torch.cat([features, residue], dim=-1)
            feature = node_linear(self.fuse, features)
        elif rt == 1:
            if use_res:
                features = torch.cat([features, residue], dim=-1)
            feature = node_linear(self.fuse2, F.leaky_relu(node_linear(self.fuse, features)))
        elif rt == 2:
            feature = node_linear(self.fuse, features)
            if use_res:
                feature = feature + node_linear(self.fuse2, residue)
        return feature if self.args.fuse_no_relu else F.leaky_relu(feature)
