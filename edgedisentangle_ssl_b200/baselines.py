"""Sparse GAT and FactorGCN layers on the same softmax + SpMM kernels (second half).

Drop-ins for the SPARSE branch of the reference's two other attention layers, as thin modules over
`edis_sp_softmax_*` / `edis_sp_matmul_*` (utils.sp_softmax / utils.sp_matmul drop-ins):
  GraphAttentionLayer   layers.py:229-296   e_ij = lrelu_alpha([hW_i || hW_j] a), alpha = sp_softmax(e)
                                            (global-max shift, +1e-10), h' = sum_j alpha_ij (hW)_j
  DisentangleLayer      layers.py:515-573   FactorGCN: per latent factor l, e = att_l([h_i || h_j]),
                                            alpha = sp_softmax(sigmoid(e)), h'_l = sum_j alpha_ij emb(x)_j
Same constructor signatures, parameter names / shapes / init order (a reference state_dict loads
unchanged).  The [E, 2D] gather + cat of the reference collapses to two per-node scalars
([h_i || h_j] a = h_i . a_top + h_j . a_bot), so no [E, .] feature temporary exists; everything per edge
that remains is the [E, 1] logit.  The dense (non --sparse) branches are out of scope like DisGALayer's.
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib, utils
from .graph import as_graph


def _indices(adj):
    g = as_graph(adj)
    return g.indices, g.n


class GraphAttentionLayer(nn.Module):
    """Sparse GAT layer (layers.py:229-296)."""

    def __init__(self, in_features, out_features, dropout, alpha, concat=True):
        super().__init__()
        self.dropout = dropout
        self.in_features = in_features
        self.out_features = out_features
        self.alpha = alpha
        self.concat = concat
        self.W = nn.Parameter(torch.zeros(size=(in_features, out_features)))
        nn.init.xavier_uniform_(self.W.data, gain=1.414)
        self.a = nn.Parameter(torch.zeros(size=(2 * out_features, 1)))
        nn.init.xavier_uniform_(self.a.data, gain=1.414)
        self.leakyrelu = nn.LeakyReLU(self.alpha)

    def forward(self, input, adj):
        if not input.is_cuda:
            raise _lib.EdisError("GraphAttentionLayer input must be a CUDA tensor; there is no CPU fallback")
        indices, n = _indices(adj)
        h = torch.mm(input, self.W)
        d = self.out_features
        s_dst, s_src = h @ self.a[:d], h @ self.a[d:]                       # [N, 1] each
        edge_e = self.leakyrelu(s_dst[indices[0]] + s_src[indices[1]])      # layers.py:255-256
        attention = utils.sp_softmax(indices, edge_e, n)
        attention = F.dropout(attention, self.dropout, training=self.training)
        h_prime = utils.sp_matmul(indices, attention, h)
        return F.elu(h_prime) if self.concat else h_prime

    def __repr__(self):
        return self.__class__.__name__ + " (" + str(self.in_features) + " -> " + str(self.out_features) + ")"


class DisentangleLayer(nn.Module):
    """Sparse FactorGCN layer (layers.py:515-573)."""

    def __init__(self, in_features, out_features, concat=True, n_latent=4):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.concat = concat
        self.n_latent = n_latent
        self.linear = nn.Linear(in_features, self.out_features)
        self.att_ls = nn.ModuleList()
        for _ in range(self.n_latent):
            self.att_ls.append(nn.Linear(self.out_features * 2, 1))
        self.emb = nn.Linear(in_features, int(self.out_features // n_latent))
        assert int(self.out_features // n_latent) * n_latent == out_features, \
            "Inconsistency in FactorGNN heads structure"

    def forward(self, input, adj):
        if not input.is_cuda:
            raise _lib.EdisError("DisentangleLayer input must be a CUDA tensor; there is no CPU fallback")
        indices, n = _indices(adj)
        h = self.linear(input)
        h_em = self.emb(input)
        d = self.out_features
        heads = []
        for att in self.att_ls:
            w = att.weight                                                    # [1, 2D]
            s_dst = h @ w[:, :d].t() + att.bias
            s_src = h @ w[:, d:].t()
            edge_e = s_dst[indices[0]] + s_src[indices[1]]                    # layers.py:547-549
            attention = utils.sp_softmax(indices, torch.sigmoid(edge_e), n)
            heads.append(utils.sp_matmul(indices, attention, h_em))
        return torch.cat(heads, dim=-1)          # no activation, whatever `concat` says (layers.py:587-597)
