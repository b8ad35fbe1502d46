"""Drop-in DISGAT model on the B200 path (mirrors /root/reference/models.py:151-373, 523-543).

Same constructor, submodule names (`attention{1,2}_{c}`, `fuser{1,2}`), parameter names and
the five traversal methods with the reference's list-of-lists return structure.  Each layer's
C channels run as ONE fused kernel launch; per-channel results are views of the fused tensors.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .graph import as_graph
from .layers import DisGALayer, FuseLayer, run_channels


class MLP(nn.Module):
    """Classifier head, same structure / state_dict keys as models.py:523-543."""

    def __init__(self, in_feat, hidden_size, out_size, layers=2, dropout=0.1):
        super().__init__()
        modules = []
        in_size = in_feat
        for _ in range(layers - 1):
            modules.append(nn.Linear(in_size, hidden_size))
            in_size = hidden_size
            modules.append(nn.LeakyReLU(0.1))
        modules.append(nn.Linear(in_size, out_size))
        self.model = nn.Sequential(*modules)

    def forward(self, features, cls=False):
        output = self.model(features)
        return F.log_softmax(output, dim=1) if cls else output


class DISGAT(nn.Module):
    """2 layers x C disentangled channels + external fusers (models.py:151-373)."""

    def __init__(self, args, nfeat, nhid, nclass, dropout, is_specific=[True, True], alpha=0.1, nheads=4):
        super().__init__()
        self.dropout = dropout
        self.args = args
        self.nheads = nheads
        self.gnn_type = args.gnn_type
        self.is_specific = is_specific
        self.attentions1 = [DisGALayer(nfeat, nhid, dropout=dropout, alpha=alpha, concat=True,
                                       att_type=args.att, gnn_type=self.gnn_type) for _ in range(nheads)]
        for i, attention in enumerate(self.attentions1):
            self.add_module("attention1_{}".format(i), attention)
        self.attentions2 = [DisGALayer(nhid, nclass, dropout=dropout, alpha=alpha, concat=True,
                                       att_type=args.att, gnn_type=self.gnn_type) for _ in range(nheads)]
        for i, attention in enumerate(self.attentions2):
            self.add_module("attention2_{}".format(i), attention)
        if args.residue:
            self.fuser1 = FuseLayer(args, nheads, nfeat=nhid, residue=nfeat)
            self.fuser2 = FuseLayer(args, nheads, nfeat=nhid, residue=nhid)
        else:
            self.fuser1 = FuseLayer(args, nheads, nfeat=nhid)
            self.fuser2 = FuseLayer(args, nheads, nfeat=nhid)

    # ------------------------------------------------------------------ shared traversal
    def _fuse(self, layer, fusers, out, x_in):
        own = self.fuser1 if layer == 0 else self.fuser2
        fuser = own if not self.is_specific[layer] else fusers[layer]
        if isinstance(fuser, FuseLayer):
            return fuser(out, x_in)               # channel-fused tensor: no cat copy
        d = out.shape[1] // self.nheads           # a reference FuseLayer wants the list
        return fuser([out[:, c * d:(c + 1) * d] for c in range(self.nheads)], x_in)

    def traverse(self, x, adj, fusers, aux=None, aux_ranges=None, need_layer2_agg=True):
        """The traversal all five reference methods share (models.py:181-373).

        Returns a dict: x_in (post-dropout layer inputs), out (fused [N, C*D] per layer, elu'd),
        edge_e ([E, C] per layer), aux (list over pair sets of [M, C'] per layer), fused
        (fuser outputs before dropout).  With need_layer2_agg=False the layer-2 aggregation and
        fuser (dead compute in predict_adjs_sparse, models.py:311-330) are skipped.
        """
        if not isinstance(fusers, list):
            fusers = [fusers]
        graph = as_graph(adj)
        if aux is not None:
            from .functional import PairList
            aux = [PairList.wrap(p) for p in aux]     # per-set derived data shared by both layers
        res = dict(x_in=[], out=[], edge_e=[], aux=[], fused=[])
        x = F.dropout(x, self.dropout, training=self.training)
        for layer, chs in enumerate((self.attentions1, self.attentions2)):
            res["x_in"].append(x)
            agg = need_layer2_agg or layer == 0
            out, edge_e, auxs = run_channels(chs, x, graph, aux, aggregate=agg, aux_ranges=aux_ranges)
            res["out"].append(out)
            res["edge_e"].append(edge_e)
            res["aux"].append(auxs)
            if not agg:
                break
            fused = self._fuse(layer, fusers, out, x)
            res["fused"].append(fused)
            x = F.dropout(fused, self.dropout, training=self.training)
        res["x_last"] = x
        return res

    def _channels(self, t):
        d = t.shape[1] // self.nheads
        return [t[:, c * d:(c + 1) * d] for c in range(self.nheads)]

    # ------------------------------------------------------------------ reference API
    def forward(self, x, adj, fusers):
        r = self.traverse(x, adj, fusers)
        return F.log_softmax(r["fused"][1], dim=1)

    def get_em(self, x, adj, fusers):
        r = self.traverse(x, adj, fusers)
        return [r["x_in"][1], r["x_last"]]

    def get_adjs(self, x, adj, fusers):
        """[[edge_e_c[E, 1]]_c]_layer, raw (pre-sigmoid) logits like models.py:254-288."""
        r = self.traverse(x, adj, fusers)
        return [[e[:, c:c + 1] for c in range(self.nheads)] for e in r["edge_e"]]

    def predict_adjs_sparse(self, x, adj, fusers, auxiliary_edges):
        """[[ [aux_c^k[M_k, 1]]_k ]_c]_layer like models.py:290-330."""
        if not isinstance(auxiliary_edges, (list, tuple)):
            auxiliary_edges = [auxiliary_edges]
        r = self.traverse(x, adj, fusers, aux=auxiliary_edges, need_layer2_agg=False)
        return [[[s[:, c:c + 1] for s in auxs] for c in range(self.nheads)] for auxs in r["aux"]]

    def get_edge_em(self, x, adj, fusers):
        """[[cat(x_in, elu(h'_c))]_c]_layer like models.py:333-373."""
        r = self.traverse(x, adj, fusers)
        return [[torch.cat((r["x_in"][l], oc), dim=-1) for oc in self._channels(r["out"][l])] for l in range(2)]
