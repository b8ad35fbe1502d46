"""Reference-facing helpers (mirrors the hot-path part of /root/reference/utils.py).

  get_parser       utils.py:23-110   the CLI surface, flag for flag (defaults and choices kept)
  sp_softmax       utils.py:192-200  -> libedis COO kernels (global-max shift, +1e-10 kept)
  sp_matmul        utils.py:203-207
  adj_mse_loss     utils.py:287-298  -> fused weighted-MSE reduction (1-D sparse form)
  accuracy, split  utils.py:243-256, 112-161 (host-side evaluation helpers used by the trainers)
"""
import argparse
import random

import numpy as np
import torch

from . import functional as Fn

_FLAGS = [
    # name, kwargs  -- same names / defaults / choices as utils.py:25-109
    ("--no-cuda", dict(action="store_true", default=False, help="Disables CUDA training.")),
    ("--sparse", dict(action="store_true", default=False, help="whether use sparse adj matrix")),
    ("--seed", dict(type=int, default=4)),
    ("--nhid", dict(type=int, default=64)),
    ("--nclass", dict(type=int, default=5)),
    ("--dataset", dict(type=str, default="dblp")),
    ("--size", dict(type=int, default=64)),
    ("--epochs", dict(type=int, default=510, help="Number of epochs to train.")),
    ("--lr", dict(type=float, default=0.01)),
    ("--weight_decay", dict(type=float, default=5e-4)),
    ("--dropout", dict(type=float, default=0.1)),
    ("--batch_nums", dict(type=int, default=6000, help="number of batches per epoch")),
    ("--load", dict(type=int, default=None)),
    ("--save", dict(type=str, default=None)),
    ("--log", dict(action="store_true", default=False, help="whether save logs and checkpoints")),
    ("--method", dict(type=str, default="no", choices=["no"])),
    ("--model", dict(type=str, default="DISGAT",
                     choices=["sage", "gcn", "GAT", "sage2", "MLP", "RGCN", "HAN", "DISGAT", "GIN", "FactorGCN",
                              "Mixhop", "H2GCN"])),
    ("--nhead", dict(type=int, default=4)),
    ("--hetero", dict(action="store_true", default=False, help="whether using multiple edge types.")),
    ("--hnn", dict(action="store_true", default=False, help="whether use heterogeneous GNN.")),
    ("--edge_num", dict(type=int, default=3, help="number of edge types")),
    ("--used_edge", dict(type=int, default=1, help="0: using all egde. 1, 2, 3...: use only that edge type.")),
    ("--cls_layer", dict(type=int, default=2, help="number of layers in classifier. Must be larger than 0")),
    ("--EdgePred_layer", dict(type=int, default=1)),
    ("--downstream", dict(nargs="+", type=str, choices=["CLS", "Edge"])),
    ("--down_weight", dict(nargs="+", type=float)),
    ("--pretrain", dict(nargs="+", type=str,
                        choices=["PredAttr", "PredDistance", "PredContext", "DisEdge", "SupEdge", "DifHead"])),
    ("--pre_weight", dict(nargs="+", type=float)),
    ("--pre_edge", dict(nargs="+", type=int)),
    ("--finetune", dict(action="store_true", default=False, help="whether to train towards target task")),
    ("--enc_layer", dict(type=int, default=2, help="number of layers in the encoder")),
    ("--fuse", dict(type=str, default="last", choices=["last", "avg", "concat"])),
    ("--pretext_dim", dict(type=int, default=16)),
    ("--cluster_num", dict(type=int, default=16)),
    ("--node_sup_ratio", dict(type=float, default=0.25, help="ratio of nodes labeled")),
    ("--reg", dict(action="store_true", default=False, help="whether to regularize weight in fusers")),
    ("--reg_weight", dict(type=float, default=0.01, help="weight of l1 norm on fusers")),
    ("--batch", dict(action="store_true", default=False, help="whether use batches of sub-graphs as data")),
    ("--batch_size", dict(type=int, default=40)),
    ("--SubgraphSize", dict(type=int, default=128)),
    ("--origin_feat", dict(action="store_true", default=False, help="whether to use original feature")),
    ("--att", dict(type=int, default=2, help="Type of attention: 1 prototype product, 2 inner product, 3 MLP")),
    ("--dis_type", dict(type=int, default=1, help="1 for homo/hetero, 2 for class-homo")),
    ("--constrain_layer", dict(type=int, default=0,
                               help="0 for all layers; otherwise compared with the 0-based layer index")),
    ("--residue", dict(action="store_true", default=False, help="whether use residue for DISGAT model")),
    ("--fuse_no_relu", dict(action="store_true", default=False, help="whether use relu in fuser layer")),
    ("--residue_type", dict(type=int, default=0)),
    ("--steps", dict(type=int, default=5)),
    ("--gnn_type", dict(type=str, default="AT", choices=["AT", "SAGE", "GCN"], help="type of GNN in DISGAT")),
    ("--case", dict(action="store_true", default=False, help="whether case study mode")),
    ("--conformT", dict(action="store_true", default=False, help="case study on label conformity loss")),
]


def get_parser():
    parser = argparse.ArgumentParser()
    for name, kw in _FLAGS:
        parser.add_argument(name, **kw)
    return parser


def sp_softmax(indices, values, N):
    """Row softmax of COO values; drop-in for utils.py:192-200 (values [E, 1] or [E])."""
    return Fn.SpSoftmax.apply(indices[0], values, N)


def sp_matmul(indices, values, mat):
    """out[i] = sum_(i,j) values_ij * mat[j]; drop-in for utils.py:203-207."""
    return Fn.SpMatmul.apply(indices[0], indices[1], values, mat)


def adj_mse_loss(adj_rec, adj_tgt):
    """Class-balanced MSE on 1-D (sampled pairs) predictions; drop-in for utils.py:287-298.

    `adj_rec` are probabilities (post-sigmoid), as in the reference's call sites."""
    n_pos = int((adj_tgt != 0).sum())
    logits = torch.logit(adj_rec.reshape(-1, 1))
    return Fn.SslWmse.apply(logits, adj_tgt.reshape(-1), n_pos)


def group_correlation(embedding):
    """Pearson correlation between the ROWS of `embedding` [G, n] -> [G, G] (utils.py:326-334), with
    the norms taken from the centred rows directly instead of the diagonal of a G x G product."""
    x = embedding - embedding.mean(dim=-1, keepdim=True)
    nrm = torch.sqrt((x * x).sum(-1))
    return (x @ x.t()) / torch.outer(nrm, nrm)


def accuracy(output, labels):
    preds = output.max(1)[1].type_as(labels)
    return preds.eq(labels).double().sum() / len(labels)


def split(labels, train_ratio=0.25):
    """Class-stratified split driven by python `random` exactly like utils.py:112-161."""
    val_ratio = (1 - train_ratio) / 4
    test_ratio = (1 - train_ratio) / 4 * 3
    num_classes = len(set(labels.tolist()))
    train_idx, val_idx, test_idx = [], [], []
    c_num_mat = np.zeros((num_classes, 3)).astype(int)
    for i in range(num_classes):
        c_idx = (labels == i).nonzero()[:, -1].tolist()
        c_num = len(c_idx)
        print("{:d}-th class sample number: {:d}".format(i, c_num))
        random.shuffle(c_idx)
        if c_num < 11:
            raise ValueError("too small class type: {}, num{}".format(i, c_num))
        c_num_mat[i] = [int(c_num * train_ratio), int(c_num * val_ratio), int(c_num * test_ratio)]
        a, b, c = c_num_mat[i]
        train_idx += c_idx[:a]
        val_idx += c_idx[a:a + b]
        test_idx += c_idx[a + b:a + b + c]
    random.shuffle(train_idx)
    return torch.LongTensor(train_idx), torch.LongTensor(val_idx), torch.LongTensor(test_idx), c_num_mat
