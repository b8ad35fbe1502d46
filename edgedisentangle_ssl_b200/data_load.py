"""Graph / feature loading on the CSR builder (mirrors /root/reference/data_load.py:22-94).

`load_data(args, path, dataset, edge_type)` keeps the reference's signature and return value
(processed adjacency as torch sparse COO float32 -- or a list of them when `args.hetero` --,
float32 features, int64 labels) but never builds an N x N matrix: the adjacency goes through
`edis_build_adjacency_host` (bit-exact, see tests/test_host_abi.py).
"""
import os

import numpy as np
import scipy.sparse as sp
import torch

from .graph import build_adjacency


def normalize(mx):
    """Row-normalise features: data_load.py:137-144 (float64, inf -> 0)."""
    mx = np.asarray(mx, dtype=np.float64)
    rowsum = mx.sum(1)
    with np.errstate(divide="ignore"):
        r_inv = np.power(rowsum, -1.0)
    r_inv[np.isinf(r_inv)] = 0.0
    return r_inv[:, None] * mx


def load_graph_arrays(path, index=1):
    """(n, rows, cols, vals) of `adj_{index}.npy` (edge list) or `adj_{index}_sp.npz` (CSR).

    Edge lists size the graph by max id + 1 like utils.edge2adj (utils.py:163-170)."""
    p_npy = os.path.join(path, "adj_{}.npy".format(index))
    if os.path.exists(p_npy):
        edge = np.load(p_npy)
        if edge.shape[1] == 2 and edge.shape[0] != 2:
            edge = edge.astype(np.int64)
            return int(edge.max()) + 1, edge[:, 0], edge[:, 1], None
        m = sp.coo_matrix(edge)
    else:
        m = sp.load_npz(os.path.join(path, "adj_{}_sp.npz".format(index))).tocoo()
    return m.shape[0], m.row.astype(np.int64), m.col.astype(np.int64), m.data.astype(np.float64)


def to_sparse_tensor(n, indices, values):
    """data_load.py:158-165: COO in CSR order, float32 values."""
    return torch.sparse_coo_tensor(torch.from_numpy(indices), torch.from_numpy(values), (n, n))


def synthetic_features(labels, dim=64, seed=0):
    """Deterministic class-informative features for graphs whose feature blobs are missing from
    the reference snapshot (cora, cora_full): SURVEY 8(d)."""
    rng = np.random.RandomState(seed)
    mu = rng.randn(int(labels.max()) + 1, dim)
    return np.abs(mu[labels] + rng.randn(labels.shape[0], dim))


def load_data(args, path="data/dblp/", dataset="dblp", edge_type=3):
    print("Loading {} dataset...".format(dataset))
    labels = np.load(os.path.join(path, "label.npy"))
    if getattr(args, "origin_feat", False):
        features = np.load(os.path.join(path, "feature.npy"))
    else:
        fpath = os.path.join(path, "feature_new.npy")
        features = np.load(fpath) if os.path.exists(fpath) else synthetic_features(labels)
        features = normalize(features)
    graphs = [load_graph_arrays(path, i + 1) for i in range(edge_type)]
    if args.hetero:
        use = graphs
    elif args.used_edge == 0:
        # union of all edge types, clipped to {0, 1} (data_load.py:56-60)
        n = graphs[0][0]
        rows = np.concatenate([g[1] for g in graphs])
        cols = np.concatenate([g[2] for g in graphs])
        use = [(n, rows, cols, None)]
    else:
        use = [graphs[args.used_edge - 1]]
    processed = []
    for n, rows, cols, vals in use:
        idx, val = build_adjacency(n, rows, cols, vals)
        if not args.sparse:
            raise NotImplementedError("the B200 path implements --sparse only (layers.py:340-416)")
        processed.append(to_sparse_tensor(n, idx, val))
    features = torch.FloatTensor(np.array(features))
    labels = torch.LongTensor(labels)
    print("Data loaded")
    return (processed, features, labels) if args.hetero else (processed[0], features, labels)
