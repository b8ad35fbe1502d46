"""Graph / feature loading on the CSR builder (mirrors /root/reference/data_load.py:22-94).

`load_data(args, path, dataset, edge_type)` keeps the reference's signature and return value
(processed adjacency as torch sparse COO float32 -- or a list of them when `args.hetero` --,
float32 features, int64 labels) but never builds an N x N matrix: the adjacency goes through
`edis_build_adjacency_host` (bit-exact, see tests/test_host_abi.py).
"""
import hashlib
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


def processed_adjacency(path, index=1, cache_dir=None):
    """(n, indices[2, E], values[E]) of edge type `index`: `load_graph_arrays` + `build_adjacency`,
    through an .npz cache when `cache_dir` is given.  The cache is keyed by a hash of the
    RAW input file's bytes, so a changed or replaced input can never be served from a stale file."""
    if not cache_dir:
        n, rows, cols, vals = load_graph_arrays(path, index)
        return (n,) + build_adjacency(n, rows, cols, vals)
    src = os.path.join(path, "adj_{}.npy".format(index))
    if not os.path.exists(src):
        src = os.path.join(path, "adj_{}_sp.npz".format(index))
    key = hashlib.blake2b(open(src, "rb").read() + b"|build_adjacency v1", digest_size=8).hexdigest()
    cpath = os.path.join(cache_dir, "adj_{}_{}.npz".format(index, key))
    if os.path.exists(cpath):
        try:
            z = np.load(cpath)
            return int(z["n"]), z["indices"], z["values"]
        except Exception:            # truncated / foreign file: rebuild
            pass
    n, rows, cols, vals = load_graph_arrays(path, index)
    idx, val = build_adjacency(n, rows, cols, vals)
    os.makedirs(cache_dir, exist_ok=True)
    tmp = cpath + ".tmp%d.npz" % os.getpid()
    np.savez(tmp, n=np.int64(n), indices=idx, values=val)
    os.replace(tmp, cpath)
    return n, idx, val


def load_data(args, path="data/dblp/", dataset="dblp", edge_type=3):
    print("Loading {} dataset...".format(dataset))
    labels = np.load(os.path.join(path, "label.npy"))
    if getattr(args, "origin_feat", False):
        features = np.load(os.path.join(path, "feature.npy"))
    else:
        fpath = os.path.join(path, "feature_new.npy")
        if os.path.exists(fpath):
            features = np.load(fpath)
        elif os.environ.get("EDIS_SYNTH_FEATURES") == "1":
            # the reference fails on a missing feature file (data_load.py:36).  cora / cora_full ship without
            # theirs (.MISSING_LARGE_BLOBS), so tests and benchmarks may opt in to class-informative synthetic
            # features -- which are built FROM THE LABELS: accuracy on them says nothing about real data
            print("WARNING: {} not found; EDIS_SYNTH_FEATURES=1 -> label-derived synthetic features "
                  "(benchmark / test use only; any accuracy on them is label-leaked)".format(fpath))
            features = synthetic_features(labels)
        else:
            raise FileNotFoundError(
                "{} not found (the reference needs it too, data_load.py:36).  For benchmarks / tests on the bundled "
                "cora / cora_full graphs, whose feature blobs are missing from the reference snapshot, set "
                "EDIS_SYNTH_FEATURES=1 to use label-derived synthetic features".format(fpath))
        features = normalize(features)
    cache_dir = os.environ.get("EDIS_CACHE_DIR") or None
    if args.hetero or args.used_edge != 0:
        want = range(edge_type) if args.hetero else [args.used_edge - 1]
        processed_coo = {i: processed_adjacency(path, i + 1, cache_dir) for i in want}
        graphs = None
    else:
        graphs = [load_graph_arrays(path, i + 1) for i in range(edge_type)]
    if not args.sparse:
        raise NotImplementedError("the B200 path implements --sparse only (layers.py:340-416)")
    if graphs is None:
        coo = [processed_coo[i] for i in sorted(processed_coo)]
    else:
        # union of all edge types, clipped to {0, 1} (data_load.py:56-60)
        n = graphs[0][0]
        rows = np.concatenate([g[1] for g in graphs])
        cols = np.concatenate([g[2] for g in graphs])
        coo = [(n,) + build_adjacency(n, rows, cols, None)]
    processed = [to_sparse_tensor(n, idx, val) for n, idx, val in coo]
    features = torch.FloatTensor(np.array(features))
    labels = torch.LongTensor(labels)
    print("Data loaded")
    return (processed, features, labels) if args.hetero else (processed[0], features, labels)
