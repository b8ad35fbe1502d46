"""Oracle (CPU, numpy) for graph construction and SSL pair sampling.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Each function restates one
piece of the reference and cites it; nothing here densifies N x N.
"""
import numpy as np
import torch


def features_row_normalize(x):
    """Row-sum normalisation of node features.

    Follows data_load.py:137-144 (`normalize`): r = rowsum**-1 with inf -> 0, x <- diag(r) x.
    Done in float64 like the reference (features are loaded as float64 .npy).
    """
    x = np.asarray(x, dtype=np.float64)
    rs = x.sum(1)
    with np.errstate(divide="ignore"):
        r = np.power(rs, -1.0)
    r[np.isinf(r)] = 0.0
    return r[:, None] * x


def build_adjacency(n, rows, cols, vals=None):
    """Processed adjacency of `load_data` as a row-major COO, without the dense detour.

    Follows data_load.py:69-77 + 12-20 + 158-165:
      * `np.fill_diagonal(adj, 1)`            -> every (i, i) present with value 1
      * `adj + adj.T*(adj.T>adj) - adj*(adj.T>adj)` -> elementwise max(A, A^T)
      * `normalize_adj`                        -> D^-1 A with float64 row sums
      * `sp.csr_matrix` -> `.tocoo()` -> float32 values, int64 indices in CSR order
    For an edge-list input the reference first runs `utils.edge2adj` (utils.py:163-170):
    duplicates collapse to 1 and the matrix side is `edgelist.max()+1`.
    Returns (indices[2, E] int64 row-major sorted, values[E] float32).
    """
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    if vals is None:
        vals = np.ones(rows.shape[0], dtype=np.float64)
    vals = np.asarray(vals, dtype=np.float64)
    # off-diagonal entries in both orientations; the diagonal is overwritten with 1
    off = rows != cols
    r = np.concatenate([rows[off], cols[off], np.arange(n, dtype=np.int64)])
    c = np.concatenate([cols[off], rows[off], np.arange(n, dtype=np.int64)])
    v = np.concatenate([vals[off], vals[off], np.ones(n)])
    # explicit zeros are not edges (csr_matrix(dense) drops them)
    keep = v != 0
    r, c, v = r[keep], c[keep], v[keep]
    key = r * n + c
    order = np.argsort(key, kind="stable")
    key, v = key[order], v[order]
    uniq, start = np.unique(key, return_index=True)
    vmax = np.maximum.reduceat(v, start)
    ur, uc = uniq // n, uniq % n
    deg = np.zeros(n, dtype=np.float64)
    np.add.at(deg, ur, vmax)
    with np.errstate(divide="ignore"):
        rinv = np.power(deg, -1.0)
    rinv[np.isinf(rinv)] = 0.0
    values = (rinv[ur] * vmax).astype(np.float32)
    return np.vstack([ur, uc]).astype(np.int64), values


def edge_list_to_coo(edges):
    """`utils.edge2adj` (utils.py:163-170) without the dense matrix: n = max id + 1."""
    edges = np.asarray(edges).astype(np.int64)
    n = int(edges.max()) + 1
    return n, edges[:, 0], edges[:, 1]


def homo_hetero_split(indices, labels):
    """DisEdge label sets (pretrainer.py:440-456) as edge masks over the processed adjacency.

    homo = edges (incl. self loops) whose endpoints share a label, hetero = the rest.
    Returns two [2, E_k] int64 arrays in row-major order.
    """
    labels = np.asarray(labels)
    same = labels[indices[0]] == labels[indices[1]]
    return indices[:, same], indices[:, ~same]


def sample_pairs(n, pos_indices, chunk_rows=256):
    """One call of `sample_train` for one label set, bit-exact and streaming.

    Follows pretrainer.py:683-707 (SupEdge) / 552-574 (DisEdge, per label set):
      thr  = (label.sum() / (N*N)).item() * 3      (float32 tensor division, python *3)
      mask = torch.rand(N, N) < thr                 (CPU default generator)
      idx  = label.nonzero()  (row-major) ; np.random.shuffle(idx) ; first E//3 forced in
      out  = mask.nonzero().T (row-major, deduped), labels = label[out]
    `torch.rand((rows, N))` drawn in consecutive row chunks consumes the CPU generator exactly
    like one `torch.rand((N, N))` (one 24-bit draw per element, serial fill; checked for chunk
    heights 1, 5, 7, 16 and in tests/test_oracle_golden.py), so memory is O(chunk * N).
    `pos_indices` is [2, E_L] int64 row-major sorted.  Returns (indices[2, M], label[M] f32).
    """
    e_l = pos_indices.shape[1]
    ratio = (torch.tensor(float(e_l), dtype=torch.float32) / (n * n)).item()
    thr = ratio * 3
    chunk_rows = max(1, int(chunk_rows))
    keys = []
    for r0 in range(0, n, chunk_rows):
        r1 = min(n, r0 + chunk_rows)
        u = torch.rand(size=(r1 - r0, n))
        nz = (u < thr).nonzero().numpy()
        keys.append((nz[:, 0].astype(np.int64) + r0) * n + nz[:, 1])
    pos = np.ascontiguousarray(pos_indices.T)  # [E_L, 2], row-major like adj.nonzero()
    np.random.shuffle(pos)
    forced = pos[: e_l // 3]
    keys.append(forced[:, 0] * n + forced[:, 1])
    key = np.unique(np.concatenate(keys))
    pos_key = np.sort(pos_indices[0] * n + pos_indices[1])
    lab = np.isin(key, pos_key, assume_unique=True).astype(np.float32)
    return np.vstack([key // n, key % n]).astype(np.int64), lab
