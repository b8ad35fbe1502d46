"""CSR/CSC graph builder and device graph handle (host side of include/edis.h).

Replaces the reference's dense / COO handling: `data_load.load_data` (data_load.py:39-77),
`utils.edge2adj` (utils.py:163-170) and the per-call `adj.coalesce().indices()` of
layers.py:344.
"""
import ctypes
import os
from ctypes import c_double, c_float, c_int32, c_int64, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, np_ptr


def build_adjacency(n, rows, cols, vals=None):
    """Processed adjacency of the reference's `load_data` without the N x N detour.

    Self loops (value 1), symmetrise by max, row-normalise; returns (indices[2,E] int64
    row-major sorted, values[E] float32) bit-identical to data_load.py:69-77 + 158-165.
    """
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    m = rows.shape[0]
    vp = None
    if vals is not None:
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        vp = np_ptr(vals, c_double)
    cap = 2 * m + n
    out_r = np.empty(cap, dtype=np.int64)
    out_c = np.empty(cap, dtype=np.int64)
    out_v = np.empty(cap, dtype=np.float32)
    e = lib.edis_build_adjacency_host(n, m, np_ptr(rows, c_int64), np_ptr(cols, c_int64), vp,
                                      np_ptr(out_r, c_int64), np_ptr(out_c, c_int64), np_ptr(out_v, c_float))
    check(e, "edis_build_adjacency_host")
    return np.stack([out_r[:e], out_c[:e]]), out_v[:e].copy()


class Graph:
    """Device-resident CSR + CSC + work schedules of one adjacency (edis_graph)."""

    def __init__(self, n, row, col, device=None, max_chunk=0, n_cols=None):
        """n destination rows; n_cols >= n source columns (rectangular = one rank's slice of a
        destination-range partition, own nodes first then halo; default square)."""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.EdisError("edgedisentangle_ssl_b200 runs on CUDA devices only (got %s); "
                                 "there is no CPU fallback" % device)
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        row = np.ascontiguousarray(row, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int64)
        self._e_in = row.shape[0]
        h = c_void_p()
        check(lib.edis_graph_create_rect(n, n if n_cols is None else int(n_cols), row.shape[0],
                                         np_ptr(row, c_int64), np_ptr(col, c_int64), int(max_chunk),
                                         self.device.index, ctypes.byref(h)), "edis_graph_create_rect")
        self._h = h
        self.max_chunk = int(max_chunk)
        self._read_info()

    def _read_info(self):
        info = np.zeros(10, dtype=np.int64)
        check(lib.edis_graph_info(self._h, np_ptr(info, c_int64)), "edis_graph_info")
        self.n, self.e, self.n_cols = int(info[0]), int(info[1]), int(info[9])
        self.info = dict(n=self.n, n_cols=self.n_cols, e=self.e, dst_items=int(info[2]), dst_slots=int(info[3]),
                         src_items=int(info[4]), src_slots=int(info[5]), max_in=int(info[6]),
                         max_out=int(info[7]), was_sorted=bool(info[8]))
        self._indices = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.edis_graph_destroy(h)
            self._h = None

    @property
    def handle(self):
        return self._h

    # ------------------------------------------------------------------ on-disk cache
    @staticmethod
    def content_key(n, row, col, max_chunk=0, n_cols=None):
        """64-bit content key of an input edge list: what a cache file of this graph is keyed by."""
        row = np.ascontiguousarray(row, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int64)
        return int(lib.edis_edge_list_key(n, n if n_cols is None else int(n_cols), row.shape[0], np_ptr(row, c_int64),
                                          np_ptr(col, c_int64), int(max_chunk)))

    def save(self, path, key):
        """Write rowptr / col / perm / CSC / schedules to `path` (atomic), keyed by `key` (any 64-bit
        integer naming the INPUT: `Graph.content_key(...)` of the edge list, or a hash of generator
        parameters) and by this graph's max_chunk."""
        os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
        check(lib.edis_graph_save(self._h, os.fsencode(path), int(key) & (2 ** 64 - 1)), "edis_graph_save")
        return path

    @classmethod
    def load(cls, path, key, device=None, max_chunk=0, verify=True):
        """Graph from a cache file written by `save` for the same key and max_chunk, memory-mapped and
        uploaded to `device`; None when there is no usable file (missing, truncated, corrupt, other
        format version, other key = stale): the caller rebuilds and saves."""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.EdisError("edgedisentangle_ssl_b200 runs on CUDA devices only (got %s); "
                                 "there is no CPU fallback" % device)
        dev = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        h = c_void_p()
        rc = lib.edis_graph_load(os.fsencode(path), int(key) & (2 ** 64 - 1), int(max_chunk), dev.index,
                                 1 if verify else 0, ctypes.byref(h))
        if rc == _lib.ERR_STALE:
            return None
        check(rc, "edis_graph_load")
        g = cls.__new__(cls)
        g.device, g._h, g.max_chunk, g._indices = dev, h, int(max_chunk), None
        g._read_info()
        g._e_in = int(lib.edis_graph_input_entries(h))
        return g

    @classmethod
    def cached(cls, path, n, row, col, device=None, max_chunk=0, n_cols=None):
        """`Graph(n, row, col, ...)` through a cache file at `path`, keyed by the content of (row, col)."""
        key = cls.content_key(n, row, col, max_chunk, n_cols)
        g = cls.load(path, key, device, max_chunk) if path else None
        if g is None:
            g = cls(n, row, col, device=device, max_chunk=max_chunk, n_cols=n_cols)
            if path:
                g.save(path, key)
        return g

    def workspace_bytes(self, width):
        return int(check(lib.edis_graph_workspace_bytes(self._h, int(width)), "edis_graph_workspace_bytes"))

    def export(self):
        """Host copies: rowptr, col, perm (input entry -> CSR slot), cscptr, cscrow, csceid."""
        rowptr = np.empty(self.n + 1, dtype=np.int64)
        col = np.empty(max(self.e, 1), dtype=np.int32)
        perm = np.empty(max(self._e_in, 1), dtype=np.int64)
        cscptr = np.empty(self.n_cols + 1, dtype=np.int64)
        cscrow = np.empty(max(self.e, 1), dtype=np.int32)
        csceid = np.empty(max(self.e, 1), dtype=np.int32)
        check(lib.edis_graph_export(self._h, np_ptr(rowptr, c_int64), np_ptr(col, c_int32),
                                    np_ptr(perm, c_int64), np_ptr(cscptr, c_int64),
                                    np_ptr(cscrow, c_int32), np_ptr(csceid, c_int32)), "edis_graph_export")
        return dict(rowptr=rowptr, col=col[: self.e], perm=perm[: self._e_in],
                    cscptr=cscptr, cscrow=cscrow[: self.e], csceid=csceid[: self.e])

    @property
    def indices(self):
        """[2, E] int64 device tensor == the reference's `adj.coalesce().indices()`."""
        if self._indices is None:
            ex = self.export()
            deg = np.diff(ex["rowptr"])
            row = np.repeat(np.arange(self.n, dtype=np.int64), deg)
            self._indices = torch.from_numpy(np.stack([row, ex["col"].astype(np.int64)])).to(self.device)
        return self._indices

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_sparse(cls, adj, max_chunk=0, cache_dir=None):
        """Graph of a torch sparse COO adjacency; cached on the tensor object and, when `cache_dir`
        (default: $EDIS_CACHE_DIR) is set, on disk keyed by the content of its index list."""
        g = getattr(adj, "_edis_graph", None)
        if g is not None:
            return g
        if isinstance(adj, Graph):
            return adj
        if not adj.is_sparse:
            raise _lib.EdisError("the B200 path implements the reference's sparse branch only "
                                 "(layers.py:340-416); pass a torch sparse COO adjacency (--sparse)")
        if not adj.is_cuda:
            raise _lib.EdisError("adjacency must live on a CUDA device; there is no CPU fallback")
        idx = adj.coalesce().indices().cpu().numpy()
        cache_dir = cache_dir or os.environ.get("EDIS_CACHE_DIR")
        if cache_dir:
            key = cls.content_key(adj.shape[0], idx[0], idx[1], max_chunk)
            g = cls.cached(os.path.join(cache_dir, "graph_%016x.edisg" % key), adj.shape[0], idx[0], idx[1],
                           device=adj.device, max_chunk=max_chunk)
        else:
            g = cls(adj.shape[0], idx[0], idx[1], device=adj.device, max_chunk=max_chunk)
        try:
            adj._edis_graph = g
        except Exception:  # pragma: no cover
            pass
        return g


def as_graph(adj):
    return adj if isinstance(adj, Graph) else Graph.from_sparse(adj)
