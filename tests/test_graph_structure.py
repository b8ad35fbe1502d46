"""CPU tests of the CSR / CSC graph builder (edis_graph_create[_rect], include/edis.h) through a
structure-only handle (device -1: host mirrors and schedules, nothing uploaded): the edge order is
the reference's `adj.coalesce().indices()` order bit for bit (layers.py:344), the permutation maps
input entries to coalesced slots, CSC is the exact transpose with CSR slot ids, long rows are cut
into chunks.  The GPU tests use the same builder with a device ordinal."""
import ctypes
import os

import numpy as np
import pytest
import torch

from edgedisentangle_ssl_b200 import _lib
from edgedisentangle_ssl_b200._lib import check, lib, np_ptr
from oracle import graph as og


class HostGraph:
    def __init__(self, n, row, col, max_chunk=0, n_cols=None):
        row = np.ascontiguousarray(row, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int64)
        self.h = ctypes.c_void_p()
        nc = n if n_cols is None else n_cols
        check(lib.edis_graph_create_rect(n, nc, row.shape[0], np_ptr(row, ctypes.c_int64), np_ptr(col, ctypes.c_int64),
                                         max_chunk, -1, ctypes.byref(self.h)), "edis_graph_create_rect")
        info = (ctypes.c_int64 * 10)()
        check(lib.edis_graph_info(self.h, info), "edis_graph_info")
        keys = ("n", "e", "dst_items", "dst_slots", "src_items", "src_slots", "max_in", "max_out", "was_sorted", "n_cols")
        self.info = dict(zip(keys, [int(v) for v in info]))
        e = self.info["e"]
        self.rowptr = np.empty(n + 1, np.int64)
        self.col = np.empty(max(e, 1), np.int32)
        self.perm = np.empty(max(row.shape[0], 1), np.int64)
        self.cscptr = np.empty(nc + 1, np.int64)
        self.cscrow = np.empty(max(e, 1), np.int32)
        self.csceid = np.empty(max(e, 1), np.int32)
        check(lib.edis_graph_export(self.h, np_ptr(self.rowptr, ctypes.c_int64), np_ptr(self.col, ctypes.c_int32),
                                    np_ptr(self.perm, ctypes.c_int64), np_ptr(self.cscptr, ctypes.c_int64),
                                    np_ptr(self.cscrow, ctypes.c_int32), np_ptr(self.csceid, ctypes.c_int32)),
              "edis_graph_export")
        self.col, self.cscrow, self.csceid = self.col[:e], self.cscrow[:e], self.csceid[:e]
        self.perm = self.perm[:row.shape[0]]

    def __del__(self):
        lib.edis_graph_destroy(self.h)

    def rows(self):
        return np.repeat(np.arange(len(self.rowptr) - 1), np.diff(self.rowptr))


def check_csc(g):
    """CSC == transpose of CSR: slot k of column c names a CSR slot whose column is c; rows ascend."""
    e = g.info["e"]
    rows = g.rows()
    assert np.array_equal(np.sort(g.csceid), np.arange(e))
    cols_of_slots = np.repeat(np.arange(len(g.cscptr) - 1), np.diff(g.cscptr))
    assert np.array_equal(g.col[g.csceid], cols_of_slots)
    assert np.array_equal(rows[g.csceid], g.cscrow)
    same_col = cols_of_slots[1:] == cols_of_slots[:-1]
    assert np.all(g.cscrow[1:][same_col] > g.cscrow[:-1][same_col])
    assert g.info["max_in"] == int(np.diff(g.rowptr).max()) and g.info["max_out"] == int(np.diff(g.cscptr).max())


def test_sorted_input_keeps_the_reference_edge_order():
    rng = np.random.RandomState(0)
    n = 500
    idx, _ = og.build_adjacency(n, rng.randint(0, n, 4000), rng.randint(0, n, 4000))
    g = HostGraph(n, idx[0], idx[1])
    assert g.info["was_sorted"] == 1 and g.info["e"] == idx.shape[1]
    assert np.array_equal(g.perm, np.arange(idx.shape[1]))
    assert np.array_equal(g.rows(), idx[0]) and np.array_equal(g.col, idx[1])
    check_csc(g)


def test_unsorted_input_with_duplicates_is_coalesced_like_torch():
    rng = np.random.RandomState(1)
    n, m = 300, 5000
    row, col = rng.randint(0, n, m), rng.randint(0, n, m)        # plenty of duplicates, random order
    g = HostGraph(n, row, col)
    ref = torch.sparse_coo_tensor(torch.from_numpy(np.stack([row, col])), torch.ones(m), (n, n)).coalesce().indices().numpy()
    assert g.info["was_sorted"] == 0 and g.info["e"] == ref.shape[1]
    assert np.array_equal(g.rows(), ref[0]) and np.array_equal(g.col, ref[1])
    # perm: input entry k lives in CSR slot perm[k] (duplicates share a slot)
    assert np.array_equal(ref[0][g.perm], row) and np.array_equal(ref[1][g.perm], col)
    check_csc(g)


@pytest.mark.parametrize("max_chunk", [4, 32, 0])
def test_schedule_cuts_long_rows_into_chunks(max_chunk):
    rng = np.random.RandomState(2)
    n = 200
    hub = rng.randint(0, 3, 1500)                                  # three hub rows
    idx, _ = og.build_adjacency(n, np.concatenate([hub, rng.randint(0, n, 600)]), rng.randint(0, n, 2100))
    g = HostGraph(n, idx[0], idx[1], max_chunk=max_chunk)
    chunk = max_chunk or 256
    deg = np.diff(g.rowptr)
    nch = np.where(deg > chunk, -(-deg // chunk), 0)               # chunks of the rows that are split
    assert g.info["dst_slots"] == int(nch.sum())
    assert g.info["dst_items"] == int(nch.sum() + (deg <= chunk).sum())
    odeg = np.diff(g.cscptr)
    onch = np.where(odeg > chunk, -(-odeg // chunk), 0)
    assert g.info["src_slots"] == int(onch.sum()) and g.info["src_items"] == int(onch.sum() + (odeg <= chunk).sum())
    ws = lib.edis_graph_workspace_bytes(g.h, 512)
    assert ws >= max(g.info["dst_slots"], g.info["src_slots"]) * 512 * 4


def test_rectangular_partition_slice():
    """Destination-range slice: own rows x (own + halo) columns (parallel.compact_columns)."""
    from edgedisentangle_ssl_b200 import parallel as par
    rng = np.random.RandomState(3)
    n = 400
    idx, _ = og.build_adjacency(n, rng.randint(0, n, 3000), rng.randint(0, n, 3000))
    lo, hi = 100, 250
    sel = (idx[0] >= lo) & (idx[0] < hi)
    row_l, col_l, halo = par.compact_columns(lo, hi, idx[0][sel], idx[1][sel])
    g = HostGraph(hi - lo, row_l, col_l, n_cols=hi - lo + len(halo))
    assert g.info["n"] == hi - lo and g.info["n_cols"] == hi - lo + len(halo) and g.info["e"] == int(sel.sum())
    # the local CSR order is still the global row-major order of the slice
    glob_col = np.where(g.col < hi - lo, g.col + lo, halo[np.maximum(g.col - (hi - lo), 0)])
    assert np.array_equal(g.rows() + lo, idx[0][sel]) and np.array_equal(np.sort(glob_col), np.sort(idx[1][sel]))
    check_csc(g)


def test_bundled_graphs_through_the_builder_match_the_golden_hashes():
    """cora / cora_full / chameleon: load_data -> CSR; rowptr / col reproduce the reference's
    coalesced indices recorded in tests/golden/graph_bundled.npz."""
    import contextlib
    import hashlib
    import io
    from edgedisentangle_ssl_b200 import data_load, utils
    here = os.path.dirname(os.path.abspath(__file__))
    gold = np.load(os.path.join(here, "golden", "graph_bundled.npz"), allow_pickle=True)
    for ds in ("cora", "cora_full", "chameleon"):
        args = utils.get_parser().parse_args(["--model=DISGAT", "--sparse", "--dataset=" + ds])
        args.hetero = False
        with contextlib.redirect_stdout(io.StringIO()):
            adj, _, _ = data_load.load_data(args, path=os.path.join(os.path.dirname(here), "data", ds) + "/", dataset=ds,
                                            edge_type=1)
        idx = adj.coalesce().indices().numpy()
        g = HostGraph(adj.shape[0], idx[0], idx[1])
        rebuilt = np.stack([g.rows().astype(np.int64), g.col.astype(np.int64)])
        assert np.array_equal(rebuilt, idx)
        assert hashlib.sha256(np.ascontiguousarray(rebuilt).tobytes()).hexdigest() == str(gold[ds + "_idx_sha"])
        assert g.info["e"] == int(gold[ds + "_e"]) and g.info["n"] == int(gold[ds + "_n"])
        check_csc(g)


def test_ops_reject_a_structure_only_handle():
    g = HostGraph(3, [0, 1, 2], [0, 1, 2])
    d = _lib.LayerDesc(att=3, C=2, D=64, Dv=64, training=0, p=0.0, seed=0)
    one = ctypes.c_void_p(16)          # non-null dummies: the call must fail before touching them
    rc = lib.edis_disga_fwd(g.h, ctypes.byref(d), one, 128, one, 128, one, one, 128, None, one, one, one, one, None,
                            one, 1 << 20, None)
    assert rc < 0 and "structure-only" in _lib.last_error()


def test_edge_cases_empty_single_row_and_bad_input():
    g = HostGraph(5, np.zeros(0, np.int64), np.zeros(0, np.int64))          # no edges at all
    assert g.info["e"] == 0 and np.array_equal(g.rowptr, np.zeros(6, np.int64)) and g.info["dst_slots"] == 0
    g = HostGraph(1, [0], [0])                                              # one node, one self loop
    assert g.info["e"] == 1 and g.info["max_in"] == 1 and g.col[0] == 0
    n = 300                                                                 # every edge in ONE row / ONE column
    g = HostGraph(n, np.full(n, 7), np.arange(n), max_chunk=16)
    assert g.info["max_in"] == n and g.info["dst_slots"] == -(-n // 16) and g.info["max_out"] == 1
    check_csc(g)
    h = ctypes.c_void_p()
    bad = np.array([0, 5], dtype=np.int64)
    rc = lib.edis_graph_create_rect(3, 3, 2, np_ptr(bad, ctypes.c_int64), np_ptr(bad, ctypes.c_int64), 0, -1,
                                    ctypes.byref(h))
    assert rc < 0 and "out of range" in _lib.last_error()
    rc = lib.edis_graph_create_rect(3, 2, 0, None, None, 0, -1, ctypes.byref(h))   # n_cols < n_rows
    assert rc < 0


def test_adjacency_builder_edge_cases():
    """edis_build_adjacency_host (data_load.py:39-77): empty edge list -> identity; duplicates keep
    the max; an explicit zero-valued entry and a negative one vanish under max(A, A^T) >= 0; the
    diagonal is overwritten with 1 before the row normalisation."""
    from edgedisentangle_ssl_b200.graph import build_adjacency
    idx, val = build_adjacency(4, np.zeros(0, np.int64), np.zeros(0, np.int64))
    assert np.array_equal(idx, np.stack([np.arange(4)] * 2)) and np.allclose(val, 1.0)
    rows = np.array([0, 0, 1, 2, 2, 3], dtype=np.int64)
    cols = np.array([1, 1, 0, 3, 2, 0], dtype=np.int64)
    vals = np.array([0.5, 2.0, 1.0, 0.0, 9.0, -1.0])
    idx, val = build_adjacency(4, rows, cols, vals)
    dense = np.zeros((4, 4))                                                # the reference's dense recipe
    seen = set()
    for r, c, v in zip(rows, cols, vals):
        dense[r, c] = max(dense[r, c], v) if (r, c) in seen else v          # duplicates keep the max
        seen.add((r, c))
    np.fill_diagonal(dense, 1)                                              # data_load.py:69
    dense = np.maximum(dense, dense.T)                                      # data_load.py:71
    dense = dense / dense.sum(1, keepdims=True)                             # normalize_adj
    got = np.zeros((4, 4))
    got[idx[0], idx[1]] = val
    assert np.allclose(got, dense, atol=1e-7) and np.count_nonzero(got) == idx.shape[1]


# ------------------------------------------------------------------ on-disk cache
def _load_host(path, key, max_chunk=0, verify=1):
    h = ctypes.c_void_p()
    rc = lib.edis_graph_load(os.fsencode(path), key, max_chunk, -1, verify, ctypes.byref(h))
    return rc, h


def _export(h, n, nc, e, e_in):
    rowptr, col, perm = np.empty(n + 1, np.int64), np.empty(max(e, 1), np.int32), np.empty(max(e_in, 1), np.int64)
    cscptr, cscrow, csceid = np.empty(nc + 1, np.int64), np.empty(max(e, 1), np.int32), np.empty(max(e, 1), np.int32)
    check(lib.edis_graph_export(h, np_ptr(rowptr, ctypes.c_int64), np_ptr(col, ctypes.c_int32),
                                np_ptr(perm, ctypes.c_int64), np_ptr(cscptr, ctypes.c_int64),
                                np_ptr(cscrow, ctypes.c_int32), np_ptr(csceid, ctypes.c_int32)), "edis_graph_export")
    return rowptr, col[:e], perm[:e_in], cscptr, cscrow[:e], csceid[:e]


@pytest.mark.parametrize("shuffle,max_chunk", [(False, 0), (True, 16)])
def test_cache_round_trip_is_identical(tmp_path, shuffle, max_chunk):
    """edis_graph_save -> edis_graph_load (memory-mapped) reproduces every array, the schedules' sizes
    and the input -> slot permutation, for sorted and unsorted (+ duplicate) inputs and split rows."""
    rng = np.random.RandomState(5)
    n = 400
    idx, _ = og.build_adjacency(n, np.concatenate([rng.randint(0, n, 3000), np.zeros(300, np.int64)]),
                                np.concatenate([rng.randint(0, n, 3000), rng.randint(0, n, 300)]))
    row, col = idx[0], idx[1]
    if shuffle:
        o = rng.permutation(len(row))
        row, col = np.concatenate([row[o], row[:50]]), np.concatenate([col[o], col[:50]])    # unsorted + duplicates
    g = HostGraph(n, row, col, max_chunk=max_chunk)
    key = int(lib.edis_edge_list_key(n, n, len(row), np_ptr(np.ascontiguousarray(row), ctypes.c_int64),
                                     np_ptr(np.ascontiguousarray(col), ctypes.c_int64), max_chunk))
    path = str(tmp_path / "g.edisg")
    check(lib.edis_graph_save(g.h, os.fsencode(path), key), "edis_graph_save")
    rc, h = _load_host(path, key, max_chunk)
    assert rc == 0
    try:
        info = (ctypes.c_int64 * 10)()
        check(lib.edis_graph_info(h, info), "edis_graph_info")
        keys = ("n", "e", "dst_items", "dst_slots", "src_items", "src_slots", "max_in", "max_out", "was_sorted", "n_cols")
        assert dict(zip(keys, [int(v) for v in info])) == g.info
        assert int(lib.edis_graph_input_entries(h)) == len(row)
        got = _export(h, n, n, g.info["e"], len(row))
        for a, b in zip(got, (g.rowptr, g.col, g.perm, g.cscptr, g.cscrow, g.csceid)):
            assert np.array_equal(a, b)
        if max_chunk:
            assert g.info["dst_slots"] > 0            # the hub row 0 was split: schedules with slots round-trip
    finally:
        lib.edis_graph_destroy(h)


def test_cache_rejects_stale_and_damaged_files(tmp_path):
    """A cache written for other input (key), another chunk size, a truncated or bit-flipped file, or a
    missing one is EDIS_ERR_STALE -- never silently used (the hazard of pretrainer.py:390-398)."""
    rng = np.random.RandomState(6)
    n = 120
    idx, _ = og.build_adjacency(n, rng.randint(0, n, 700), rng.randint(0, n, 700))
    g = HostGraph(n, idx[0], idx[1])
    key = int(lib.edis_edge_list_key(n, n, idx.shape[1], np_ptr(np.ascontiguousarray(idx[0]), ctypes.c_int64),
                                     np_ptr(np.ascontiguousarray(idx[1]), ctypes.c_int64), 0))
    other = idx.copy()
    other[1, 5] = (other[1, 5] + 1) % n                 # one changed edge -> another key
    key2 = int(lib.edis_edge_list_key(n, n, idx.shape[1], np_ptr(np.ascontiguousarray(other[0]), ctypes.c_int64),
                                      np_ptr(np.ascontiguousarray(other[1]), ctypes.c_int64), 0))
    assert key != key2
    path = str(tmp_path / "g.edisg")
    check(lib.edis_graph_save(g.h, os.fsencode(path), key), "edis_graph_save")
    assert _load_host(path, key)[0] == 0
    assert _load_host(path, key2)[0] == _lib.ERR_STALE and "stale" in _lib.last_error()
    assert _load_host(path, key, max_chunk=64)[0] == _lib.ERR_STALE
    assert _load_host(str(tmp_path / "missing.edisg"), key)[0] == _lib.ERR_STALE
    blob = open(path, "rb").read()
    open(str(tmp_path / "cut.edisg"), "wb").write(blob[: len(blob) // 2])
    assert _load_host(str(tmp_path / "cut.edisg"), key)[0] == _lib.ERR_STALE
    flipped = bytearray(blob)
    flipped[len(blob) - 40] ^= 0x10
    open(str(tmp_path / "flip.edisg"), "wb").write(bytes(flipped))
    assert _load_host(str(tmp_path / "flip.edisg"), key)[0] == _lib.ERR_STALE and "checksum" in _lib.last_error()
    assert _load_host(str(tmp_path / "flip.edisg"), key, verify=0)[0] == 0       # checksum is opt-out
