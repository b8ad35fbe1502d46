"""CPU tests: the C-ABI library loads, exports every symbol include/edis.h declares, and the
host-side (integer) entry points are bit-exact against the reference goldens.  No GPU calls."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch

import edgedisentangle_ssl_b200 as edis
from edgedisentangle_ssl_b200 import _lib
from edgedisentangle_ssl_b200.sampler import sample_pairs, homo_hetero_split
from helpers import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "edis.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(edis_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = declared_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libedis.so does not export %s" % s
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert "sm_100a" in _lib.version()


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_ops_fail_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.EdisError):
        edis.Graph(3, [0, 1, 2], [0, 1, 2], device="cuda:0")
    with pytest.raises(_lib.EdisError):
        edis.Graph(3, [0, 1, 2], [0, 1, 2], device="cpu")
    lay = edis.DisGALayer(4, 8, 0.0, 0.1, att_type=2)
    adj = torch.sparse_coo_tensor(torch.tensor([[0, 1], [1, 0]]), torch.ones(2), (2, 2))
    with pytest.raises(_lib.EdisError):
        lay(torch.randn(2, 4), adj)


def test_bad_arguments_report_errors():
    r = _lib.lib.edis_build_adjacency_host(0, 0, None, None, None, None, None, None)
    assert r < 0 and "bad arguments" in _lib.last_error()
    with pytest.raises(_lib.EdisError):
        edis.build_adjacency(3, [0, 5], [1, 1])  # out of range


# ------------------------------------------------------------------ graph builder (bit-exact)
def test_build_adjacency_edge_list():
    g = load("graph_small")
    e = g["el_edges"]
    idx, val = edis.build_adjacency(int(e.max()) + 1, e[:, 0], e[:, 1])
    assert np.array_equal(idx, g["el_indices"]) and np.array_equal(val, g["el_values"])


def test_build_adjacency_weighted_csr():
    g = load("graph_small")
    idx, val = edis.build_adjacency(int(g["csr_n"]), g["csr_row"], g["csr_col"], g["csr_val"])
    assert np.array_equal(idx, g["csr_indices"]) and np.array_equal(val, g["csr_values"])


def test_build_adjacency_empty_and_isolated():
    idx, val = edis.build_adjacency(4, [], [])
    assert np.array_equal(idx, np.stack([np.arange(4), np.arange(4)])) and np.all(val == 1.0)


@pytest.mark.parametrize("ds", ["cora", "chameleon", "cora_full"])
def test_build_adjacency_bundled(ds):
    from edgedisentangle_ssl_b200.data_load import load_graph_arrays
    g = load("graph_bundled")
    n, r, c, v = load_graph_arrays(os.path.join(ROOT, "data", ds) + "/")
    idx, val = edis.build_adjacency(n, r, c, v)
    assert n == int(g[ds + "_n"]) and idx.shape[1] == int(g[ds + "_e"])
    assert sha(idx) == str(g[ds + "_idx_sha"]) and sha(val) == str(g[ds + "_val_sha"])


# ------------------------------------------------------------------ sampler (bit-exact)
def seed_all(s):
    import random
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


@pytest.mark.parametrize("tag", ["model_a3_AT", "model_a2_GCN"])
def test_sampler_small(tag):
    g = load(tag)
    n, idx = int(g["n"]), g["indices"]
    seed_all(8)
    pairs, lab = sample_pairs(n, idx, chunk_rows=7)
    assert np.array_equal(pairs, g["sup.sample_idx"]) and np.array_equal(lab, g["sup.sample_lab"])
    homo, het = homo_hetero_split(idx, g["labels"])
    seed_all(9)
    p0, l0 = sample_pairs(n, homo)
    p1, l1 = sample_pairs(n, het, chunk_rows=5)
    assert np.array_equal(p0, g["dis.sample_idx0"]) and np.array_equal(l0, g["dis.sample_lab0"])
    assert np.array_equal(p1, g["dis.sample_idx1"]) and np.array_equal(l1, g["dis.sample_lab1"])


def test_sampler_cora():
    from edgedisentangle_ssl_b200.data_load import load_graph_arrays
    g = load("graph_bundled")
    n, r, c, v = load_graph_arrays(os.path.join(ROOT, "data", "cora") + "/")
    idx, _ = edis.build_adjacency(n, r, c, v)
    seed_all(4)
    pairs, lab = sample_pairs(n, idx)
    assert pairs.shape[1] == int(g["cora_sample_m"])
    assert sha(pairs) == str(g["cora_sample_idx_sha"]) and sha(lab) == str(g["cora_sample_lab_sha"])


@pytest.mark.parametrize("seed,n,thr", [(0, 1000, 0.013), (4, 3001, 0.00217), (11, 50, 0.5), (3, 700, 1e-9),
                                        (5, 333, 1.0), (9, 1, 0.3)])
def test_rand_hits_replays_torch_cpu_generator_bit_exactly(seed, n, thr):
    """edis_rand_hits_host == (torch.rand(N, N) < thr).nonzero() on the CPU default generator,
    hit for hit, and leaves the generator in the same state (the reference samples its SSL pairs
    from exactly this stream, pretrainer.py:692)."""
    import numpy as np
    import torch
    from edgedisentangle_ssl_b200.sampler import bernoulli_hits
    torch.manual_seed(seed)
    torch.rand(seed * 7 + 3)                      # start in the middle of a 624-word block
    st0 = torch.get_rng_state().clone()
    ref = (torch.rand(size=(n, n)) < thr).nonzero()
    ref_key = (ref[:, 0] * n + ref[:, 1]).numpy()
    after_ref = torch.rand(7)
    torch.set_rng_state(st0)
    got = bernoulli_hits(n * n, thr)
    assert got is not None and np.array_equal(got, ref_key)
    assert torch.equal(torch.rand(7), after_ref)
