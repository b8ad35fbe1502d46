"""INTEGRATION.md section 2 is a ctypes stub "a reference maintainer would add".  This test extracts
that code block from the document and runs it AS WRITTEN (no `_lib.py`), then compares its output
with the package's own layer on the same inputs -- so the document cannot drift from include/edis.h."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def doc_stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "edis_disga_fwd" in b]
    assert len(stub) == 1, "INTEGRATION.md must hold exactly one ctypes stub calling edis_disga_fwd"
    return stub[0]


def test_doc_stub_declares_the_whole_descriptor():
    """CPU: the stub's LayerDesc has the header's nine fields / 40 bytes (it asserts so itself)."""
    src = doc_stub_source()
    head = src.split("def graph_from_adj")[0].replace('ctypes.CDLL("edgedisentangle_ssl_b200/libedis.so")',
                                                      "ctypes.CDLL(%r)" % os.path.join(ROOT, "edgedisentangle_ssl_b200", "libedis.so"))
    ns = {}
    exec(head, ns)
    import ctypes
    assert ctypes.sizeof(ns["LayerDesc"]) == 40
    assert [f[0] for f in ns["LayerDesc"]._fields_] == ["att", "C", "D", "Dv", "training", "p", "seed", "flags", "reserved"]


@pytest.mark.gpu
def test_doc_stub_runs_and_matches_the_package_layer():
    import edgedisentangle_ssl_b200 as edis
    from edgedisentangle_ssl_b200.layers import run_channels
    from oracle import graph as og
    src = doc_stub_source().replace('ctypes.CDLL("edgedisentangle_ssl_b200/libedis.so")',
                                    "ctypes.CDLL(%r)" % os.path.join(ROOT, "edgedisentangle_ssl_b200", "libedis.so"))
    ns = {}
    exec(src, ns)
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(0)
    n, fin, C, D = 700, 48, 4, 64
    idx, val = og.build_adjacency(n, rng.randint(0, n, 6000), rng.randint(0, n, 6000))
    adj = torch.sparse_coo_tensor(torch.from_numpy(idx), torch.from_numpy(val), (n, n)).to(dev)
    torch.manual_seed(0)
    chs = [edis.DisGALayer(fin, D, 0.0, 0.1, True, 3, "AT").to(dev).eval() for _ in range(C)]
    x = torch.randn(n, fin, device=dev)
    w = torch.cat([torch.cat([l.W[:fin] for l in chs], 1), torch.cat([l.W[fin:] for l in chs], 1),
                   torch.cat([l.W_em for l in chs], 1)], 1)
    proj = x @ w
    a = torch.cat([l.a.reshape(1, D) for l in chs], 0).contiguous()
    h = ns["graph_from_adj"](adj)
    try:
        out, edge_e = ns["disga_forward"](h, n, idx.shape[1], proj, a, C, D, False, 0.0, 0)
        torch.cuda.synchronize()
    finally:
        ns["lib"].edis_graph_destroy(h)
    with torch.no_grad():
        ref_out, ref_e, _ = run_channels(chs, x, edis.Graph.from_sparse(adj))
    assert float((out - ref_out).abs().max() / ref_out.abs().max()) < 1e-5
    assert float((edge_e - ref_e).abs().max() / ref_e.abs().max()) < 1e-5
