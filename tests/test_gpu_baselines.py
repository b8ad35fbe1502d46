"""Sparse GAT / FactorGCN layers (edgedisentangle_ssl_b200/baselines.py) on the COO softmax + SpMM kernels
against golden vectors of the unmodified reference layers (tests/golden/make_golden_baselines.py):
outputs, input gradient and every parameter gradient; a reference state_dict loads strict."""
import pytest
import torch

from edgedisentangle_ssl_b200 import baselines
from helpers import load, t, assert_close, params_from, group_floor

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", ["gat", "gat_nc", "factor"])
def test_baseline_layer_vs_reference_golden(name):
    g = load("baseline_layers")
    n = int(g["n"])
    fin, dout = g["x"].shape[1], g[name + ".out"].shape[1]
    if name == "factor":
        lay = baselines.DisentangleLayer(fin, dout, concat=True, n_latent=4)
    else:
        lay = baselines.GraphAttentionLayer(fin, dout, dropout=0.3, alpha=0.2, concat=(name == "gat"))
    lay.load_state_dict(params_from(g, name + ".p."), strict=True)
    lay = lay.to(DEV).eval()
    idx = torch.as_tensor(g["indices"])
    adj = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (n, n)).to(DEV)
    x = t(g["x"]).to(DEV).requires_grad_(True)
    y = lay(x, adj)
    assert_close(y.cpu(), g[name + ".out"], 1e-5, "out")
    (y * t(g[name + ".r"]).to(DEV)).sum().backward()
    assert_close(x.grad.cpu(), g[name + ".gx"], 2e-5, "gx")
    floor = group_floor([v for k, v in g.items() if k.startswith(name + ".g.")])
    for k, prm in lay.named_parameters():
        assert_close(prm.grad.cpu(), g[name + ".g." + k], 2e-5, "g." + k, floor)


def test_cpu_input_is_refused():
    lay = baselines.GraphAttentionLayer(4, 4, 0.0, 0.2)
    with pytest.raises(Exception):
        lay(torch.randn(3, 4), torch.eye(3).to_sparse())
