"""Device-side validation metrics (trainer.roc_f_device) against sklearn, the reference's own
metric code path (utils.Roc_F, /root/reference/utils.py:258-284).  Pure torch: runs on CPU here."""
import pytest
import torch

from edgedisentangle_ssl_b200.trainer import roc_f, roc_f_device


@pytest.mark.parametrize("n,k,ties", [(300, 7, False), (500, 70, False), (200, 5, True), (120, 2, False),
                                      (150, 2, True), (64, 3, True)])
def test_roc_f_device_matches_sklearn(n, k, ties):
    torch.manual_seed(n + k)
    out = torch.randn(n, k)
    if ties:
        out = (out * 2).round() / 2          # many tied scores: average ranks must equal the trapezoid
    y = torch.randint(0, k, (n,))
    y[:k] = torch.arange(k)                  # sklearn's multi-class AUC needs every class present
    auc, f1 = roc_f_device(out, y)
    auc_ref, f1_ref = roc_f(out, y)
    assert abs(float(auc) - auc_ref) < 1e-9
    assert abs(float(f1) - f1_ref) < 1e-12


def test_macro_f1_ignores_classes_absent_from_truth_and_prediction():
    out = torch.tensor([[5.0, 0.0, 0.0, 0.0], [0.0, 5.0, 0.0, 0.0], [5.0, 0.0, 0.0, 0.0]])
    y = torch.tensor([0, 1, 1])
    _, f1 = roc_f_device(out, y)
    from sklearn.metrics import f1_score
    assert abs(float(f1) - f1_score(y, out.argmax(-1), average="macro")) < 1e-12


@pytest.mark.parametrize("rt", [0, 1, 2])
@pytest.mark.parametrize("use_res", [0, 1])
@pytest.mark.parametrize("no_relu", [0, 1])
def test_fuse_layer_variants_match_reference_golden(rt, use_res, no_relu):
    """layers.FuseLayer (layers.py:876-921): every --residue_type x residue x --fuse_no_relu variant,
    same state_dict keys, outputs and gradients as the unmodified reference
    (tests/golden/make_golden_fuse.py).  Pure torch at this size: runs on CPU."""
    import os
    import numpy as np
    import edgedisentangle_ssl_b200 as edis
    from edgedisentangle_ssl_b200.utils import get_parser
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuse_variants.npz"))
    tag = "rt%d_res%d_norelu%d" % (rt, use_res, no_relu)
    args = get_parser().parse_args(["--residue_type=%d" % rt] + (["--fuse_no_relu"] if no_relu else []))
    x = [torch.from_numpy(a).clone().requires_grad_(True) for a in g["x"]]
    r = torch.from_numpy(g["r"])
    f = edis.FuseLayer(args, len(x), nfeat=x[0].shape[1], residue=r.shape[1] if use_res else 0)
    ref_sd = {k[len(tag) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + ".sd.")}
    assert set(ref_sd) == set(f.state_dict())
    f.load_state_dict(ref_sd)
    y = f(x, r if use_res else None)
    y.pow(2).sum().backward()
    assert torch.allclose(y.detach(), torch.from_numpy(g[tag + ".y"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(x[0].grad, torch.from_numpy(g[tag + ".gx0"]), rtol=1e-5, atol=1e-7)
    for k, v in f.named_parameters():
        ref = torch.from_numpy(g[tag + ".g." + k])
        assert torch.allclose(v.grad, ref, rtol=1e-5, atol=1e-6), k
    # the channel-fused [N, C*D] tensor is accepted directly (the cat is free on the B200 path)
    y2 = f(torch.cat([t.detach() for t in x], 1), r if use_res else None)
    assert torch.allclose(y2, y.detach(), rtol=0, atol=0)


def test_disedge_conformt_label_sets_match_reference_golden():
    """--conformT (pretrainer.py:466-506): homo / hetero edge sets restricted to edges between
    label-known nodes, bit-exact with the unmodified reference on bundled chameleon, including the
    position it leaves python's RNG at (tests/golden/make_golden_conformt.py).  Host-side logic."""
    import contextlib
    import io
    import os
    import random
    import numpy as np
    from edgedisentangle_ssl_b200 import data_load, utils
    from edgedisentangle_ssl_b200.trainer import disedge_label_sets
    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, "golden", "conformt_chameleon.npz"))
    args = utils.get_parser().parse_args(["--model=DISGAT", "--sparse", "--dataset=chameleon", "--conformT"])
    args.hetero = False
    with contextlib.redirect_stdout(io.StringIO()):
        adj, _, labels = data_load.load_data(args, path=os.path.join(os.path.dirname(here), "data", "chameleon") + "/",
                                             dataset="chameleon", edge_type=1)
        idx = adj.coalesce().indices().numpy()
        random.seed(11)
        homo, het = disedge_label_sets(adj.shape[0], idx, labels, True, args.node_sup_ratio)
    assert np.array_equal(homo, g["set0"]) and np.array_equal(het, g["set1"])
    assert random.random() == float(g["py_random_after"])
    # without --conformT every edge takes part (pretrainer.py:440-456)
    homo_all, het_all = disedge_label_sets(adj.shape[0], idx, labels, False)
    assert homo_all.shape[1] + het_all.shape[1] == idx.shape[1]


def test_group_correlation_is_pearson_between_rows():
    """utils.group_correlation (utils.py:326-334) == numpy's corrcoef of the rows."""
    import numpy as np
    from edgedisentangle_ssl_b200.utils import group_correlation
    torch.manual_seed(3)
    e = torch.randn(7, 400, dtype=torch.float64)
    assert np.allclose(group_correlation(e).numpy(), np.corrcoef(e.numpy()), atol=1e-12)
