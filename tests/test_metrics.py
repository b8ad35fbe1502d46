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
