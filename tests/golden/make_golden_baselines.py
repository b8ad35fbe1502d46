"""Golden vectors for the sparse GAT and FactorGCN layers (layers.py:229-296, 515-597) from the UNMODIFIED
reference -> tests/golden/baseline_layers.npz.  Build container only (needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_baselines.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference"]
sys.path.append(os.path.dirname(os.path.dirname(HERE)))

import layers  # noqa: E402  (reference)
from oracle import graph as og  # noqa: E402


def main():
    rng = np.random.RandomState(7)
    n, fin, dout = 120, 20, 32
    src, dst = rng.randint(0, n, 900), rng.randint(0, n, 900)
    dst[:60] = 2                                           # a hub row
    idx, val = og.build_adjacency(n, dst, src)
    adj = torch.sparse_coo_tensor(torch.from_numpy(idx), torch.from_numpy(val), (n, n)).coalesce()
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(n, fin, generator=gen)
    out = {"n": np.int64(n), "indices": idx, "x": x.numpy()}
    for name, make in (("gat", lambda: layers.GraphAttentionLayer(fin, dout, dropout=0.3, alpha=0.2, concat=True)),
                       ("gat_nc", lambda: layers.GraphAttentionLayer(fin, dout, dropout=0.3, alpha=0.2, concat=False)),
                       ("factor", lambda: layers.DisentangleLayer(fin, dout, concat=True, n_latent=4))):
        torch.manual_seed(11)
        lay = make().eval()
        xi = x.clone().requires_grad_(True)
        y = lay(xi, adj)
        r = torch.randn(y.shape, generator=gen)
        (y * r).sum().backward()
        out[name + ".out"] = y.detach().numpy()
        out[name + ".r"] = r.numpy()
        out[name + ".gx"] = xi.grad.numpy()
        for k, v in lay.state_dict().items():
            out[name + ".p." + k] = v.numpy().copy()
        for k, v in lay.named_parameters():
            out[name + ".g." + k] = v.grad.numpy().copy()
    path = os.path.join(HERE, "baseline_layers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
