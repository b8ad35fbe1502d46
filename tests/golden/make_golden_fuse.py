"""Golden fixture for every FuseLayer variant (layers.py:876-921 of the UNMODIFIED reference):
residue_type 0 / 1 / 2, with and without a residue input, with and without --fuse_no_relu.
Build-container only (imports /root/reference with the stubs of make_golden.py):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_fuse.py  ->  tests/golden/fuse_variants.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference"]

import layers  # noqa: E402  (reference)
import utils  # noqa: E402


def main():
    out = {}
    torch.manual_seed(0)
    n, heads, nfeat, res = 37, 4, 8, 5
    x = [torch.randn(n, nfeat) for _ in range(heads)]
    r = torch.randn(n, res)
    out["x"] = torch.stack(x).numpy()
    out["r"] = r.numpy()
    for rt in (0, 1, 2):
        for use_res in (0, 1):
            for no_relu in (0, 1):
                argv = ["--residue_type=%d" % rt] + (["--fuse_no_relu"] if no_relu else [])
                args = utils.get_parser().parse_args(argv)
                torch.manual_seed(100 + rt * 4 + use_res * 2 + no_relu)
                f = layers.FuseLayer(args, heads, nfeat=nfeat, residue=res if use_res else 0)
                xs = [t.clone().requires_grad_(True) for t in x]
                y = f(xs, r if use_res else None)
                y.pow(2).sum().backward()
                tag = "rt%d_res%d_norelu%d" % (rt, use_res, no_relu)
                out[tag + ".y"] = y.detach().numpy()
                out[tag + ".gx0"] = xs[0].grad.numpy()
                for k, v in f.state_dict().items():
                    out[tag + ".sd." + k] = v.numpy()
                for k, v in f.named_parameters():
                    out[tag + ".g." + k] = v.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "fuse_variants.npz"), **out)
    print("wrote fuse_variants.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
