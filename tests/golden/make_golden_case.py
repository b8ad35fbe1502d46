"""Golden fixture for the --case study (`Trainer.analyze_disentangle`, trainer.py:82-134 of the
UNMODIFIED reference) on the model / graph of tests/golden/model_a3_AT.npz, eval mode.
Build-container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_case.py  ->  tests/golden/case_a3_AT.npz
"""
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference"]

import models  # noqa: E402  (reference)
import trainer  # noqa: E402
import utils  # noqa: E402


def main():
    g = np.load(os.path.join(HERE, "model_a3_AT.npz"))
    args = utils.get_parser().parse_args([str(a) for a in g["argv"]])
    args.cuda, args.hetero, args.edge_num = False, True, 1
    args.size = g["x"].shape[1]
    x, labels = torch.from_numpy(g["x"]), torch.from_numpy(g["labels"])
    args.nclass = int(labels.max()) + 1
    n = int(g["n"])
    idx = torch.from_numpy(g["indices"])
    adj = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (n, n)).coalesce()
    random.seed(4)
    enc = models.DISGAT(args, nfeat=args.size, nhid=args.nhid, nclass=args.nhid, nheads=args.nhead, dropout=args.dropout)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        cls = trainer.ClsTrainer(args, enc, labels, 1.0)
    finally:
        os.chdir(cwd)
    enc.load_state_dict({k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("enc0.")})
    cls.fuse1.load_state_dict({k[11:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("cls0.fuse1.")})
    cls.fuse2.load_state_dict({k[11:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("cls0.fuse2.")})
    for m in cls.models:
        m.eval()
    torch.manual_seed(21)
    np.random.seed(21)
    with torch.no_grad():
        dist, at_cor, feat_cor = cls.analyze_disentangle(x, adj)
    out = {"at_distance": np.array(dist, dtype=np.float64), "torch_rand_after": torch.rand(3).numpy()}
    for l in range(2):
        out["at_cor%d" % l] = at_cor[l].numpy()
        out["feat_cor%d" % l] = feat_cor[l].numpy()
    np.savez_compressed(os.path.join(HERE, "case_a3_AT.npz"), **out)
    print(dist, at_cor[0].shape, feat_cor[0].shape)


if __name__ == "__main__":
    main()
