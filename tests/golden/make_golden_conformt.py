"""Golden fixture for DisEdge's --conformT label sets (pretrainer.py:466-506 of the UNMODIFIED
reference): homo / hetero edge sets restricted to edges between label-known nodes, on bundled
chameleon.  Build-container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_conformt.py  ->  tests/golden/conformt_chameleon.npz
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

import data_load  # noqa: E402  (reference)
import models  # noqa: E402
import pretrainer  # noqa: E402
import utils  # noqa: E402


def main():
    argv = ["--model=DISGAT", "--sparse", "--att=3", "--gnn_type=AT", "--nhead=4", "--dataset=chameleon", "--conformT",
            "--pretrain", "DisEdge", "--pre_weight", "1", "--pre_edge", "1"]
    args = utils.get_parser().parse_args(argv)
    args.cuda = False
    args.hetero = False
    adj, feat, labels = data_load.load_data(args, path="/root/reference/data/chameleon/", dataset="chameleon", edge_type=1)
    args.size = feat.shape[1]
    args.nclass = int(labels.max()) + 1
    enc = models.DISGAT(args, nfeat=args.size, nhid=args.nhid, nclass=args.nhid, nheads=args.nhead, dropout=args.dropout)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())          # the trainer creates ./resource/<dataset>/
    try:
        random.seed(11)
        np.random.seed(11)
        torch.manual_seed(11)
        tr = pretrainer.GeneratedEdgeTrainer(args, enc, 1.0)
        dis = tr.get_label_all(feat, adj, labels)
        after = random.random()           # the python RNG position after the reference's utils.split call
    finally:
        os.chdir(cwd)
    out = {"py_random_after": np.float64(after)}
    for k, a in enumerate(dis):
        nz = a.nonzero().numpy()          # row-major
        out["set%d" % k] = nz.T.astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "conformt_chameleon.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
