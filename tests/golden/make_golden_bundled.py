"""Golden summaries of the UNMODIFIED reference on the BUNDLED graphs -> tests/golden/bundled_ref.npz.

Build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_bundled.py

For cora / cora_full (synthetic 64-d features of SURVEY 8(d): their feature blobs are missing from the
reference snapshot) with att 3 / AT, and chameleon (REAL features, |x| up to ~892: the hard case for
fp32 parity) with all nine att x gnn_type combinations, the reference's own `data_load.load_data`,
`models.DISGAT.{get_adjs, get_em, predict_adjs_sparse, get_edge_em}`, `utils.sp_softmax`,
`pretrainer.{SupEdgeTrainer, GeneratedEdgeTrainer}.sample_train`, `utils.adj_mse_loss` and `models.MLP`
produce, in eval mode: raw logits / alpha at 512 fixed edges, elu(h') and feature_2 at 64 fixed nodes,
alpha row-sum deviation, the sampled pair sets (count + sha256) and the three SSL losses.  The full
tensors are far too large to commit; the CUDA path and the oracle are compared against these samples
and, over ALL entries, against each other (tests/test_gpu_bundled.py).  Weights are not stored: both
sides build the model from the same torch seed in the same constructor order and the golden carries a
checksum of the reference's state dict.
"""
import hashlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path[:0] = [os.path.join(HERE, "_stubs"), REF]

import data_load  # noqa: E402  (reference)
import layers  # noqa: E402
import models  # noqa: E402
import pretrainer  # noqa: E402
import utils  # noqa: E402

torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", "4")))
SEED = 11
C, D = 4, 64


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synthetic_features(labels, dim=64, seed=0):
    rng = np.random.RandomState(seed)
    mu = rng.randn(int(labels.max()) + 1, dim)
    return np.abs(mu[labels] + rng.randn(labels.shape[0], dim))


def scratch(ds):
    tmp = tempfile.mkdtemp(prefix="gold_")
    src = os.path.join(REF, "data", ds)
    for f in os.listdir(src):
        os.symlink(os.path.join(src, f), os.path.join(tmp, f))
    if not os.path.exists(os.path.join(tmp, "feature_new.npy")):
        np.save(os.path.join(tmp, "feature_new.npy"), synthetic_features(np.load(os.path.join(src, "label.npy"))))
    return tmp + "/"


def ref_args(att, gnn):
    a = utils.get_parser().parse_args(["--model=DISGAT", "--sparse", "--att=%d" % att, "--gnn_type=" + gnn,
                                       "--nhead=%d" % C, "--nhid=%d" % D, "--dropout=0.1", "--no-cuda"])
    a.cuda, a.hetero = False, False
    return a


def build(att, gnn, fin):
    """Same constructor order as tests/test_gpu_bundled.py::build_edis."""
    args = ref_args(att, gnn)
    args.size = fin
    torch.manual_seed(SEED)
    enc = models.DISGAT(args, nfeat=fin, nhid=D, nclass=D, nheads=C, dropout=0.1)
    fus = [layers.FuseLayer(args, C, nfeat=D), layers.FuseLayer(args, C, nfeat=D)]
    clf = [models.MLP(in_feat=D + fin, hidden_size=D, out_size=C, layers=2),
           models.MLP(in_feat=2 * D, hidden_size=D, out_size=C, layers=2)]
    for m in [enc] + fus + clf:
        m.eval()
    return args, enc, fus, clf


def state_checksum(mods):
    tot = 0.0
    for m in mods:
        for v in m.state_dict().values():
            tot += float(v.double().abs().sum())
    return tot


def main():
    out = {}
    for ds in ("cora", "chameleon", "cora_full"):
        args0 = ref_args(3, "AT")
        adj, x, labels = data_load.load_data(args0, path=scratch(ds), dataset=ds, edge_type=1)
        adj = adj.coalesce()
        idx = adj.indices()
        n, e, fin = adj.shape[0], idx.shape[1], x.shape[1]
        rng = np.random.RandomState(5)
        sel_e = np.sort(rng.choice(e, 512, replace=False))
        sel_n = np.sort(rng.choice(n, 64, replace=False))
        out[ds + ".n"], out[ds + ".e"] = np.int64(n), np.int64(e)
        out[ds + ".sel_e"], out[ds + ".sel_n"] = sel_e, sel_n
        out[ds + ".indices_sha"] = np.array(sha(idx.numpy()))
        out[ds + ".x_sha"] = np.array(sha(x.numpy()))
        # ---- SSL pair sets from the reference's own samplers (RNG seeded right before each call)
        dense = adj.to_dense()
        label_sup = (dense != 0).float()
        torch.manual_seed(SEED)
        np.random.seed(SEED)
        y_sup, m_sup = pretrainer.SupEdgeTrainer.sample_train(None, label_sup)
        pairs_sup = m_sup[0]
        out[ds + ".sup_m"] = np.int64(pairs_sup.shape[1])
        out[ds + ".sup_pairs_sha"] = np.array(sha(pairs_sup.numpy()))
        out[ds + ".sup_label_sha"] = np.array(sha(y_sup.numpy()))
        homo = (labels.unsqueeze(0).expand(dense.shape) == labels.unsqueeze(-1).expand(dense.shape)).int()
        adj_ind = (dense != 0).int()
        dis = [((homo + adj_ind) == 2).float(), (((1 - homo) + adj_ind) == 2).float()]   # pretrainer.py:447-456
        del homo, adj_ind, dense
        stub = types.SimpleNamespace(dis_adjs=dis, args=types.SimpleNamespace(sparse=True))
        torch.manual_seed(SEED + 1)
        np.random.seed(SEED + 1)
        y_dis, m_dis = pretrainer.GeneratedEdgeTrainer.sample_train(stub)
        for k in range(2):
            out[ds + ".dis%d_m" % k] = np.int64(m_dis[k].shape[1])
            out[ds + ".dis%d_pairs_sha" % k] = np.array(sha(m_dis[k].numpy()))
            out[ds + ".dis%d_label_sha" % k] = np.array(sha(y_dis[k].numpy()))
        del dis, stub, label_sup
        combos = [(3, "AT")] if ds != "chameleon" else [(a, g) for a in (1, 2, 3) for g in ("AT", "SAGE", "GCN")]
        for att, gnn in combos:
            k = "%s.a%d_%s." % (ds, att, gnn)
            args, enc, fus, clf = build(att, gnn, fin)
            out[k + "state_checksum"] = np.float64(state_checksum([enc] + fus + clf))
            with torch.no_grad():
                edge_e = enc.get_adjs(x, adj, fus)                       # [[E,1]]_c per layer
                feats = enc.get_em(x, adj, fus)
                em = enc.get_edge_em(x, adj, fus)
                aux_sup = enc.predict_adjs_sparse(x, adj, fus, [pairs_sup])
                aux_dis = enc.predict_adjs_sparse(x, adj, fus, m_dis)
                dev_rowsum = 0.0
                for l in range(2):
                    ee = torch.cat(edge_e[l], 1)                         # [E, C]
                    al = torch.cat([utils.sp_softmax(idx, torch.sigmoid(edge_e[l][c]), n) for c in range(C)], 1)
                    rs = torch.zeros(n, C).index_add_(0, idx[0], al)
                    dev_rowsum = max(dev_rowsum, float((rs - 1).abs().max()))
                    out[k + "e%d" % l] = ee[sel_e].numpy()
                    out[k + "alpha%d" % l] = al[sel_e].numpy()
                    out[k + "e%d_absmax" % l] = np.float32(ee.abs().max())
                    out[k + "out%d" % l] = torch.cat([em[l][c][sel_n, -D:] for c in range(C)], 1).numpy()
                    out[k + "out%d_absmax" % l] = np.float32(max(float(em[l][c][:, -D:].abs().max()) for c in range(C)))
                out[k + "alpha_rowsum_dev"] = np.float64(dev_rowsum)
                out[k + "feat2"] = feats[1][sel_n].numpy()
                out[k + "feat2_absmax"] = np.float32(feats[1].abs().max())
                # SupEdge (pretrainer.py:726-747), DisEdge (596-627), DifHead (819-832), constrain_layer 0
                l_sup = sum(utils.adj_mse_loss(torch.sigmoid(torch.sum(torch.stack([h[0] for h in aux_sup[l]]), 0)).squeeze(),
                                               y_sup) for l in range(2))
                l_dis = 0.0
                for l in range(2):
                    ph = torch.sigmoid(torch.sum(torch.stack([h[0] for h in aux_dis[l]][: int(C / 2)]), 0)).squeeze()
                    pt = torch.sigmoid(torch.sum(torch.stack([h[1] for h in aux_dis[l]][int(C / 2):]), 0)).squeeze()
                    l_dis = l_dis + utils.adj_mse_loss(ph, y_dis[0]) + utils.adj_mse_loss(pt, y_dis[1])
                l_dif = 0.0
                for l in range(2):
                    for c in range(C):
                        lab = torch.full((n,), c, dtype=torch.long)
                        l_dif = l_dif + torch.nn.functional.nll_loss(clf[l](em[l][c], cls=True), lab)
                out[k + "loss_sup"] = np.float64(float(l_sup))
                out[k + "loss_dis"] = np.float64(float(l_dis))
                out[k + "loss_dif"] = np.float64(float(l_dif))
            print(k, "e absmax", float(out[k + "e0_absmax"]), "rowsum dev", dev_rowsum, "losses", float(l_sup), float(l_dis),
                  float(l_dif), flush=True)
    path = os.path.join(HERE, "bundled_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
