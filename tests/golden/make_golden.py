"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference is imported from /root/reference with three stub modules (ipdb, tensorboardX,
matplotlib: tests/golden/_stubs) because those are not installed.  Everything the reference
computes is produced by its own functions/classes: `data_load.load_data`, `layers.DisGALayer`,
`models.DISGAT`, `pretrainer.{SupEdge,GeneratedEdge,DifHead}Trainer`, `trainer.ClsTrainer`,
`utils.{sp_softmax,sp_matmul,adj_mse_loss}`.  Fixtures total about 10 MB.
"""
import hashlib
import os
import random
import sys
import tempfile

import numpy as np
import scipy.sparse as sp
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
import trainer  # noqa: E402
import utils  # noqa: E402

torch.set_num_threads(8)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_args(extra=()):
    args = utils.get_parser().parse_args(list(extra))
    args.cuda = False
    args.hetero = False
    return args


def np_(t):
    return t.detach().cpu().numpy().copy()  # copy: params are updated in place later


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", name, os.path.getsize(path) // 1024, "KiB")


def write_dataset(tmp, n, feat, labels, edges=None, csr=None):
    os.makedirs(tmp, exist_ok=True)
    np.save(os.path.join(tmp, "label.npy"), labels)
    np.save(os.path.join(tmp, "feature_new.npy"), feat)
    if edges is not None:
        np.save(os.path.join(tmp, "adj_1.npy"), edges)
    else:
        sp.save_npz(os.path.join(tmp, "adj_1_sp.npz"), csr)


def random_edges(rng, n, m, hub=True):
    src = rng.randint(0, n, size=m)
    dst = rng.randint(0, n, size=m)
    if hub:  # a hub row, duplicates and explicit self loops
        src[: m // 6] = 3
        src[m // 6: m // 6 + 4] = dst[m // 6: m // 6 + 4]
        src[-5:] = src[:5]
        dst[-5:] = dst[:5]
    e = np.stack([src, dst], 1).astype(np.int64)
    e[0] = [n - 1, n - 2]  # make sure max id + 1 == n (edge2adj sizes by max id)
    return e


# ------------------------------------------------------------------ graph construction
def golden_graph(tmpdir):
    rng = np.random.RandomState(11)
    out = {}
    # (i) edge list with duplicates / self loops / asymmetric entries  (data_load.py:40-49)
    n = 37
    edges = random_edges(rng, n, 150)
    feat = np.abs(rng.randn(n, 6)) + 0.1
    lab = rng.randint(0, 3, size=n)
    d = os.path.join(tmpdir, "g_edges/")
    write_dataset(d, n, feat, lab, edges=edges)
    args = ref_args(["--sparse"])
    adj, features, labels = data_load.load_data(args, path=d, dataset="g_edges", edge_type=1)
    adj = adj.coalesce()
    out.update(el_n=n, el_edges=edges, el_feat=feat, el_indices=np_(adj.indices()),
               el_values=np_(adj.values()), el_features=np_(features))
    # (ii) weighted asymmetric scipy CSR  (data_load.py:43-45)
    n2 = 29
    dense = np.zeros((n2, n2), dtype=np.float32)
    for _ in range(90):
        i, j = rng.randint(0, n2, 2)
        dense[i, j] = rng.choice([0.5, 1.0, 2.0, 3.0])
    csr = sp.csr_matrix(dense)
    d = os.path.join(tmpdir, "g_csr/")
    write_dataset(d, n2, np.abs(rng.randn(n2, 4)) + 0.1, rng.randint(0, 2, n2), csr=csr)
    adj2, _, _ = data_load.load_data(args, path=d, dataset="g_csr", edge_type=1)
    adj2 = adj2.coalesce()
    coo = csr.tocoo()
    out.update(csr_n=n2, csr_row=coo.row.astype(np.int64), csr_col=coo.col.astype(np.int64),
               csr_val=coo.data.astype(np.float64), csr_indices=np_(adj2.indices()),
               csr_values=np_(adj2.values()))
    save("graph_small", **out)


def golden_bundled(tmpdir):
    """Hashes of the processed adjacency of the three bundled graphs (inputs are in data/)."""
    out = {}
    for ds in ("cora", "chameleon", "cora_full"):
        src = os.path.join(REF, "data", ds)
        d = os.path.join(tmpdir, "b_" + ds + "/")
        os.makedirs(d)
        for f in os.listdir(src):
            if f.startswith("adj_1") or f == "label.npy":
                os.symlink(os.path.join(src, f), os.path.join(d, f))
        lab = np.load(os.path.join(src, "label.npy"))
        if os.path.exists(os.path.join(src, "feature_new.npy")):
            os.symlink(os.path.join(src, "feature_new.npy"), os.path.join(d, "feature_new.npy"))
        else:  # deterministic synthetic features for cora / cora_full
            r = np.random.RandomState(0)
            mu = r.randn(int(lab.max()) + 1, 64)
            np.save(os.path.join(d, "feature_new.npy"), np.abs(mu[lab] + r.randn(lab.shape[0], 64)))
        args = ref_args(["--sparse"])
        adj, features, labels = data_load.load_data(args, path=d, dataset=ds, edge_type=1)
        adj = adj.coalesce()
        idx, val = np_(adj.indices()), np_(adj.values())
        out[ds + "_n"] = adj.shape[0]
        out[ds + "_e"] = idx.shape[1]
        out[ds + "_idx_sha"] = sha(idx)
        out[ds + "_val_sha"] = sha(val)
        out[ds + "_feat_sha"] = sha(np_(features))
        out[ds + "_feat_row0"] = np_(features)[0]
        if ds == "cora":
            # sampler on a real graph: SupEdge.sample_train (pretrainer.py:683-707)
            a = ref_args(["--sparse", "--model=DISGAT"])
            a.size = features.shape[1]
            enc = models.DISGAT(a, nfeat=a.size, nhid=a.nhid, nclass=a.nhid, nheads=a.nhead, dropout=0.1)
            st = pretrainer.SupEdgeTrainer(a, enc, 1.0)
            gt = st.get_label_all(features, adj)
            torch.manual_seed(4)
            np.random.seed(4)
            lab_used, ind = st.sample_train(gt)
            out["cora_sample_m"] = ind[0].shape[1]
            out["cora_sample_idx_sha"] = sha(np_(ind[0]))
            out["cora_sample_lab_sha"] = sha(np_(lab_used))
    save("graph_bundled", **out)


# ------------------------------------------------------------------ single layer
def golden_layers(tmpdir):
    rng = np.random.RandomState(5)
    for tag, (n, f, dd, m) in {"s8": (48, 12, 8, 260), "s64": (64, 20, 64, 420)}.items():
        edges = random_edges(rng, n, m)
        feat = rng.randn(n, f)
        d = os.path.join(tmpdir, "l_%s/" % tag)
        write_dataset(d, n, feat, rng.randint(0, 2, n), edges=edges)
        args = ref_args(["--sparse", "--origin_feat"])
        np.save(os.path.join(d, "feature.npy"), feat)
        adj, features, _ = data_load.load_data(args, path=d, dataset=tag, edge_type=1)
        idx = adj.coalesce().indices()
        e = idx.shape[1]
        aux = [torch.from_numpy(rng.randint(0, n, size=(2, 70))).long(),
               torch.from_numpy(rng.randint(0, n, size=(2, 33))).long()]
        r_out = torch.from_numpy(rng.randn(n, dd)).float()
        r_e = torch.from_numpy(rng.randn(e, 1)).float()
        r_aux = [torch.from_numpy(rng.randn(70, 1)).float(), torch.from_numpy(rng.randn(33, 1)).float()]
        out = dict(n=n, x=np_(features), indices=np_(idx), aux0=np_(aux[0]), aux1=np_(aux[1]),
                   r_out=np_(r_out), r_e=np_(r_e), r_aux0=np_(r_aux[0]), r_aux1=np_(r_aux[1]))
        for att in (1, 2, 3):
            for gnn in ("AT", "SAGE", "GCN"):
                torch.manual_seed(100 * att + len(gnn))
                lay = layers.DisGALayer(f, dd, dropout=0.3, alpha=0.1, concat=True, att_type=att, gnn_type=gnn)
                lay.eval()
                x = features.clone().requires_grad_(True)
                o, ee, au = lay(x, adj, aux)
                loss = (o * r_out).sum() + (ee * r_e).sum() + (au[0] * r_aux[0]).sum() + (au[1] * r_aux[1]).sum()
                loss.backward()
                k = "a%d_%s_" % (att, gnn)
                out[k + "out"] = np_(o)
                out[k + "edge_e"] = np_(ee)
                out[k + "aux0"] = np_(au[0])
                out[k + "aux1"] = np_(au[1])
                out[k + "gx"] = np_(x.grad)
                for name, prm in lay.named_parameters():
                    out[k + "p." + name] = np_(prm)
                    if prm.grad is not None:
                        out[k + "g." + name] = np_(prm.grad)
                # alpha itself (utils.sp_softmax on sigmoid(e)), and the plain sp_matmul
                alpha = utils.sp_softmax(idx, torch.sigmoid(ee.detach()), n)
                out[k + "alpha"] = np_(alpha)
        mat = torch.from_numpy(rng.randn(n, 5)).float()
        vals = torch.from_numpy(rng.rand(e, 1)).float()
        out["spmm_mat"], out["spmm_vals"] = np_(mat), np_(vals)
        out["spmm_out"] = np_(utils.sp_matmul(idx, vals, mat))
        out["spsm_out"] = np_(utils.sp_softmax(idx, vals, n))
        save("layer_" + tag, **out)


# ------------------------------------------------------------------ model + trainers
def seed_all(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def golden_model(tmpdir):
    rng = np.random.RandomState(21)
    n, f, m = 72, 20, 420
    edges = random_edges(rng, n, m)
    feat = np.abs(rng.randn(n, f)) + 0.05
    labels_np = np.arange(n) % 3
    rng.shuffle(labels_np)
    d = os.path.join(tmpdir, "m/")
    write_dataset(d, n, feat, labels_np, edges=edges)
    cwd = os.getcwd()
    os.chdir(tmpdir)
    try:
        for att, gnn, extra in ((3, "AT", []), (1, "SAGE", []), (2, "GCN", []),
                                (3, "AT", ["--residue", "--residue_type=2", "--constrain_layer=1"])):
            tag = "model_a%d_%s%s" % (att, gnn, "_res" if extra else "")
            argv = ["--sparse", "--model=DISGAT", "--dataset=gold_" + tag, "--att=%d" % att,
                    "--gnn_type=" + gnn, "--nhead=4", "--nhid=64", "--dropout=0.0",
                    "--pretrain", "SupEdge", "DisEdge", "DifHead", "--pre_weight", "1", "1", "1",
                    "--pre_edge", "1", "1", "1", "--downstream", "CLS", "--down_weight", "1.0",
                    "--finetune"] + extra
            args = ref_args(argv)
            args.hetero = True  # main.py:29-30
            args.edge_num = 1
            adjs, features, labels = data_load.load_data(args, path=d, dataset="m", edge_type=1)
            args.size = features.shape[1]
            args.nclass = labels.max().item() + 1
            adj = adjs[0]
            seed_all(4)
            enc = models.DISGAT(args, nfeat=args.size, nhid=args.nhid, nclass=args.nhid,
                                nheads=args.nhead, dropout=args.dropout)
            out = dict(n=n, x=np_(features), indices=np_(adj.coalesce().indices()), labels=np_(labels),
                       argv=np.array(argv))
            for k, v in enc.state_dict().items():
                out["enc0." + k] = np_(v)
            # trainers in main.py order: SSL trainers first (237-251), then CLS (253-258)
            sup = pretrainer.SupEdgeTrainer(args, enc, 1.0)
            sup_lab = sup.get_label_all(features, adj)
            dis = pretrainer.GeneratedEdgeTrainer(args, enc, 1.0)
            dis_lab = dis.get_label_all(features, adj, labels)
            dif = pretrainer.DifHeadTrainer(args, enc, 1.0)
            cls = trainer.ClsTrainer(args, enc, labels, 1.0)
            out["cls_idx_train"] = np_(cls.idx_train)
            out["cls_idx_val"] = np_(cls.idx_val)
            out["cls_idx_test"] = np_(cls.idx_test)
            for nm, tr in (("sup", sup), ("dis", dis), ("dif", dif), ("cls", cls)):
                for k, v in tr.fuse1.state_dict().items():
                    out["%s0.fuse1.%s" % (nm, k)] = np_(v)
                for k, v in tr.fuse2.state_dict().items():
                    out["%s0.fuse2.%s" % (nm, k)] = np_(v)
            for k, v in dif.classifier1.state_dict().items():
                out["dif0.classifier1." + k] = np_(v)
            for k, v in dif.classifier2.state_dict().items():
                out["dif0.classifier2." + k] = np_(v)
            for k, v in cls.classifier.state_dict().items():
                out["cls0.classifier." + k] = np_(v)

            # eval-mode traversal outputs with the CLS trainer's fusers (models.py:181-373)
            enc.eval()
            fusers = [cls.fuse1, cls.fuse2]
            aux = [torch.from_numpy(rng.randint(0, n, size=(2, 90))).long(),
                   torch.from_numpy(rng.randint(0, n, size=(2, 41))).long()]
            out["aux0"], out["aux1"] = np_(aux[0]), np_(aux[1])
            with torch.no_grad():
                em = enc.get_em(features, adj, fusers)
                out["get_em_1"], out["get_em_2"] = np_(em[0]), np_(em[1])
                out["forward"] = np_(enc(features, adj, fusers))
                ga = enc.get_adjs(features, adj, fusers)
                out["get_adjs"] = np.stack([np.stack([np_(h) for h in lay]) for lay in ga])
                pa = enc.predict_adjs_sparse(features, adj, fusers, aux)
                out["pred_aux0"] = np.stack([np.stack([np_(h[0]) for h in lay]) for lay in pa])
                out["pred_aux1"] = np.stack([np.stack([np_(h[1]) for h in lay]) for lay in pa])
                ge = enc.get_edge_em(features, adj, fusers)
                out["edge_em_l0"] = np.stack([np_(h) for h in ge[0]])
                out["edge_em_l1"] = np.stack([np_(h) for h in ge[1]])

            # one train_step of each trainer (dropout = 0 so train mode is deterministic);
            # record loss, encoder grads and encoder params after the Adam step.
            # (fixture size: full gradients only for the first case; channels 0 and 3 otherwise.)
            full = (att, gnn, bool(extra)) == (3, "AT", False)

            def record(nm, tr, log):
                for k, v in log.items():
                    out["%s.log.%s" % (nm, k)] = np.float64(v)
                for k, prm in enc.named_parameters():
                    if prm.grad is not None and (full or "_0." in k or "_3." in k):
                        out["%s.encgrad.%s" % (nm, k)] = np_(prm.grad)
                for k, prm in tr.fuse1.named_parameters():
                    if prm.grad is not None:
                        out["%s.fuse1grad.%s" % (nm, k)] = np_(prm.grad)

            seed_all(7)
            record("cls", cls, cls.train_step([features, adj], labels, 0))
            seed_all(8)
            record("sup", sup, sup.train_step([features, adj], sup_lab))
            seed_all(8)  # same seed: lets the test replay the sampler in isolation
            lab_used, ind = sup.sample_train(sup_lab)
            out["sup.sample_idx"], out["sup.sample_lab"] = np_(ind[0]), np_(lab_used)
            seed_all(9)
            record("dis", dis, dis.train_step([features, adj], dis_lab))
            seed_all(9)
            labs, inds = dis.sample_train()
            out["dis.sample_idx0"], out["dis.sample_lab0"] = np_(inds[0]), np_(labs[0])
            out["dis.sample_idx1"], out["dis.sample_lab1"] = np_(inds[1]), np_(labs[1])
            seed_all(10)
            record("dif", dif, dif.train_step([features, adj], None))
            for k, v in dif.classifier1.named_parameters():
                out["dif.cls1grad." + k] = np_(v.grad)
            # encoder after the four sequential Adam steps (one Adam state per trainer)
            for k, v in enc.state_dict().items():
                out["enc_final." + k] = np_(v)
            save(tag, **out)
    finally:
        os.chdir(cwd)


def golden_loss():
    rng = np.random.RandomState(3)
    pred = torch.from_numpy(rng.rand(500)).float()
    tgt = torch.from_numpy((rng.rand(500) < 0.15).astype(np.float32))
    save("loss_small", pred=np_(pred), tgt=np_(tgt), loss=np_(utils.adj_mse_loss(pred, tgt)))


if __name__ == "__main__":
    which = sys.argv[1:] or ["graph", "layers", "model", "loss", "bundled"]
    with tempfile.TemporaryDirectory() as tmp:
        if "graph" in which:
            golden_graph(tmp)
        if "layers" in which:
            golden_layers(tmp)
        if "loss" in which:
            golden_loss()
        if "model" in which:
            golden_model(tmp)
        if "bundled" in which:
            golden_bundled(tmp)
