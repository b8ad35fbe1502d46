"""GPU parity of the trainers' `train_step`s against the reference's recorded steps.

tests/golden/model_*.npz hold, for four DISGAT configurations, one CLS, SupEdge, DisEdge and
DifHead `train_step` run by the UNMODIFIED reference (dropout 0 so train mode is deterministic;
RNG seeded right before each step): logged losses, encoder gradients after each step and the
encoder after the four Adam updates.  This replays them on the B200 path -- sampler (CPU RNG
order), fused pair scoring / loss kernels, DifHead tail, per-trainer Adam states.
"""
import random

import numpy as np
import pytest
import torch

import edgedisentangle_ssl_b200 as edis
from edgedisentangle_ssl_b200 import trainer as T
from edgedisentangle_ssl_b200.utils import get_parser
from helpers import load, t, assert_close, params_from, group_floor

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MODEL_TAGS = ["model_a3_AT", "model_a1_SAGE", "model_a2_GCN", "model_a3_AT_res"]


def seed_all(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def load_into(module, g, prefix):
    module.load_state_dict({k: v.to(DEV) for k, v in params_from(g, prefix).items()}, strict=True)


@pytest.mark.parametrize("tag", MODEL_TAGS)
def test_train_steps_vs_reference_golden(tag):
    g = load(tag)
    args = get_parser().parse_args([str(a) for a in g["argv"]])
    args.cuda, args.hetero, args.edge_num = True, True, 1
    args.size = g["x"].shape[1]
    x, labels = t(g["x"]).to(DEV), t(g["labels"]).to(DEV)
    args.nclass = int(labels.max()) + 1
    n = int(g["n"])
    idx = torch.as_tensor(g["indices"])
    adj = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (n, n)).to(DEV)

    seed_all(4)
    enc = edis.DISGAT(args, nfeat=args.size, nhid=args.nhid, nclass=args.nhid, nheads=args.nhead,
                      dropout=args.dropout).to(DEV)
    sup = T.SupEdgeTrainer(args, enc, 1.0)
    sup_lab = sup.get_label_all(x, adj)
    dis = T.GeneratedEdgeTrainer(args, enc, 1.0)
    dis_lab = dis.get_label_all(x, adj, labels)
    dif = T.DifHeadTrainer(args, enc, 1.0)
    cls = T.ClsTrainer(args, enc, labels, 1.0)
    # same initial state as the recorded run (constructors above only fixed shapes / optimisers)
    load_into(enc, g, "enc0.")
    for nm, tr in (("sup", sup), ("dis", dis), ("dif", dif), ("cls", cls)):
        load_into(tr.fuse1, g, nm + "0.fuse1.")
        load_into(tr.fuse2, g, nm + "0.fuse2.")
    load_into(dif.classifier1, g, "dif0.classifier1.")
    load_into(dif.classifier2, g, "dif0.classifier2.")
    load_into(cls.classifier, g, "cls0.classifier.")
    cls.idx_train, cls.idx_val, cls.idx_test = (t(g["cls_idx_" + k]).to(DEV) for k in ("train", "val", "test"))
    # the split itself replays utils.split's python-RNG order
    assert int(g["cls_idx_train"].shape[0]) == int(cls.idx_train.shape[0])

    from oracle import disgat as od
    from helpers import rel_err, kink_sensitivity

    def snapshot(tr):
        """float64 copies of the weights a step is about to use (encoder, the trainer's fusers / heads)."""
        c64 = lambda sd: {k: v.detach().cpu().double() for k, v in sd.items()}
        snap = {"enc": c64(enc.state_dict()), "fus": [c64(tr.fuse1.state_dict()), c64(tr.fuse2.state_dict())]}
        for nm_ in ("classifier", "classifier1", "classifier2"):
            if hasattr(tr, nm_):
                snap[nm_] = c64(getattr(tr, nm_).state_dict())
        return snap

    def arbiter(nm, snap):
        """Encoder gradient of step `nm` evaluated in FLOAT64 by the CPU oracle at the snapshot weights."""
        p = {k: v.clone().requires_grad_(True) for k, v in snap["enc"].items() if k.startswith("attention")}
        kw = dict(residue=bool(args.residue), residue_type=args.residue_type, no_relu=bool(args.fuse_no_relu))
        xi, ii = x.cpu().double(), idx
        if nm == "cls":
            r = od.disgat_traverse(p, snap["fus"], xi, ii, args.nhead, args.att, args.gnn_type, **kw)
            out = od.mlp(snap["classifier"], r["feats"][-1], cls=True)
            it = cls.idx_train.cpu()
            loss = torch.nn.functional.nll_loss(out[it], labels.cpu()[it])
        elif nm == "sup":
            r = od.disgat_traverse(p, snap["fus"], xi, ii, args.nhead, args.att, args.gnn_type,
                                   aux=[torch.as_tensor(g["sup.sample_idx"])], **kw)
            loss = od.supedge_loss(r["aux"], torch.as_tensor(g["sup.sample_lab"]).double(), args.constrain_layer)
        elif nm == "dis":
            aux = [torch.as_tensor(g["dis.sample_idx%d" % k]) for k in range(2)]
            r = od.disgat_traverse(p, snap["fus"], xi, ii, args.nhead, args.att, args.gnn_type, aux=aux, **kw)
            loss = od.disedge_loss(r["aux"], [torch.as_tensor(g["dis.sample_lab%d" % k]).double() for k in range(2)],
                                   args.constrain_layer)
        else:
            r = od.disgat_traverse(p, snap["fus"], xi, ii, args.nhead, args.att, args.gnn_type, **kw)
            loss = od.difhead_loss(r["edge_em"], [snap["classifier1"], snap["classifier2"]])
        loss.backward()
        return {k: v.grad for k, v in p.items() if v.grad is not None}

    def check(nm, log, snap, rt_loss=1e-5, rt_grad=1e-4):
        for k, v in log.items():
            key = "%s.log.%s" % (nm, k)
            if key in g and k.startswith("loss"):
                assert_close(v, g[key], rt_loss, key)
            elif key in g and k.startswith("acc"):
                # accuracies are counts over a few dozen nodes: allow one borderline prediction
                assert abs(v - float(g[key])) <= 1.0 / 12 + 1e-9, (key, v, float(g[key]))
            # roc_val / macroF_val are rank statistics over ~12 validation nodes whose logits are
            # nearly tied at initialisation: not a numerical parity quantity, only checked for presence
            elif key in g:
                assert np.isfinite(v), key
        refs = {k[len(nm) + 9:]: v for k, v in g.items() if k.startswith(nm + ".encgrad.")}
        floor = group_floor(refs.values())
        seen = 0
        for name, prm in enc.named_parameters():
            if name in refs:
                assert prm.grad is not None, name
                assert_close(prm.grad.cpu(), refs[name], rt_grad, "%s grad %s" % (nm, name), floor)
                seen += 1
        assert seen == len(refs) and seen > 0
        # float64 arbiter AT THE WEIGHTS THIS STEP USED: the recorded reference gradients (above, 1e-4) belong
        # to the reference's own weights, which after the first Adam update differ from ours by ~1e-4 of the
        # weight scale (a gradient entry at rounding-noise level can flip the sign of its first Adam step).
        # Against the true gradient at OUR weights the bound is 2e-5, or 16x the error the reference's fp32
        # arithmetic itself makes against float64 on the first step, where both sides share the weights.
        f64 = arbiter(nm, snap)
        floor64 = group_floor(f64.values())
        sens = kink_sensitivity(lambda: arbiter(nm, snap), f64, floor64)     # helpers.rounding_noise: why
        for name, prm in enc.named_parameters():
            if name in f64:
                # 1e-4, not 2e-5: the kernels take the softmax backward's row term t_i = sum_k alpha_ik dalpha_ik
                # from the stored aggregate (<gh_i, agg_i>, single pass) -- on the near-constant layer-2
                # messages of these toy graphs that form carries ~6e-5 on the tiny `a` gradients (1e-3 of the
                # step's largest) in PLAIN fp32 torch as well, against 5e-6 for the reference's two-reduction
                # form (tests/test_oracle_golden.py::test_single_pass_softmax_backward_conditioning)
                tol = max(1e-4 if name.endswith(".a") else 2e-5, 2.0 * sens[name])
                if nm == "cls" and name in refs:
                    tol = max(tol, 16.0 * rel_err(refs[name], f64[name], floor64))
                err = rel_err(prm.grad.cpu(), f64[name], floor64)
                assert err <= tol, "%s grad %s vs float64: %.3e > %.3e" % (nm, name, err, tol)
        return log

    seed_all(7)
    snap = snapshot(cls)
    check("cls", cls.train_step([x, adj], labels, 0), snap)
    seed_all(8)
    snap = snapshot(sup)
    check("sup", sup.train_step([x, adj], sup_lab), snap)
    seed_all(9)
    snap = snapshot(dis)
    check("dis", dis.train_step([x, adj], dis_lab), snap)
    seed_all(10)
    snap = snapshot(dif)
    check("dif", dif.train_step([x, adj], None), snap)
    for name, prm in dif.classifier1.named_parameters():
        assert_close(prm.grad.cpu(), g["dif.cls1grad." + name], 5e-5, "dif classifier1 " + name)
    # encoder after four Adam updates (one Adam state per trainer, like the reference).  Adam's first
    # steps move every weight by ~lr * g / |g|: a gradient entry at rounding-noise level can flip the
    # direction of its update, so the bound is a fraction of 4 * lr relative to the weight scale, not
    # the 1e-5 of the gradients themselves (checked above)
    final = enc.state_dict()
    for k, v in params_from(g, "enc_final.").items():
        assert_close(final[k].cpu(), v, 4e-4, "enc_final." + k)


def test_sample_train_matches_reference_sets():
    """Trainer-level sampler: pairs and labels bit-exact vs the reference's sample_train."""
    g = load("model_a3_AT")
    args = get_parser().parse_args([str(a) for a in g["argv"]])
    args.cuda, args.hetero, args.size = True, True, g["x"].shape[1]
    n = int(g["n"])
    idx = torch.as_tensor(g["indices"])
    adj = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (n, n)).to(DEV)
    enc = edis.DISGAT(args, nfeat=args.size, nhid=args.nhid, nclass=args.nhid, nheads=args.nhead, dropout=0.0).to(DEV)
    sup = T.SupEdgeTrainer(args, enc, 1.0)
    lab = sup.get_label_all(None, adj)
    seed_all(8)
    y, masks = sup.sample_train(lab)
    assert np.array_equal(masks[0].cpu().numpy(), g["sup.sample_idx"])
    assert np.array_equal(y.cpu().numpy(), g["sup.sample_lab"])
    dis = T.GeneratedEdgeTrainer(args, enc, 1.0)
    dis.get_label_all(None, adj, t(g["labels"]))
    seed_all(9)
    ys, ms = dis.sample_train()
    for k in range(2):
        assert np.array_equal(ms[k].cpu().numpy(), g["dis.sample_idx%d" % k])
        assert np.array_equal(ys[k].cpu().numpy(), g["dis.sample_lab%d" % k])


def test_cli_epoch_on_cora():
    """The CLI loop runs end to end on bundled cora with the example script's flags
    (Example_cora_full.sh:38, 2 epochs) and trains (loss decreases, finite)."""
    import os
    from edgedisentangle_ssl_b200.main import run
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")
    hist = run(["--seed=4", "--model=DISGAT", "--used_edge=1", "--finetune", "--downstream=CLS", "--down_weight=1.0",
                "--steps=5", "--nhead=4", "--dataset=cora", "--pretrain", "SupEdge", "DisEdge", "DifHead",
                "--pre_weight", "1", "1", "1", "--pre_edge", "1", "1", "1", "--sparse", "--att=3",
                "--constrain_layer=0", "--epochs=3", "--gnn_type=AT"], data_root=root)
    assert len(hist) == 3
    for h in hist:
        for k in ("loss_train", "loss_heads_sup", "loss_head_disen", "loss_head_diversity"):
            assert np.isfinite(h[k]), (k, h[k])
    assert hist[-1]["loss_train"] < hist[0]["loss_train"]


def test_device_sampler_has_the_reference_distribution():
    """O(M) device sampler vs the bit-exact one: same pair count (Binomial mean), same share of
    positives, forced third of the positives present, sorted unique keys, exact labels."""
    from edgedisentangle_ssl_b200.sampler import sample_pairs, sample_pairs_device
    from oracle import graph as og
    rng = np.random.RandomState(0)
    n = 3000
    idx, _ = og.build_adjacency(n, rng.randint(0, n, 20000), rng.randint(0, n, 20000))
    seed_all(1)
    exact = [sample_pairs(n, idx)[0].shape[1] for _ in range(3)]
    pos_key = torch.from_numpy(idx[0] * n + idx[1]).to(DEV)
    counts, pos_frac = [], []
    for s in range(6):
        gen = torch.Generator(device=DEV).manual_seed(s)
        pairs, lab = sample_pairs_device(n, pos_key, gen)
        key = pairs[0] * n + pairs[1]
        assert torch.all(key[1:] > key[:-1])
        assert torch.equal(lab.bool(), torch.isin(key, pos_key))
        counts.append(pairs.shape[1])
        pos_frac.append(lab.mean().item())
        assert int(lab.sum()) >= idx.shape[1] // 3
    e = idx.shape[1]
    expect = 3 * e + e // 3          # Bernoulli hits + forced third (overlap is O(rho) small)
    assert abs(np.mean(counts) - expect) < 0.02 * expect and abs(np.mean(exact) - expect) < 0.02 * expect
    assert abs(np.mean(pos_frac) - (e // 3 + 3 * e * e / n / n) / expect) < 0.01


# (key in tests/golden/accuracy_ref.json, dataset, gnn_type, index of the compared checkpoint, sampler)
ACCURACY_CASES = [
    ("chameleon", "chameleon", "AT", -1, None),          # real features, 81 epochs: final accuracy
    ("chameleon_SAGE", "chameleon", "SAGE", 1, None),    # gnn_type SAGE (BASELINE config[2]), epoch 40
    ("cora", "cora", "AT", 1, None),                     # synthetic features; epoch 40 (see docstring)
    ("cora_full", "cora_full", "AT", 1, "device"),       # epoch 40; O(M) device sampler keeps the suite short
]


@pytest.mark.parametrize("key,ds,gnn,at,sampler", ACCURACY_CASES)
def test_accuracy_parity_within_seed_noise(key, ds, gnn, at, sampler, monkeypatch):
    """North-star: node-classification accuracy 'within seed noise' of the reference on cora / cora_full /
    chameleon, gnn_type AT and SAGE, >= 5 seeds where the reference curves exist.

    tests/golden/accuracy_ref.json holds the UNMODIFIED reference CLI's test accuracy every 40 epochs
    (tests/golden/run_reference_accuracy.py: example flags of Example_cora_full.sh:38, CPU, seeds 4-8).
    The same flags run here on the B200 path.  Train-mode dropout and Adam make single runs differ by a
    few points on both sides, so the check is on the MEAN over the seeds with the reference's own
    seed-to-seed spread as the band, and every run must be far above chance.  cora / cora_full use the
    label-derived synthetic features of SURVEY 8(d) (their blobs are missing), on which the reference
    itself is unstable after ~epoch 80 (4 of 5 cora seeds collapse to the majority class 0.3024): the
    comparison is made at epoch 40, where both sides are still in the regime the hot path decides.  On
    cora_full (70 classes) the reference is at 0.047 = the majority-class share for every seed at epoch 40
    (49 s per CPU epoch bounded how far the reference curves could be recorded); the B200 path lands on the
    same 0.047."""
    import contextlib
    import io
    import json
    import os
    from edgedisentangle_ssl_b200.main import run
    here = os.path.dirname(os.path.abspath(__file__))
    table = json.load(open(os.path.join(here, "golden", "accuracy_ref.json")))
    if key not in table:
        pytest.skip("no reference curve recorded for %s" % key)
    ref = table[key]
    if sampler:
        monkeypatch.setenv("EDIS_SAMPLER", sampler)
    root = os.path.join(os.path.dirname(here), "data")
    ours, theirs = [], []
    for name in sorted(ref):
        seed, r = int(name[4:]), ref[name]
        n_ckpt = len(r["test_acc_every_40"])
        upto = n_ckpt - 1 if at < 0 else at
        epochs = 40 * upto + 1
        with contextlib.redirect_stdout(io.StringIO()):
            hist = run(["--seed=%d" % seed, "--model=DISGAT", "--used_edge=1", "--finetune", "--downstream=CLS",
                        "--down_weight=1.0", "--steps=5", "--nhead=4", "--dataset=" + ds, "--pretrain", "SupEdge",
                        "DisEdge", "DifHead", "--pre_weight", "1", "1", "1", "--pre_edge", "1", "1", "1", "--sparse",
                        "--att=3", "--constrain_layer=0", "--epochs=%d" % epochs, "--gnn_type=" + gnn], data_root=root)
        accs = [h["acc_test"] for h in hist if "acc_test" in h]
        assert len(accs) == upto + 1
        ours.append(accs[-1])
        theirs.append(r["test_acc_every_40"][upto])
    spread = max(theirs) - min(theirs)
    n_class = {"chameleon": 5, "cora": 7, "cora_full": 70}[ds]
    print("%s test accuracy at epoch %d over %d seeds: ours %s, reference %s" % (key, epochs - 1, len(ours), ours, theirs))
    if len(ours) < 3:
        pytest.skip("only %d reference seeds recorded for %s so far" % (len(ours), key))
    # both sides must have learnt (mean far above chance 1 / n_class) and the means must agree within the
    # reference's own seed-to-seed spread (floor 0.05)
    assert np.mean(ours) > 2.0 / n_class and np.mean(theirs) > 2.0 / n_class
    assert abs(np.mean(ours) - np.mean(theirs)) <= max(spread, 0.05), (ours, theirs)


def test_checkpoint_round_trip_resumes_bit_exactly(tmp_path):
    """save_model / load_model (main.py:214-235 + per-trainer fusers, classifiers and Adam states):
    2 epochs + save, then a fresh process state loading it must hold identical weights, keep the
    reference's 'encoder' key layout, and continue training."""
    import contextlib
    import io
    import os
    from edgedisentangle_ssl_b200 import main as M
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")
    flags = ["--seed=4", "--model=DISGAT", "--used_edge=1", "--finetune", "--downstream=CLS", "--down_weight=1.0",
             "--steps=2", "--nhead=4", "--dataset=cora", "--pretrain", "SupEdge", "DifHead", "--pre_weight", "1", "1",
             "--pre_edge", "1", "1", "--sparse", "--att=3", "--constrain_layer=0", "--gnn_type=AT"]
    with contextlib.redirect_stdout(io.StringIO()):
        M.run(flags + ["--epochs=2"], data_root=root, ckpt_root=str(tmp_path), save_every=2)
    args = M.utils.get_parser().parse_args(flags + ["--epochs=2", "--load=1"])
    _, path = M.checkpoint_path(args, 1, str(tmp_path))
    saved = torch.load(path, map_location="cpu")
    assert set(saved) >= {"encoder", "trainers"} and len(saved["trainers"]) == 3
    assert "attention1_0.W" in saved["encoder"] and "attention2_3.W_em" in saved["encoder"]
    assert saved["trainers"][2]["class"] == "ClsTrainer" and len(saved["trainers"][2]["optimizers"]) == 4
    with contextlib.redirect_stdout(io.StringIO()):
        hist = M.run(flags + ["--epochs=1", "--load=1"], data_root=root, ckpt_root=str(tmp_path), save_every=1)
    assert np.isfinite(hist[0]["loss_train"])
    # the resumed run started from the saved weights: its first test (epoch 0, before any step) sees them
    resumed = torch.load(M.checkpoint_path(args, 0, str(tmp_path))[1], map_location="cpu")
    moved = max(float((resumed["encoder"][k] - v).abs().max()) for k, v in saved["encoder"].items())
    assert 0.0 < moved < 0.1          # one more epoch of Adam steps from the loaded state, not a re-init


def test_case_study_matches_reference_golden():
    """`Trainer.analyze_disentangle` (--case, trainer.py:82-134): same sampled pairs (RNG stream
    replayed), same channel / feature correlation maps as the unmodified reference
    (tests/golden/make_golden_case.py), eval mode."""
    g, c = load("model_a3_AT"), load("case_a3_AT")
    args = get_parser().parse_args([str(a) for a in g["argv"]])
    args.cuda, args.hetero, args.edge_num = True, True, 1
    args.size = g["x"].shape[1]
    x, labels = t(g["x"]).to(DEV), t(g["labels"]).to(DEV)
    args.nclass = int(labels.max()) + 1
    n = int(g["n"])
    idx = torch.as_tensor(g["indices"])
    adj = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), (n, n)).to(DEV)
    seed_all(4)
    enc = edis.DISGAT(args, nfeat=args.size, nhid=args.nhid, nclass=args.nhid, nheads=args.nhead,
                      dropout=args.dropout).to(DEV)
    cls = T.ClsTrainer(args, enc, labels, 1.0)
    load_into(enc, g, "enc0.")
    load_into(cls.fuse1, g, "cls0.fuse1.")
    load_into(cls.fuse2, g, "cls0.fuse2.")
    for m in cls.models:
        m.eval()
    torch.manual_seed(21)
    np.random.seed(21)
    with torch.no_grad():
        dist, at_cor, feat_cor = cls.analyze_disentangle(x, adj)
    assert torch.equal(torch.rand(3), t(c["torch_rand_after"]))          # consumed the same RNG stream
    for layer in range(2):
        assert abs(dist[layer] - float(c["at_distance"][layer])) < 1e-5
        assert_close(at_cor[layer].cpu(), t(c["at_cor%d" % layer]), 2e-5, "at_cor%d" % layer)
        assert_close(feat_cor[layer].cpu(), t(c["feat_cor%d" % layer]), 2e-5, "feat_cor%d" % layer)
