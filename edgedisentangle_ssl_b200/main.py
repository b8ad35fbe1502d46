"""CLI entry: `python -m edgedisentangle_ssl_b200.main --model=DISGAT --sparse ...`.

Same flags and epoch-loop order as /root/reference/main.py (arg post-processing 27-42, seeding
44-48, dataset dispatch 70-111, trainer construction 237-258, loop 270-352): every 40 epochs a
CLS test, then `--steps` CLS train steps, then one train step per SSL trainer.  Only the DISGAT /
--sparse route is built (everything else in main.py is out of scope).
"""
import os
import random
import sys
import time

import numpy as np
import torch

from . import data_load, models, trainer, utils

DATASETS = ("chameleon", "squirrel", "cora_full", "deezer", "arxiv", "BlogCatalog", "cora")


def checkpoint_path(args, epoch, root="."):
    """Same location and name as the reference's save_model / load_model (main.py:214-235)."""
    d = os.path.join(root, "checkpoint", args.dataset, "{}_used_edge{}_weight{}_reg{}".format(
        args.model, args.used_edge, args.pre_weight, args.reg))
    return d, os.path.join(d, "pretrain_{}_{}.pth".format(args.pretrain, epoch))


def save_model(args, encoder, trainers, epoch, root="."):
    """main.py:214-227 writes {'encoder': state_dict}.  That key keeps its meaning (the reference's
    load_model reads these files unchanged); the per-trainer fusers / classifiers and every Adam
    state ride along under 'trainers', so a run can be RESUMED, not only warm-started."""
    d, path = checkpoint_path(args, epoch, root)
    os.makedirs(d, exist_ok=True)
    from .layers import dropout_stream_state
    content = {"encoder": encoder.state_dict(), "epoch": epoch, "trainers": [],
               "dropout_stream": dropout_stream_state()}
    for tr in trainers:
        content["trainers"].append({"class": type(tr).__name__,
                                    "models": [m.state_dict() for m in tr.models[1:]],     # [0] is the encoder
                                    "optimizers": [o.state_dict() for o in tr.models_opt]})
    torch.save(content, path)
    print("successfully saved: {}".format(epoch))
    return path


def load_model(args, encoder, trainers=(), root="."):
    """main.py:229-235 (+ the trainer states when the file carries them)."""
    _, path = checkpoint_path(args, args.load, root)
    content = torch.load(path, map_location=lambda storage, loc: storage)
    encoder.load_state_dict(content["encoder"])
    for tr, st in zip(trainers, content.get("trainers", [])):
        if st["class"] != type(tr).__name__:
            raise SystemExit("checkpoint trainer order differs: {} vs {}".format(st["class"], type(tr).__name__))
        for m, sd in zip(tr.models[1:], st["models"]):
            m.load_state_dict(sd)
        for o, sd in zip(tr.models_opt, st["optimizers"]):
            o.load_state_dict(sd)
    if "dropout_stream" in content:
        from .layers import set_dropout_stream_state
        set_dropout_stream_state(content["dropout_stream"])
    print("successfully loaded: {}".format(args.load))
    return content.get("epoch", args.load)


def run(argv=None, data_root="data", epoch_hook=None, ckpt_root=".", save_every=0):
    args = utils.get_parser().parse_args(argv)
    args.log = True
    args.cuda = not args.no_cuda and torch.cuda.is_available()
    if not args.cuda:
        raise SystemExit("edgedisentangle_ssl_b200 needs a CUDA device (no CPU fallback)")
    if args.model != "DISGAT" or not args.sparse:
        raise SystemExit("only --model=DISGAT --sparse is built on the B200 path")
    if args.pretrain is not None or args.hnn:
        args.hetero = True
    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    if args.pre_weight is None:
        args.pretrain = []
    if args.dataset not in DATASETS:
        raise SystemExit("no this dataset: {}".format(args.dataset))
    args.edge_num = 1
    adjs, features, labels = data_load.load_data(args, path=os.path.join(data_root, args.dataset) + "/",
                                                 dataset=args.dataset, edge_type=args.edge_num)
    args.size = features.shape[1]
    print("feature dimension: {}".format(args.size))
    args.nclass = labels.max().item() + 1
    print(args)

    encoder = models.DISGAT(args, nfeat=args.size, nhid=args.nhid, nclass=args.nhid, nheads=args.nhead,
                            dropout=args.dropout).cuda()
    features, labels = features.cuda(), labels.cuda()
    adjs = [a.cuda() for a in adjs] if args.hetero else adjs.cuda()

    def adj_of(k):
        return adjs[k - 1] if args.hetero else adjs

    ssl_trainers, ssl_labels = [], []
    for i, name in enumerate(args.pretrain or []):
        assert args.pre_edge[i] > 0, "edge index begins from 1"
        tr = trainer.SSL_TRAINERS[name](args, encoder, args.pre_weight[i])
        ssl_trainers.append(tr)
        if name != "DisEdge":
            ssl_labels.append(tr.get_label_all(features, adj_of(args.used_edge)))
        else:
            ssl_labels.append(tr.get_label_all(features, adj_of(args.used_edge), labels))
    down = []
    for i, name in enumerate(args.downstream or []):
        if name != "CLS":
            raise SystemExit("downstream 'Edge' is unfinished in the reference (README.md:22) and not built")
        down.append(trainer.ClsTrainer(args, encoder, labels, args.down_weight[0]))

    if args.load is not None:
        load_model(args, encoder, ssl_trainers + down, ckpt_root)
    t_total = time.time()
    history = []
    for epoch in range(args.epochs):
        t_epoch = time.time()
        log = {}
        if epoch % 40 == 0:
            for tr in down:
                log.update(tr.test([features, adj_of(args.used_edge)], labels, epoch))
            if args.case and down:
                # case study on disentanglement (main.py:290-302): inside the every-40-epochs block, right
                # after `test` (so every model is still in eval mode: no dropout in the statistics, and the
                # CPU / numpy RNG streams are consumed at the reference's positions), for the LAST
                # downstream trainer like the reference's leftover loop variable.  The scalars are what the
                # reference sends to tensorboard; its matplotlib heat maps are out of scope, the maps
                # themselves are returned
                with torch.no_grad():
                    dist, at_cor, feat_cor = down[-1].analyze_disentangle(features, adj_of(args.used_edge))
                log["att_correlation_layer1"], log["att_correlation_layer2"] = dist
                log["att_correlation_maps"] = [c.cpu() for c in at_cor]
                log["feat_correlation_maps"] = [c.cpu() for c in feat_cor]
        if args.finetune:
            for step in range(args.steps):
                for tr in down:
                    log.update(tr.train_step([features, adj_of(args.used_edge)], labels, epoch))
        for i, tr in enumerate(ssl_trainers):
            log.update(tr.train_step([features, adj_of(args.pre_edge[i])], ssl_labels[i]))
        torch.cuda.synchronize()
        log["epoch_ms"] = (time.time() - t_epoch) * 1e3
        history.append(log)
        if save_every and (epoch + 1) % save_every == 0:
            save_model(args, encoder, ssl_trainers + down, epoch, ckpt_root)
        if epoch_hook:
            epoch_hook(epoch, log)
    print("Optimization Finished!")
    print("Total time elapsed: {:.4f}s".format(time.time() - t_total))
    return history


if __name__ == "__main__":
    run(sys.argv[1:])
