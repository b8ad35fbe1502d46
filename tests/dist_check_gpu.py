"""Multi-GPU correctness check of the partitioned path with the real CUDA kernels and NCCL
(not a pytest: needs N GPUs).  Run under torchrun on a B200 box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29531 tests/dist_check_gpu.py

Every rank builds the same seeded hub graph; the ranks compute get_em and the SupEdge pair loss
over their destination ranges (halo exchange, all-gather / reduce-scatter, grad all-reduce) and rank
0 compares features, loss and encoder gradients against its own single-GPU run on the full graph.
Prints one JSON line with the maximum relative errors (relative to each tensor's max); exits non-zero
above 2e-5 (features, losses) / 5e-4 (gradients: fp32 sums cut differently per rank + atomic pair updates)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def main():
    import edgedisentangle_ssl_b200 as edis
    from edgedisentangle_ssl_b200 import functional as Fn
    from edgedisentangle_ssl_b200 import parallel as par
    from edgedisentangle_ssl_b200.graph import build_adjacency
    from edgedisentangle_ssl_b200.utils import get_parser
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, fin, C, D = 20000, 100, 8, 64
    rng = np.random.RandomState(0)
    hubs = rng.randint(0, 50, 60000)
    rows = np.concatenate([rng.randint(0, n, 200000), hubs])
    cols = np.concatenate([rng.randint(0, n, 200000), rng.randint(0, n, 60000)])
    idx, _ = build_adjacency(n, rows, cols)
    args = get_parser().parse_args(["--model=DISGAT", "--sparse", "--att=3", "--gnn_type=AT", "--nhead=%d" % C,
                                    "--nhid=%d" % D, "--dropout=0.0"])
    torch.manual_seed(0)
    enc = edis.DISGAT(args, nfeat=fin, nhid=D, nclass=D, nheads=C, dropout=0.0).to(dev).eval()
    fus = [edis.FuseLayer(args, C, nfeat=D).to(dev), edis.FuseLayer(args, C, nfeat=D).to(dev)]
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(n, fin, generator=gen).to(dev)
    R = torch.randn(n, D, generator=gen).to(dev)
    key = np.unique(np.concatenate([rng.randint(0, n * n, 600000), idx[0][::3] * n + idx[1][::3]]))
    lab = np.isin(key, idx[0] * n + idx[1]).astype(np.float32)
    params = [p for m in [enc] + fus for p in m.parameters()]

    rowptr = np.concatenate([[0], np.cumsum(np.bincount(idx[0], minlength=n))])
    bounds = par.row_ranges(rowptr, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    sel = (idx[0] >= lo) & (idx[0] < hi)
    part = par.Partition(rank, world, bounds, idx[0][sel], idx[1][sel]).attach_graph(dev)
    mine = (key // n >= lo) & (key // n < hi)
    pairs = torch.from_numpy(np.stack([key[mine] // n - lo, key[mine] % n])).to(dev)
    labels = torch.from_numpy(lab[mine]).to(dev)

    feats = par.get_em_partitioned(enc, fus, x[lo:hi], part)
    loss_em = (feats[-1] * R[lo:hi]).sum()
    loss_ssl = par.ssl_pair_loss_partitioned(enc, fus, x[lo:hi], part, [pairs], [labels], [(0, C)],
                                             [int(lab.sum())], [len(key)])
    (loss_em + 1000.0 * loss_ssl).backward()
    par.allreduce_grads(params)
    tot = torch.stack([loss_em.detach(), loss_ssl.detach()])
    dist.all_reduce(tot)
    got_feat = feats[-1].detach()
    got_grads = {k: v.grad.clone() for k, v in enc.named_parameters() if v.grad is not None}
    for p in params:
        p.grad = None

    ok = True
    if rank == 0:
        graph = edis.Graph(n, idx[0], idx[1], device=dev)
        f_ref = enc.get_em(x, graph, fus)
        l_em = (f_ref[-1] * R).sum()
        r = enc.traverse(x, graph, fus, aux=[torch.from_numpy(np.stack([key // n, key % n])).to(dev)],
                         need_layer2_agg=False)
        y = torch.from_numpy(lab).to(dev)
        l_ssl = sum(Fn.SslWmse.apply(a[0], y, int(lab.sum())) for a in r["aux"])
        (l_em + 1000.0 * l_ssl).backward()
        errs = {"feat": rel(got_feat, f_ref[-1][lo:hi].detach()), "loss_em": abs(float(tot[0]) - float(l_em)) / abs(float(l_em)),
                "loss_ssl": abs(float(tot[1]) - float(l_ssl)) / abs(float(l_ssl))}
        gerr = {k: rel(got_grads[k], v.grad) for k, v in enc.named_parameters() if v.grad is not None}
        errs["grad_max"] = max(gerr.values())
        errs["grad_worst"] = max(gerr, key=gerr.get)
        ok = errs["feat"] < 2e-5 and errs["loss_em"] < 2e-5 and errs["loss_ssl"] < 2e-5 and errs["grad_max"] < 5e-4
        print(json.dumps({"world": world, "n": n, "edges": int(idx.shape[1]), "pairs": int(len(key)),
                          "halo_rank0": int(len(part.halo_ids)), "errors": errs, "ok": ok}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
