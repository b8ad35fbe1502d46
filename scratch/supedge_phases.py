"""Phase timing of one SupEdge step on the config-A graph (tuning aid, not a bench)."""
import os, sys, time, contextlib, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import edgedisentangle_ssl_b200 as edis
from edgedisentangle_ssl_b200 import trainer as T, functional as Fn
from edgedisentangle_ssl_b200.synthetic import power_law_graph
from edgedisentangle_ssl_b200.utils import get_parser
dev = torch.device("cuda:0")
n, raw = int(os.environ.get("N", 2_400_000)), int(os.environ.get("RAW", 30_600_000))
idx = power_law_graph(n, raw, seed=0)
graph = edis.Graph(n, idx[0], idx[1], device=dev)
args = get_parser().parse_args(["--model=DISGAT", "--sparse", "--att=3", "--gnn_type=AT", "--nhead=8", "--nhid=64", "--dropout=0.1"])
args.size, args.cuda = 100, True
torch.manual_seed(4)
enc = edis.DISGAT(args, nfeat=100, nhid=64, nclass=64, nheads=8, dropout=0.1).to(dev)
x = torch.randn(n, 100, device=dev)
os.environ["EDIS_SAMPLER"] = "device"
tr = T.SupEdgeTrainer(args, enc, 1.0)
lab = tr.get_label_all(x, graph)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(3):
    Fn.TIMER.reset(); Fn.TIMER.enabled = (it == 2)
    tr._begin_step()
    t0 = ev()
    label, masks = tr.sample_train(lab)
    t1 = ev()
    loss = tr._loss([x, graph], [label], masks)
    t2 = ev()
    (loss * tr.loss_weight).backward()
    t3 = ev()
    for opt in tr.models_opt: opt.step()
    t4 = ev()
    torch.cuda.synchronize()
    print("it %d pairs %d: sample %.1f fwd %.1f bwd %.1f adam %.1f ms  loss %.3e" % (it, masks[0].shape[1], t0.elapsed_time(t1), t1.elapsed_time(t2), t2.elapsed_time(t3), t3.elapsed_time(t4), float(loss)))
print({k: [round(x, 1) for x in v] for k, v in Fn.TIMER.durations_ms().items()})
Fn.TIMER.enabled = False
pj = masks[0][1]
torch.cuda.synchronize(); t=time.time(); perm = torch.sort(pj.to(torch.int32))[1].to(torch.int32); torch.cuda.synchronize(); print("col sort %.1f ms" % ((time.time()-t)*1e3))
# sampler internals
from edgedisentangle_ssl_b200.sampler import sample_pairs_device
key = lab.keys_on(dev)[0]
torch.cuda.synchronize(); t=time.time(); sample_pairs_device(n, key); torch.cuda.synchronize(); print("sampler alone %.1f ms" % ((time.time()-t)*1e3))
