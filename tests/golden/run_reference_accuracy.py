"""Run the UNMODIFIED reference CLI (main.py) on CPU for a few seeds and record the test accuracy
curve -> tests/golden/accuracy_ref.json.  Build-container only (needs /root/reference).

    python tests/golden/run_reference_accuracy.py cora 201 4 5 6 7 8
    python tests/golden/run_reference_accuracy.py chameleon:SAGE 41 4 5 6 7 8     (gnn_type after a colon)

The reference tree is read-only, so a scratch directory gets symlinks to its sources and data;
cora / cora_full get the deterministic synthetic features of SURVEY 8(d) (their feature blobs are
missing from the snapshot).  Flags = example_bashs/Example_cora_full.sh:38.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def scratch_tree(ds):
    from edgedisentangle_ssl_b200.data_load import synthetic_features
    tmp = tempfile.mkdtemp(prefix="ref_acc_")
    for f in os.listdir(REF):
        if f.endswith(".py"):
            os.symlink(os.path.join(REF, f), os.path.join(tmp, f))
    d = os.path.join(tmp, "data", ds)
    os.makedirs(d)
    src = os.path.join(REF, "data", ds)
    for f in os.listdir(src):
        os.symlink(os.path.join(src, f), os.path.join(d, f))
    if not os.path.exists(os.path.join(d, "feature_new.npy")):
        np.save(os.path.join(d, "feature_new.npy"), synthetic_features(np.load(os.path.join(src, "label.npy"))))
    return tmp


def main():
    ds, epochs, seeds = sys.argv[1], int(sys.argv[2]), [int(s) for s in sys.argv[3:]]
    ds, _, gnn = ds.partition(":")
    gnn = gnn or "AT"
    key = ds if gnn == "AT" else "%s_%s" % (ds, gnn)
    out_path = os.path.join(HERE, "accuracy_ref.json")
    res = json.load(open(out_path)) if os.path.exists(out_path) else {}
    tmp = scratch_tree(ds)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=os.path.join(HERE, "_stubs"))
    for seed in seeds:
        cmd = [sys.executable, "main.py", "--no-cuda", "--seed=%d" % seed, "--model=DISGAT", "--used_edge=1",
               "--finetune", "--downstream=CLS", "--down_weight=1.0", "--steps=5", "--nhead=4", "--dataset=" + ds,
               "--pretrain", "SupEdge", "DisEdge", "DifHead", "--pre_weight", "1", "1", "1", "--pre_edge", "1", "1",
               "1", "--sparse", "--att=3", "--constrain_layer=0", "--epochs=%d" % epochs, "--gnn_type=" + gnn]
        p = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True)
        if p.returncode != 0:
            print(p.stderr[-2000:])
            raise SystemExit("reference run failed")
        accs = [float(m) for m in re.findall(r"Test set results: loss= [-\d.e]+ accuracy= ([\d.]+)", p.stdout)]
        res = json.load(open(out_path)) if os.path.exists(out_path) else res
        res.setdefault(key, {})["seed%d" % seed] = {"epochs": epochs, "test_acc_every_40": accs}
        print(key, seed, accs, flush=True)
        json.dump(res, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
