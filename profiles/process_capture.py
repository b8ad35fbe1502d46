#!/usr/bin/env python
"""Turn the gpurun_out/<tag>_* files written by profiles/capture.sh into the committed summaries:
    python profiles/process_capture.py r1d
-> profiles/<tag>_launches_full.csv, <tag>_dram_full.csv (copies), <tag>_launch_shares.json,
   <tag>_dram_full_summary.json, <tag>_med_set_full_summary.json, traffic.json."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def rows_of(path):
    return [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]


# ---- DRAM traffic per layer-kernel launch (full size)
d = collections.OrderedDict()
for r in rows_of(os.path.join(G, tag + "_dram_full.csv")):
    k = (int(r[0]), r[4].replace("(LayerArgs)", "").replace("void ", ""))
    d.setdefault(k, {})[r[12]] = float(r[14])
dram, agg = [], collections.defaultdict(list)
for (i, name), m in d.items():
    t = m["gpu__time_duration.sum"] / 1e6
    rd, wr = m["dram__bytes_read.sum"] / 1e9, m["dram__bytes_write.sum"] / 1e9
    dram.append({"launch": i, "kernel": name, "ms": round(t, 3), "dram_read_GB": round(rd, 2),
                 "dram_write_GB": round(wr, 2), "dram_GBps": round((rd + wr) / t * 1e3, 1),
                 "l2_hit_pct": round(m.get("lts__t_sector_hit_rate.pct", 0), 1)})
    agg["disga_" + re.match(r"k_disga_(\w+?)<", name).group(1)].append((rd + wr) * 1e9)
json.dump(dram, open(os.path.join(P, tag + "_dram_full_summary.json"), "w"), indent=1)
traffic = {k: sum(v) / len(v) for k, v in agg.items()}
traffic["_source"] = ("profiles/%s_dram_full.csv: ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                      "config A full size (N=2.4M, E=63.06M), mean over the captured launches" % tag)
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)

# ---- launch shares of one step (full size)
rows = rows_of(os.path.join(G, tag + "_launches_full.csv"))
idx = [i for i, r in enumerate(rows) if "k_disga_fwd" in r[4]]
step_len, first = idx[2] - idx[0], idx[0]
step = rows[idx[-2] - first: idx[-2] - first + step_len]
tot = sum(float(r[-1]) for r in step) / 1e6
grp = collections.OrderedDict()
for r in step:
    n = r[4]
    if "k_disga" in n or "k_combine" in n:
        key = re.sub(r"[<(].*", "", n).replace("void ", "").replace("edis::", "")
    elif "gemm" in n.lower() or "cutlass" in n.lower() or "cublas" in n.lower():
        key = "dense GEMMs (cuBLAS/CUTLASS via torch)"
    else:
        key = "torch elementwise / reductions"
    g = grp.setdefault(key, [0, 0.0])
    g[0] += 1
    g[1] += float(r[-1]) / 1e6
json.dump({"launches_in_step": len(step), "sum_ms_serialised": tot,
           "groups": {k: {"launches": v[0], "ms": v[1], "share": v[1] / tot} for k, v in grp.items()}},
          open(os.path.join(P, tag + "_launch_shares.json"), "w"), indent=1)
for f in ("_launches_full.csv", "_dram_full.csv"):
    shutil.copy(os.path.join(G, tag + f), os.path.join(P, tag + f))
subprocess.check_call([sys.executable, os.path.join(P, "ncu_summary.py"), os.path.join(G, tag + "_med_set_full.ncu-rep"),
                       "--stalls", "--json", os.path.join(P, tag + "_med_set_full_summary.json")],
                      stdout=subprocess.DEVNULL)
print(json.dumps({"traffic": {k: round(v / 1e9, 1) for k, v in traffic.items() if not k.startswith("_")},
                  "step_ms": round(tot, 1), "groups": {k: round(v[1], 1) for k, v in grp.items()}}, indent=1))
for r in dram[:4]:
    print(r)
for r in json.load(open(os.path.join(P, tag + "_med_set_full_summary.json")))[1:4]:
    print({k: r[k] for k in ("kernel", "time_ms", "dram_GBps", "issue_active_pct", "warps_active_pct", "regs_per_thread", "inst_executed", "l2_hit_pct")}, r["stalls_warps_per_issue"])
