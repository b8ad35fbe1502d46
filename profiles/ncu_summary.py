#!/usr/bin/env python
"""Summarise an .ncu-rep (captured with `ncu --set full`) into the few numbers DESIGN.md cites.

    python profiles/ncu_summary.py gpurun_out/x.ncu-rep [--stalls] [--json out.json]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import json
import subprocess
import sys

KEYS = [
    ("time_ms", "gpu__time_duration.sum"),
    ("dram_read_GB", "dram__bytes_read.sum"),
    ("dram_write_GB", "dram__bytes_write.sum"),
    ("dram_pct_of_ncu_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs_per_thread", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("static_smem_B", "launch__shared_mem_per_block_static"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("inst_executed", "smsp__inst_executed.sum"),
    ("local_load_sectors", "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    path = sys.argv[1]
    hdr, units, rows = load(path)
    res = []
    for r in rows:
        d = {"kernel": r[hdr.index("Kernel Name")][:90]}
        for name, key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = float(v.replace(",", ""))
                except ValueError:
                    pass
                u = units[i]
                if isinstance(v, float):
                    if u == "Gbyte":
                        pass
                    elif u == "Mbyte":
                        v /= 1e3
                    elif u == "byte" and name.endswith("_GB"):
                        v /= 1e9
                    elif u in ("us", "usecond") and name == "time_ms":
                        v /= 1e3
                    elif u in ("ns", "nsecond") and name == "time_ms":
                        v /= 1e6
                d[name] = v
        if "dram_read_GB" in d and "time_ms" in d:
            d["dram_GBps"] = (d["dram_read_GB"] + d.get("dram_write_GB", 0.0)) / (d["time_ms"] * 1e-3)
        if "--stalls" in sys.argv:
            st = {}
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        st[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
                    except ValueError:
                        pass
            d["stalls_warps_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:8])
        res.append(d)
    for d in res:
        print(json.dumps(d))
    if "--json" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
