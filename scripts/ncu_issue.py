"""Turn `ncu --set full` raw CSV (ncu -i X.ncu-rep --page raw --csv) into per-kernel-class issue / pipe statistics:
profiles/<round>_issue.json (bench.py roofline.issue) + a markdown table.   usage: ncu_issue.py out.json tag raw1.csv [raw2.csv ...]"""
import collections
import csv
import json
import re
import sys

import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_traffic import CLASSES, klass, short  # noqa: F401,E402

M = {"us": "gpu__time_duration.sum", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fmai": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "inst": "smsp__inst_executed.sum",
     "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum", "regs": "launch__registers_per_thread",
     "l1": "l1tex__throughput.avg.pct_of_peak_sustained_active", "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed"}
UNIT = {"us": {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}, "dram_rd": {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9},
        "dram_wr": {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}}


def main():
    out, tag, paths = sys.argv[1], sys.argv[2], sys.argv[3:]
    path = ", ".join(paths)
    recs = []
    for one in paths:                       # units are chosen per report: convert file by file
        rows = list(csv.reader(l for l in open(one) if not l.startswith("==")))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            d = {"name": r[col["Kernel Name"]]}
            for k, m in M.items():
                if m in col and r[col[m]] not in ("", "n/a"):
                    v = float(r[col[m]].replace(",", ""))
                    v *= UNIT.get(k, {}).get(units[col[m]], 1)
                    d[k] = v
            recs.append(d)
    by = collections.defaultdict(list)
    for d in recs:
        by[klass(d["name"])].append(d)
    res = {}
    for c, ds in by.items():
        t = sum(x.get("us", 0) for x in ds)
        w = lambda k: sum(x.get(k, 0) * x.get("us", 0) for x in ds) / t if t else 0.0   # time-weighted mean
        res[c] = {"launches": len(ds), "us_under_ncu": round(t, 1), "issue_frac": round(w("issue") / 100, 4), "fma_pipe_frac": round(w("fma") / 100, 4),
                  "fma_inst_frac": round(w("fmai") / 100, 4),
                  "tensor_pipe_frac": round(w("tensor") / 100, 4), "alu_frac": round(w("alu") / 100, 4), "xu_frac": round(w("xu") / 100, 4),
                  "warp_inst_per_launch_M": round(sum(x.get("inst", 0) for x in ds) / len(ds) / 1e6, 2)}
    res["_provenance"] = f"ncu --set full --clock-control none, {tag}; {path}"
    json.dump(res, open(out, "w"), indent=1)
    print("| kernel | us | DRAM rd MB | DRAM wr MB | L1 % | L2 % | tensor % | issue % | fma pipe % | regs | warp inst (M) |\n|---|---|---|---|---|---|---|---|---|---|---|")
    for d in recs:
        g = lambda k, s=1.0: f"{d.get(k, 0) * s:.1f}"
        print(f"| `{short(d['name'])}` | {g('us')} | {g('dram_rd', 1e-6)} | {g('dram_wr', 1e-6)} | {g('l1')} | {g('l2')} | {g('tensor')} | {g('issue')} | {g('fma')} | "
              f"{int(d.get('regs', 0))} | {g('inst', 1e-6)} |")


if __name__ == "__main__":
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    main()
