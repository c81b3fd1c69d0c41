"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total ms, share.
usage: python scripts/launch_summary.py launches.csv [first_fraction last_fraction]   (e.g. 0.5 1.0 = the second half of the launches)"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    lo, hi = (float(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (0.0, 1.0)
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    rows = [d for d in rows if "gpu__time_duration" in d.get("Metric Name", "gpu__time_duration")]
    rows = rows[int(len(rows) * lo):int(len(rows) * hi)]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in rows:
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        v, u = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
        ms = v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot >= 0.002:
            print(f"| `{k[:100]}` | {v[0]} | {v[1]:.3f} | {v[1] / tot:.3f} |")
    print(f"| total | {sum(v[0] for v in agg.values())} | {tot:.3f} | 1 |")


if __name__ == "__main__":
    main()
