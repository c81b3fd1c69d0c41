"""Turn an ncu launch list (csv of `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`) into
(1) per-kernel-class DRAM traffic per launch -> profiles/<round>_traffic.json (bench.py's roofline.traffic) and
(2) a markdown launch table (kernel, launches, total ms, share).   usage: ncu_traffic.py launches.csv out.json micro_batch tag"""
import csv, json, sys, collections, re

CLASSES = [("k_pwdw_", "fused_conv1x1_dwconv3x3_tcgen05"), ("k_conv_gemm_tc", "conv_gemm_tcgen05"), ("k_conv3_tc", "conv_gemm_tcgen05"),
           ("k_conv_gemm_simt", "conv_gemm_simt"), ("k_dwconv", "dwconv3x3"), ("k_mdta_gram", "mdta_gram"),
           ("k_mdta_softmax", "mdta_softmax_fold"), ("k_mdta_project", "mdta_softmax_fold"), ("k_ln_stats", "ln_stats"),
           ("k_conv_few", "small_channel_conv")]

def klass(name):
    for pat, c in CLASSES:
        if pat in name:
            return c
    return "other"

def short(name):
    m = re.search(r"(k_[A-Za-z0-9_]+(<[^>]*>)?)", name)
    return m.group(1) if m else name[:40]

def main():
    path, out, mb, tag = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    iname, imet, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iid = hdr.index("ID")
    per = collections.defaultdict(dict)
    names = {}
    for r in rows[1:]:
        if len(r) <= ival or not r[iid].isdigit():
            continue
        v = float(r[ival].replace(",", ""))
        u = r[iunit]
        if r[imet].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        elif r[imet].startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)     # -> us
        per[int(r[iid])][r[imet]] = v
        names[int(r[iid])] = r[iname]
    cls = collections.defaultdict(lambda: [0, 0.0, 0.0])
    kern = collections.defaultdict(lambda: [0, 0.0])
    for i, m in per.items():
        c = cls[klass(names[i])]
        c[0] += 1
        c[1] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
        c[2] += m.get("gpu__time_duration.sum", 0)
        k = kern[short(names[i])]
        k[0] += 1; k[1] += m.get("gpu__time_duration.sum", 0)
    tj = {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / v[0], "us_per_launch_under_ncu": v[2] / v[0]} for k, v in cls.items()}
    tj["_micro_batch"] = mb
    tj["_provenance"] = (f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one "
                         f"KDLAE-T forward of {mb} images (bench.py --batch {mb} --steps 1), {tag}; {path}")
    json.dump(tj, open(out, "w"), indent=1)
    tot = sum(v[1] for v in kern.values())
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.3f} | {v[1] / tot:.3f} |")
    print(f"| total | {sum(v[0] for v in kern.values())} | {tot / 1e3:.3f} | 1 |")

if __name__ == "__main__":
    main()
