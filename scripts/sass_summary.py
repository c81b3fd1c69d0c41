"""SASS evidence for the hot kernels of libkdlae_b200.so: per kernel the counts of the Blackwell-specific instructions (UTCHMMA =
tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store, FFMA2 / FMUL2 = packed fp32 math, SYNCS = mbarrier) and an
excerpt of the densest FFMA2 / UTCHMMA region.   usage: python scripts/sass_summary.py > profiles/r02_sass.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rethink_acoustic_image_enhancement_b200", "libkdlae_b200.so")
HOT = ["k_pwdw_f2ILi1ELi0", "k_pwdw_tILi1", "k_pwdw_tILi0", "k_conv_gemm_tcILi1ELi20", "k_conv3_tcILi1", "k_mdta_gram_tcILi1",
       "k_gemm_tf32", "k_wgrad_tf32", "k_gram_tf32"]
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "FFMA2", "FMUL2", "FFMA", "MUFU", "SYNCS", "LDS", "STS", "LDG", "STG", "F2FP", "BAR"]

sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip()))
print("# SASS summary of the hot kernels (cuobjdump -sass libkdlae_b200.so, sm_100a)\n")
print("| kernel | instructions | " + " | ".join(OPS) + " |\n|---|---|" + "---|" * len(OPS))
picked = []
for key in HOT:
    name = next((f for f in funcs if key in f), None)
    if not name:
        continue
    ins = funcs[name]
    cnt = collections.Counter()
    for l in ins:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            cnt[m.group(1)] += 1
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.search(r"(k_[a-z0-9_]+<[^>]*>)", demangled)
    label = short.group(1) if short else demangled[:60]
    print(f"| `{label}` | {len(ins)} | " + " | ".join(str(cnt.get(o, 0)) for o in OPS) + " |")
    picked.append((label, ins))
for label, ins in picked:
    # densest 40-instruction window for the kernel's characteristic op
    key = "UTCHMMA" if "gemm" in label or "conv3" in label or "gram" in label else "FFMA2"
    best, bi = -1, 0
    for i in range(0, max(1, len(ins) - 40)):
        c = sum(1 for l in ins[i:i + 40] if key in l)
        if c > best:
            best, bi = c, i
    print(f"\n## `{label}`: densest {key} window ({best} of 40 instructions)\n\n```")
    print("\n".join(l[:110] for l in ins[bi:bi + 40]))
    print("```")
