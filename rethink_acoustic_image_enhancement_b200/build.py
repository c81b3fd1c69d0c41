"""In-tree nvcc build of libkdlae_b200.so (sm_100a only).

    python -m rethink_acoustic_image_enhancement_b200.build [--force] [--verbose]

The shared library lands next to this file so it travels with the repo snapshot; it is never
installed into site-packages.  Objects are rebuilt only when a source or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "libkdlae_b200.so")
OBJ_DIR = os.path.join(PKG_DIR, "build")
SOURCES = ["api.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_tf32.cu", "conv3x3_tc.cu", "glue.cu", "dwconv_f2.cu", "pwdw_f2.cu", "pwdw_t.cu", "gram_tc.cu", "pack.cu", "prepost.cu", "metrics.cu", "train.cu", "teacher.cu", "student.cu", "asdqe.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libkdlae_b200.so")


def _newest_header() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return max(os.path.getmtime(h) for h in hs)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link the C-ABI shared library. Returns its path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = _newest_header()
    jobs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or not os.path.exists(op) or os.path.getmtime(op) < max(os.path.getmtime(sp), hdr_t):
            jobs.append((sp, op))

    def compile_one(job):
        sp, op = job
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", sp, "-o", op]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {sp}:\n{r.stdout}\n{r.stderr}")
        return sp, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for sp, log in ex.map(compile_one, jobs):
            if verbose:
                print(f"== {os.path.basename(sp)}\n{log}")
            else:
                spills = [l for l in log.splitlines() if "spill" in l and "0 bytes spill stores, 0 bytes spill loads" not in l]
                if spills:
                    print(f"[build] {os.path.basename(sp)}: {len(spills)} kernels with register spills")

    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
