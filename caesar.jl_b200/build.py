"""Builds libcaesar_b200.so (sm_100a only) in-tree with nvcc.  `python caesar.jl_b200/build.py [--force]`."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libcaesar_b200.so")
SRCS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
DEPS = SRCS + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + [os.path.join(HERE, "..", "include", "caesar_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-ccbin", "/usr/bin/g++", "-Xptxas", "-v" if os.environ.get("CB2_PTXAS_V") else "-O3"] + os.environ.get("CB2_NVCC_EXTRA", "").split()
HDRS = [f for f in DEPS if not f.endswith(".cu")]


def build(force=False, verbose=False):
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(f) for f in DEPS):
        return SO
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SRCS:
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(f) for f in [src] + HDRS):
            continue  # object is newer than its source and every header
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    ok = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose or os.environ.get("CB2_PTXAS_V"):
            sys.stderr.write(out)
        ok = ok and p.returncode == 0
    if not ok:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, "-ccbin", "/usr/bin/g++"] + objs)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
