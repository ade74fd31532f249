"""Builds an experiment variant of libcaesar_b200.so: `python bench/micro/build_variant.py NAME [-DFLAG ...]`.
Only product.cu (tree build + Gibbs kernel) is recompiled with the extra flags, by default with -DCB2_FAST_BUILD (the FP32 Pose3
instance only); every other object comes from the regular build.  Output: caesar.jl_b200/variants/lib_NAME.so (git-ignored,
travels to the GPU box); select it with CB2_LIB=... (caesar.jl_b200/_capi.py)."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, "caesar.jl_b200")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    if "--full" in flags:
        flags.remove("--full")
    else:
        flags.append("-DCB2_FAST_BUILD")
    os.makedirs(os.path.join(PKG, "variants"), exist_ok=True)
    obj = os.path.join(PKG, "variants", f"product_{name}.o")
    cmd = [NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin",
           "/usr/bin/g++"] + flags + ["-c", os.path.join(PKG, "csrc", "product.cu"), "-o", obj]
    subprocess.check_call(cmd)
    others = [o for o in glob.glob(os.path.join(PKG, "build", "*.o")) if not o.endswith("product.o")]
    so = os.path.join(PKG, "variants", f"lib_{name}.so")
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, "-ccbin", "/usr/bin/g++", obj] + others)
    print(so)


if __name__ == "__main__":
    main()
