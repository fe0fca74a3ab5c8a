#!/usr/bin/env python
"""Build libbitsandbytes_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python bitsandbytes-sycl_b200/build.py [--force] [--verbose]

The .so lands in bitsandbytes-sycl_b200/bnb_b200/ (git-ignored, travels to the GPU box with gpurun).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "bnb_b200")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(OUT_DIR, "libbitsandbytes_b200.so")
SOURCES = ["c_api.cu", "quant_blockwise.cu", "gemv_4bit.cu", "gemm_4bit.cu", "int8_quant.cu", "int8_fused.cu", "igemm.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    # bit-exactness: never contract a*b+c behind our back; IEEE div/sqrt; keep denormals
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
] + os.environ.get("BNB_EXTRA_NVCC_FLAGS", "").split()


def _digest(paths):
    h = hashlib.sha1()
    for p in sorted(paths):
        h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "bnb_b200.h")]
    stamp = os.path.join(OBJ_DIR, "stamp")
    digest = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):   # GPU box without a matching source change: use the shipped .so
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libbitsandbytes_b200.so")

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp, "w").write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
