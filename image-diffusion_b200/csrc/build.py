"""Builds libidf_b200.so (sm_100a only) in-tree with nvcc. No torch dependency: the library is a plain C ABI."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
SOURCES = ["host.cu", "igemm.cu", "attention.cu", "norm.cu", "pointwise.cu", "wgrad.cu", "train.cu", "attention_bwd.cu"]
HEADERS = ["common.cuh", "host.h", os.path.join(ROOT, "include", "idf_b200.h")]
LIB = os.path.join(PKG, "idf_b200", "libidf_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build the sm_100a kernels")
    return exe


def _stamp() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS + [os.path.abspath(__file__)]:
        path = f if os.path.isabs(f) else os.path.join(HERE, f)
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build_trace(tag: str = "trace", defines=("-DIDF_ATTN_TRACE",)) -> str:
    """Variant library for experiments, selected at run time with IDF_B200_LIB (the product library is never touched):
    `--trace`: the attention kernel's clock64 trace hooks (tools/trace_attn.py); `--variant TAG -DNAME=VALUE ...`: any
    compile-time switch (e.g. -DIDF_ATTN_POLY=4)."""
    nvcc = _nvcc()
    lib = os.path.join(PKG, "idf_b200", f"libidf_b200_{tag}.so")
    objs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, f"{tag}_" + src.replace(".cu", ".o"))
        res = subprocess.run([nvcc, *NVCC_FLAGS, *defines, "-c", os.path.join(HERE, src), "-o", obj],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
        objs.append(obj)
    res = subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp_file = os.path.join(OBJ_DIR, "stamp.txt")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    if "--trace" in sys.argv:
        os.makedirs(OBJ_DIR, exist_ok=True)
        print(build_trace())
    elif "--variant" in sys.argv:
        os.makedirs(OBJ_DIR, exist_ok=True)
        k = sys.argv.index("--variant")
        print(build_trace(sys.argv[k + 1], tuple(a for a in sys.argv[k + 2:] if a.startswith("-D"))))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
