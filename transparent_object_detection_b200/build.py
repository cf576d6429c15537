"""Build libtod.so (sm_100a only) in-tree with nvcc.  No torch headers are involved: the library is a
plain C ABI (include/tod.h) over hand-written CUDA."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libtod.so")

# file -> extra flags.  -fmad=false where results must reproduce the reference's float32 arithmetic bit for bit.
SOURCES = {
    "capi.cu": [],
    "conv_tcgen05.cu": [],
    "conv_halo_tcgen05.cu": [],
    "stem_conv.cu": [],
    "stem_u8_tcgen05.cu": [],
    "sppf_pool.cu": [],
    "letterbox.cu": [],
    "cbam.cu": [],
    "softmax.cu": [],
    "attention_tcgen05.cu": [],
    "head_decode.cu": ["-fmad=false"],
    "nms.cu": ["-fmad=false"],
}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libtod.so cannot be built (there is no fallback path)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "tod.h"))
    objs, jobs = [], []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(PKG, "build", src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def compile_one(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    if jobs:   # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            list(pool.map(compile_one, jobs))
    if force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
