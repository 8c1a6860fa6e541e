"""Build recipe for libtlb200.so (hand-written sm_100a CUDA + the C ABI of include/tlb200.h).

    python -m tensorly_b200.build          # incremental
    python -m tensorly_b200.build --force

nvcc cross-compiles for sm_100a without a GPU; the resulting .so lives in-tree
(tensorly_b200/csrc/libtlb200.so) so that it travels to the GPU box with the repo.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libtlb200.so")
OBJ_DIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: tensorly_b200 needs the CUDA toolkit to build libtlb200.so")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "tlb200.h"))
    return hs


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link libtlb200.so. Returns its path."""
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    # stamp of the sources this library is built from (checked by _lib.load()); rewritten only when it changes,
    # so that api.o is recompiled exactly then
    from ._lib import source_hash
    stamp = os.path.join(OBJ_DIR, "source_hash.inc")
    text = f'#define TLB200_SOURCE_HASH "{source_hash()}"\n'
    if not os.path.exists(stamp) or open(stamp).read() != text:
        with open(stamp, "w") as f:
            f.write(text)
    hdrs = _headers() + [stamp]
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB, *objs])
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print(path)
