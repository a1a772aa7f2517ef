"""Builds ``libmp2gpu.so`` in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmp2gpu.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["api.cu", "ntt.cu", "merkle.cu", "tables.cu", "prof.cu", "fri.cu", "selftest.cu", "sharded.cu", "quotient.cu", "permutation.cu", "transcript.cpp", "prover.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newest_dep() -> float:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "mp2gpu.h"))
    return max(os.path.getmtime(d) for d in deps)


def build_library(force: bool = False, verbose: bool = False, out: str = None) -> str:
    global LIB, OBJ_DIR
    if out:
        LIB = os.path.abspath(out)
        OBJ_DIR = LIB + ".objs"
        force = True
    extra = os.environ.get("MP2_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DMP2_HASH_BLOCK=512
    if extra:
        force = True
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_dep():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    env.pop("CC", None)   # the image's CC points at a gcc without libgomp specs; let nvcc pick PATH g++
    env.pop("CXX", None)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    import sys
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out))
