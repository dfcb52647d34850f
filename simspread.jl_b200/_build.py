"""In-tree build of libsimspread_b200.so (nvcc, sm_100a only).  No torch, no JIT cache: the built
library lives in simspread.jl_b200/lib/ so that it travels with the repository snapshot."""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libsimspread_b200.so"
SOURCES = ["ss_api.cu", "ss_elementwise.cu", "ss_gemm.cu", "ss_csr.cu", "ss_rank.cu", "ss_sparse.cu", "ss_umma.cu", "ss_recsys.cu", "ss_transfer.cu", "ss_comm.cu", "ss_folds.cu", "ss_similarity.cu", "ss_io.cu", "ss_tsparse.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def lib_path() -> str:
    return os.path.join(LIBDIR, LIBNAME)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libsimspread_b200.so")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, "ss_common.cuh"),
               os.path.join(HERE, "..", "include", "simspread_b200.h")]

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return o

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    out = lib_path()
    if force or _stale(out, objs):
        cmd = [nvcc, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                      "-Xcompiler", "-fPIC", "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in os.sys.argv, verbose=True))
