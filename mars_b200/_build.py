"""Build libmarsgb.so in-tree with nvcc for sm_100a (the only target; no multi-arch dispatch)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmarsgb.so")
SOURCES = ["mgb_util.cu", "mgb_partition.cu", "mgb_split.cu", "mgb_agg.cu", "mgb_bucket.cu", "mgb_dense.cu", "mgb_mid.cu", "mgb_tiny.cu", "mgb_nccl.cu", "mgb_pack.cu", "mgb_hostpack.cpp",
           "mgb_host.cu"]
# (source, object suffix, extra defines): the dense kernels are instantiated in four units to compile in parallel
VARIANTS = [("mgb_dense_inst.cu", f"_{nk}_{th}", [f"-DDN_NK={nk}", f"-DDN_TH={th}"]) for nk in (1, 2) for th in (256, 512)]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh"))
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
] + os.environ.get("MGB_NVCC_EXTRA", "").split()      # experiments only (e.g. -DBK_THREADS=128)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmarsgb.so cannot be built (set NVCC=/path/to/nvcc)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, v[0]) for v in VARIANTS] + [
        os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(HERE, "..", "include", "mgb.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(job):
        src, suffix, defines = job
        obj = os.path.join(objdir, os.path.splitext(src)[0] + suffix + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    jobs = VARIANTS + [(s, "", []) for s in SOURCES]   # longest first
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, jobs))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", LIB, *objs, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
