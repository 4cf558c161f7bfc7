"""Build libswb200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.  No torch involved.

The wavefront kernels are templates over the rows-per-lane count R; csrc/sw_inst.cu is compiled once per R
(in parallel) and the objects are linked with the host side csrc/swb200.cu."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libswb200.so")
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
R_SET = [2, 4, 5, 8, 12, 16, 19, 24, 32]   # must match kRSet / SWB_DECL in csrc/swb200.cu
DEPS = [os.path.join(CSRC, f) for f in ("swb200.cu", "sw_inst.cu", "sw_core.cuh", "sw_qs.cuh")] + [os.path.join(os.path.dirname(HERE), "include", "swb200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _run(cmd):
    print("[swb200 build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force=False, verbose=False, checked=False):
    """checked=True: the same library with -DSWB_CHECKED (own bounds checks at every work-buffer store and pass-2 ring access,
    see sw_core.cuh) as libswb200_checked.so — the stand-in for compute-sanitizer memcheck, run by a GPU test."""
    global LIB, OBJ
    if checked:
        lib, obj = os.path.join(HERE, "libswb200_checked.so"), os.path.join(HERE, "build", "checked")
    else:
        lib, obj = os.path.join(HERE, "libswb200.so"), os.path.join(HERE, "build")
    if not force and os.path.isfile(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in DEPS + [__file__]):
        return lib
    LIB, OBJ = lib, obj
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    extra = (["-Xptxas", "-v"] if verbose else []) + os.environ.get("SWB_NVCC_EXTRA", "").split() + (["-DSWB_CHECKED"] if checked else [])
    jobs = [[nvcc] + CFLAGS + extra + ["-c", os.path.join(CSRC, "swb200.cu"), "-o", os.path.join(OBJ, "swb200.o")]]
    for r in R_SET:
        jobs.append([nvcc] + CFLAGS + extra + [f"-DSWB_R={r}", "-c", os.path.join(CSRC, "sw_inst.cu"), "-o", os.path.join(OBJ, f"sw_inst_r{r}.o")])
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        list(ex.map(_run, jobs))
    objs = [os.path.join(OBJ, "swb200.o")] + [os.path.join(OBJ, f"sw_inst_r{r}.o") for r in R_SET]
    _run([nvcc] + ARCH + ["-shared", "-cudart", "static", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, checked="--checked" in sys.argv))
