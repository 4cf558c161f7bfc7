#!/usr/bin/env python3
"""Build oracle/_ref/libref_aligner.so from the reference sources where they lie.

TEST INFRASTRUCTURE ONLY.  Recipe: extract the Eigen 3.3.7 headers the reference
vendors as cmake/eigen-3.3.7.zip (python zipfile — there is no unzip here), then g++ the three
aligner translation units of /root/reference/src/aligner directly (we do not run the reference's
cmake) together with our C-ABI wrapper oracle/ref_harness.cpp.

Flags follow CMakeLists.txt:9 of the reference, with two documented deviations:
  * `-include cstdint -include functional`: similaritymatrix.h:9 uses uint8_t / std::function
    without the headers (GCC 13 no longer pulls them in transitively);
  * `-march=x86-64-v3` (AVX2, what the u8 kernel needs) instead of `-march=native`, so that the
    .so built in this container also runs on the GPU box's host CPU.
The library is the SERIAL build (no -DUSEOMP): SURVEY F7 — the USEOMP build of
OMPParallelLocalAligner is racy and its results are not deterministic.  OpenMP is enabled only
for the harness-level loop over reads in ref_bench_reads (the reference's own `#pragma omp`
lines sit behind `#ifdef USEOMP`, so -fopenmp does not change them).

Outputs: oracle/_ref/libref_aligner.so (git-ignored; travels to the GPU box with gpurun) and
oracle/_eigen/ (git-ignored AND gpurun-ignored header scratch).
"""
import hashlib
import os
import subprocess
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PGS_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
EIGEN = os.path.join(HERE, "_eigen")
EIGEN_MD5 = None  # cmake/GetEigen.cmake pins an md5; we report the one we saw


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REF, "src", "aligner", "similaritymatrix.cpp"))


def build(force: bool = False, useomp: bool = False) -> str:
    """useomp=False: the SERIAL build (the oracle).  useomp=True: a second library compiled with the reference's own
    -DUSEOMP switch (CMakeLists.txt:38-44) — its OpenMP pragmas are live; used for TIMING ONLY."""
    lib = os.path.join(OUT, "libref_aligner_omp.so" if useomp else "libref_aligner.so")
    if not have_reference():
        if os.path.isfile(lib):
            return lib  # GPU box: prebuilt file travelled with the snapshot
        raise RuntimeError(f"reference not found at {REF} and no prebuilt {lib}")
    srcs = [os.path.join(REF, "src", "aligner", f) for f in ("similaritymatrix.cpp", "smithwaterman.cpp", "plocalaligner.cpp")]
    harness = os.path.join(HERE, "ref_harness.cpp")
    if not force and os.path.isfile(lib) and all(os.path.getmtime(lib) > os.path.getmtime(s) for s in srcs + [harness, __file__]):
        return lib
    os.makedirs(OUT, exist_ok=True)
    inc = os.path.join(EIGEN, "eigen-eigen-323c052e1731")
    if not os.path.isdir(inc):
        zpath = os.path.join(REF, "cmake", "eigen-3.3.7.zip")
        with open(zpath, "rb") as f:
            md5 = hashlib.md5(f.read()).hexdigest()
        print(f"[build_ref] extracting {zpath} (md5 {md5})")
        with zipfile.ZipFile(zpath) as z:
            z.extractall(EIGEN)
    cmd = ["g++", "-Ofast", "-march=x86-64-v3", "-std=c++17", "-mavx", "-ffast-math", "-ftree-loop-if-convert",
           "-fopenmp", "-fPIC", "-shared", "-include", "cstdint", "-include", "functional",
           "-I", inc, "-I", os.path.join(REF, "src", "aligner"), "-o", lib, harness] + srcs
    if useomp:
        cmd.insert(1, "-DUSEOMP")
    print("[build_ref]", " ".join(cmd))
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build(force="--force" in sys.argv, useomp=True))
