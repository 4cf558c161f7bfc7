#!/usr/bin/env python3
"""Build tests/cpp/ref_test_localaligner: the reference's own test/test_localaligner.cpp with the aligner type swapped
(see ref_test_swap.cpp).  Needs /root/reference (build container only); googletest 1.8.1 and Eigen 3.3.7 come from the
zips the reference vendors under cmake/.  Outputs are git-ignored and travel to the GPU box with the snapshot."""
import os
import subprocess
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PGS_REFERENCE_DIR", "/root/reference")
GTEST = os.path.join(ROOT, "oracle", "_eigen", "googletest-release-1.8.1", "googletest")
EIGEN = os.path.join(ROOT, "oracle", "_eigen", "eigen-eigen-323c052e1731")
EXE = os.path.join(HERE, "ref_test_localaligner")


def build(force=False):
    if not os.path.isfile(os.path.join(REF, "test", "test_localaligner.cpp")):
        if os.path.isfile(EXE):
            return EXE
        raise RuntimeError("reference not found and no prebuilt ref_test_localaligner")
    src = os.path.join(HERE, "ref_test_swap.cpp")
    shim = os.path.join(ROOT, "parallel-genomeseq_b200", "cpp", "cuda_aligner.h")
    lib = os.path.join(ROOT, "parallel-genomeseq_b200", "libswb200.so")
    if not force and os.path.isfile(EXE) and all(os.path.getmtime(EXE) > os.path.getmtime(f) for f in (src, shim, lib, __file__)):
        return EXE
    for zname, probe in (("googletest-release-1.8.1.zip", GTEST), ("eigen-3.3.7.zip", EIGEN)):
        if not os.path.isdir(probe):
            with zipfile.ZipFile(os.path.join(REF, "cmake", zname)) as z:
                z.extractall(os.path.join(ROOT, "oracle", "_eigen"))
    gobj = os.path.join(ROOT, "oracle", "_eigen", "gtest-all.o")
    if not os.path.isfile(gobj):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-w", "-I", os.path.join(GTEST, "include"), "-I", GTEST, "-c", os.path.join(GTEST, "src", "gtest-all.cc"), "-o", gobj])
    cmd = ["g++", "-std=c++17", "-O2", "-w", "-include", "cstdint", "-include", "functional",
           f'-DREF_TEST_FILE="{os.path.join(REF, "test", "test_localaligner.cpp")}"',
           "-I", os.path.join(GTEST, "include"), "-I", EIGEN, "-I", os.path.join(REF, "src", "aligner"),
           src, gobj, "-o", EXE, "-L", os.path.dirname(lib), "-lswb200", "-lpthread", "-Wl,-rpath,$ORIGIN/../../parallel-genomeseq_b200"]
    print("[build_ref_tests]", " ".join(cmd))
    subprocess.check_call(cmd)
    return EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
