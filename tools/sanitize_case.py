#!/usr/bin/env python3
"""A small batch through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck) and for the
SWB_CHECKED build of the library (own bounds checks; SWB_LIB_PATH=parallel-genomeseq_b200/libswb200_checked.so): batched
kernels (both arithmetic modes, chunked), row strips run by one warp, pipelined strips (warps synchronising through HBM
flags), the query-stationary kernels, wide lanes and the dense-matrix replay.  Sizes are tiny: the tools slow kernels
down by one to two orders of magnitude.  Results are checked against the oracle so that a clean log means a clean RUN.

    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as o  # noqa: E402

pkg = importlib.import_module("parallel-genomeseq_b200")


def check(r, xs, y, **kw):
    for i, x in enumerate(xs):
        w = o.align(x, y, **kw)
        got = (int(r["score"][i]), int(r["pos"][i]), r["cx"][i], r["cy"][i])
        assert got == (w["score"], w["pos"], w["cx"], w["cy"]), (i, got[:2], w["score"], w["pos"])


def main():
    rng = np.random.default_rng(1)
    eng = pkg.Engine(0)
    y = "".join(rng.choice(list("ACGT"), size=1200))
    xs = [y[s:s + 125] for s in (0, 300, 777, 1000)] + ["".join(rng.choice(list("ACGT"), size=90))]
    for mode, omode in ((pkg.MODE_SAT_U8, o.MODE_SAT_U8), (pkg.MODE_EXACT, o.MODE_EXACT)):
        eng.set_scoring_match(mode, 3, -3, 2)
        eng.set_reference(y)
        check(eng.align(xs, cons_stride=1500), xs, y, mode=omode)
        r = eng.align(xs[:2], npiece=3, ratio=2.0, cons_stride=1500)
        for i, x in enumerate(xs[:2]):
            w = o.align_chunked(x, y, 3, 2.0, mode=omode)
            assert (int(r["score"][i]), int(r["pos"][i])) == (w["score"], w["pos"])
    # row strips by one warp (many pairs is not needed: forced off the pipelined path) and pipelined strips
    ref = pkg.synth.c3_reference(6000, seed=41)
    reads = pkg.synth.mutated_reads(ref, 2, 1200, seed=42, sub=0.03, ins=0.003, dele=0.003)
    for env in ({"SWB_NO_PIPELINE": "1"}, {}, {"SWB_COLS": "8"}, {"SWB_FORCE_WIDE": "1"}):
        os.environ.update(env)
        for mode, omode in ((pkg.MODE_EXACT, o.MODE_EXACT), (pkg.MODE_SAT_U8, o.MODE_SAT_U8)):
            eng.set_scoring_match(mode, 3, -3, 2)
            eng.set_reference(ref)
            check(eng.align(reads, cons_stride=4000), reads, ref, mode=omode)
        for k in env:
            os.environ.pop(k)
    # query-stationary kernels (forced: the batch is small)
    os.environ["SWB_QSTAT"] = "1"
    q = pkg.synth.c4_queries(1, 120, seed=5)[0]
    db = [p[:200] for p in pkg.synth.c4_database(24, seed=6)]
    t = pkg.synth.blosum62_table()
    eng.set_scoring_table(pkg.MODE_EXACT, t, 10)
    eng.set_reference(q)
    r = eng.align(db, cons_stride=800)
    for i, p in enumerate(db):
        w = o.align(p, q, mode=o.MODE_EXACT, table=t, gap=10)
        if w["score"] > 0:
            assert (int(r["score"][i]), int(r["pos"][i]), r["cx"][i]) == (w["score"], w["pos"], w["cx"])
    os.environ.pop("SWB_QSTAT")
    eng.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
    eng.set_reference("TGTTACGG")
    assert (eng.matrix("GGTTGACTA") == o.matrix("GGTTGACTA", "TGTTACGG")).all()
    eng.close()
    print("sanitize case ok:", pkg.load_library().swb_version().decode())


if __name__ == "__main__":
    main()
