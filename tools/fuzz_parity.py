#!/usr/bin/env python3
"""Randomised parity fuzzer: CUDA path vs the CPU oracle over random shapes, scorings, modes, chunking and
kernel knobs (select, columns per step, sub-batch size).  Test infrastructure; prints the first mismatch and exits 1."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as o  # noqa: E402

pkg = importlib.import_module("parallel-genomeseq_b200")


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    rng = np.random.default_rng(seed)
    eng = pkg.Engine(0)
    t0 = time.time()
    checked = 0
    for it in range(iters):
        mode = int(rng.integers(0, 2))
        omode = o.MODE_SAT_U8 if mode == 0 else o.MODE_EXACT
        alpha = list("ACGT") if rng.random() < 0.6 else (list("ACGTN") if rng.random() < 0.5 else list("ARNDCQEGHILKMFPSTWYV"))
        n = int(rng.choice([rng.integers(1, 40), rng.integers(40, 700), rng.integers(700, 3500)]))
        y = "".join(rng.choice(alpha, size=n))
        nreads = int(rng.choice([1, 2, 3, 7, 20, 45]))
        xs = []
        for _ in range(nreads):
            kind = rng.random()
            m = int(rng.integers(1, 60)) if kind < 0.3 else (int(rng.integers(60, 420)) if kind < 0.85 else int(rng.integers(1025, 2600)))
            if rng.random() < 0.6 and m < n:
                s0 = int(rng.integers(0, n - m + 1)); x = list(y[s0:s0 + m])
                for q in range(m):
                    if rng.random() < 0.08:
                        x[q] = str(rng.choice(alpha))
                xs.append("".join(x))
            else:
                xs.append("".join(rng.choice(alpha, size=m)))
        table = None
        if mode == 1 and rng.random() < 0.35:
            t = rng.integers(-6, 3, size=(256, 256)).astype(np.int32)
            if rng.random() < 0.5:
                t = np.minimum(t, t.T)      # symmetric like a substitution matrix; otherwise fn(x, y) != fn(y, x)
            np.fill_diagonal(t, rng.integers(1, 12, size=256))
            table = t
            gap = int(rng.integers(0, 8))
        ma, mi, gap2 = int(rng.integers(0, 10)), -int(rng.integers(0, 10)), int(rng.integers(0, 7))
        npiece = 0
        ratio = 2.0
        if rng.random() < 0.3 and n > 200 and table is None:
            npiece = int(rng.integers(1, 7)); ratio = float(rng.choice([1.0, 1.5, 2.0]))
        os.environ["SWB_SELECT"] = str(rng.choice(["profile", "compare"]))
        os.environ["SWB_COLS"] = str(rng.choice(["1", "2"]))
        os.environ["SWB_CHUNK_PAIRS"] = str(rng.choice(["64", "37888"]))
        if rng.random() < 0.5:
            os.environ.pop("SWB_SELECT", None); os.environ.pop("SWB_COLS", None)
        # pass-2 knobs: ring depth, band lanes (2 = a new session every few rows), checkpoint period
        for k, choices in (("SWB_TRACE_WC", ["32", "64"]), ("SWB_TRACE_NB", ["2", "3", "5"]), ("SWB_FORCE_B", ["32", "64", "256", "1024"]), ("SWB_FORCE_WIDE", ["1"])):
            if rng.random() < 0.4:
                os.environ[k] = str(rng.choice(choices))
            else:
                os.environ.pop(k, None)
        try:
            if table is not None:
                eng.set_scoring_table(mode, table, gap)
            else:
                eng.set_scoring_match(mode, ma, mi, gap2)
            eng.set_reference(y)
            r = eng.align(xs, npiece=npiece, ratio=ratio, cons_stride=max(len(x) for x in xs) + n + 2)
        except pkg.SwbError as e:
            if e.code in (-3, -5):      # range precondition / unsupported shape: both are legitimate refusals
                continue
            raise
        for i, x in enumerate(xs):
            kw = dict(mode=omode, table=table, gap=gap) if table is not None else dict(mode=omode, match=ma, mismatch=mi, gap=gap2)
            w = o.align_chunked(x, y, npiece, ratio, **kw) if npiece else o.align(x, y, **kw)
            if w.get("err", 0) <= -10:
                continue
            if w["score"] == 0:
                ok = int(r["score"][i]) == 0 and int(r["len"][i]) == 0
            else:
                ok = (int(r["score"][i]), int(r["pos"][i]), r["cx"][i], r["cy"][i]) == (w["score"], w["pos"], w["cx"], w["cy"])
                if ok and not npiece:
                    ok = tuple(int(v) for v in r["end"][i]) == tuple(w["end"])
                ok = ok and int(r["flags"][i]) == 0
            checked += 1
            if not ok:
                print("MISMATCH iter", it, "read", i, "mode", mode, "m", len(x), "n", n, "npiece", npiece, ratio, "scoring", (ma, mi, gap2) if table is None else ("table", gap),
                      "env", {k: os.environ.get(k) for k in ("SWB_SELECT", "SWB_COLS", "SWB_CHUNK_PAIRS", "SWB_QSTAT", "SWB_TRACE_WC", "SWB_TRACE_NB", "SWB_FORCE_B", "SWB_FORCE_WIDE")})
                print(" flags", int(r["flags"][i]))
                print(" got ", int(r["score"][i]), int(r["pos"][i]), tuple(r["end"][i]), len(r["cx"][i]))
                print(" want", w["score"], w["pos"], w.get("end"), len(w["cx"]))
                sys.exit(1)
    print(f"fuzz ok: {iters} iterations, {checked} alignments checked in {time.time() - t0:.1f}s (seed {seed})")


if __name__ == "__main__":
    main()
