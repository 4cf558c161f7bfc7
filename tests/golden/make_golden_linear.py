#!/usr/bin/env python3
"""Golden vectors for the shapes the reference itself cannot materialise, computed with the
LINEAR-MEMORY oracle (oracle/sw_oracle_linear.c — pinned against sw_oracle.c, oracle/_ref and every other golden
in tests/test_oracle.py).  Runs on CPU only; takes tens of core-minutes, which is why its outputs are committed:

  c5_full.json   BASELINE config 5 at its STATED size: the 16 seeded 10 kbp reads (synth.c5_reads) against the seeded
                 51 Mbp reference (synth.c5_reference), EXACT and SAT_U8, default scoring +3/-3/2.  Per read and mode:
                 score, pos, arg-max cell, consensus length and sha256 of both consensus strings (full strings for the
                 first two reads).
  c3_10k.json    BASELINE config 3 on 10 000 reads (synth.c3_reads, indels included) against the seeded 1 Mbp reference,
                 SAT_U8: score, pos, arg-max cell, consensus length and a 64-bit digest of the two consensus strings.

    python tests/golden/make_golden_linear.py [c5] [c3] [--procs N] [--c5-reads N]
"""
import argparse
import hashlib
import json
import os
import sys
import time
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyoracle as o  # noqa: E402
import synth  # noqa: E402

_REF = None


def sha(s):
    return hashlib.sha256(s.encode("latin-1")).hexdigest()


def digest64(cx, cy):
    return hashlib.sha256((cx + "|" + cy).encode("latin-1")).hexdigest()[:16]


def _init(kind):
    global _REF
    _REF = synth.c5_reference() if kind == "c5" else synth.c3_reference()


def _c5_job(job):
    idx, x, mode = job
    t0 = time.time()
    w = o.align(x, _REF, mode=mode, linear=True)
    return idx, mode, w, time.time() - t0


def _c3_job(job):
    lo, xs = job
    out = []
    for x in xs:
        w = o.align(x, _REF, mode=o.MODE_SAT_U8, linear=True)
        out.append([w["score"], w["pos"], w["end"][0], w["end"][1], len(w["cx"]), digest64(w["cx"], w["cy"])])
    return lo, out


def make_c5(procs, n_reads):
    ref = synth.c5_reference()
    reads = synth.c5_reads(ref, 16, 10_000)[:n_reads]
    jobs = [(i, x, mode) for mode in (o.MODE_EXACT, o.MODE_SAT_U8) for i, x in enumerate(reads)]
    res = {}
    with Pool(procs, initializer=_init, initargs=("c5",)) as pool:
        for idx, mode, w, dt in pool.imap_unordered(_c5_job, jobs):
            print(f"c5 read {idx} mode {mode}: score {w['score']} pos {w['pos']} len {len(w['cx'])}  ({dt:.0f} s)", flush=True)
            e = dict(score=w["score"], pos=w["pos"], end=list(w["end"]), len=len(w["cx"]), cx_sha256=sha(w["cx"]), cy_sha256=sha(w["cy"]))
            if idx < 2:
                e["cx"], e["cy"] = w["cx"], w["cy"]
            res[(idx, mode)] = e
    doc = dict(ref_len=len(ref), ref_seed=26, ref_sha256=sha(ref), read_len=10_000, reads_seed=27, n_reads=len(reads),
               reads_sha256=[sha(x) for x in reads], scoring=dict(match=3, mismatch=-3, gap=2),
               exact=[res[(i, o.MODE_EXACT)] for i in range(len(reads))], sat_u8=[res[(i, o.MODE_SAT_U8)] for i in range(len(reads))])
    with open(os.path.join(HERE, "c5_full.json"), "w") as f:
        json.dump(doc, f, indent=0)


def make_c3(procs, n_reads=10_000):
    ref = synth.c3_reference()
    reads = synth.c3_reads(ref, n_reads)
    step = 50
    jobs = [(lo, reads[lo:lo + step]) for lo in range(0, n_reads, step)]
    rows = [None] * n_reads
    done = 0
    with Pool(procs, initializer=_init, initargs=("c3",)) as pool:
        for lo, out in pool.imap_unordered(_c3_job, jobs):
            rows[lo:lo + len(out)] = out
            done += len(out)
            if done % 1000 == 0:
                print(f"c3 {done}/{n_reads}", flush=True)
    doc = dict(ref_len=len(ref), ref_seed=22, ref_sha256=sha(ref), read_len=150, reads_seed=23, n_reads=n_reads,
               reads_sha256=sha("".join(reads)), scoring=dict(match=3, mismatch=-3, gap=2),
               columns=["score", "pos", "end_x", "end_y", "len", "digest64(cx|cy)"], rows=rows)
    with open(os.path.join(HERE, "c3_10k.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["c3", "c5"])
    ap.add_argument("--procs", type=int, default=max(1, (os.cpu_count() or 2) - 2))
    ap.add_argument("--c5-reads", type=int, default=16)
    a = ap.parse_args()
    o.build_oracle()
    if "c3" in a.what:
        make_c3(a.procs)
    if "c5" in a.what:
        make_c5(a.procs, a.c5_reads)
