#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by running the REFERENCE ITSELF.

Runs only in the build container (needs /root/reference): the reference aligner classes are compiled
by oracle/build_ref.py into oracle/_ref/libref_aligner.so and driven through oracle/pyoracle.py.
The GPU box never reads /root/reference — it reads the small files this script commits:

  data_small/genome.chr22.5K.fa, data_small/data_small_ground_truth.csv
        the reference's own integration fixture (data/data_small/, data/data_small_ground_truth.csv);
        DATA, copied byte-for-byte so BASELINE config 1/2 can run anywhere.
  data_small_sw_skewed.csv    SWAligner<Similarity_Matrix_Skewed>(read, ref)            (config 1)
  data_small_sw_float.csv     SWAligner<Similarity_Matrix>(read, ref)                   (EXACT semantics)
  data_small_p4.csv           serial OMPParallelLocalAligner<Skewed,SWAligner<Skewed>>(read, ref, 4, 2.0)   (config 2)
  data_small_p17.csv          ... (read, ref, 17, 2.0)  (the value sw_solve_small.cpp:82 uses)
  random_pairs.json           seeded random non-square pairs, several scorings, both SMTs, chunked and not
  c3_sample.json              config-3-shaped sample: 150 bp reads vs a seeded 1 Mbp synthetic reference
  c4_sample.json              config-4-shaped sample: protein DB entries vs a 300-aa query, BLOSUM62 callback
  SHA256SUMS                  sha256 of every file above
"""
import csv
import hashlib
import json
import os
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyoracle as o  # noqa: E402
import synth  # noqa: E402

REF = "/root/reference"


def read_fasta(path):
    with open(path) as f:
        lines = f.read().split("\n")
    return "".join(lines[1:])  # sw_solve_small.cpp:25-30: skip line 0, concatenate the rest


def read_truth(path):
    rows = []
    with open(path) as f:
        for i, line in enumerate(f):
            if i == 0 or not line.strip():
                continue
            r = line.rstrip("\n").split(",")  # sw_solve_small.cpp:56-67: field 2 = SEQ
            rows.append((int(r[0]), r[1], r[2], int(r[3])))
    return rows


def dump(path, rows):
    with open(path, "w", newline="") as f:
        w = csv.writer(f, lineterminator="\n")
        w.writerow(["index", "score", "pos", "consensus_x", "consensus_y"])
        for r in rows:
            w.writerow(r)


def main():
    assert o.ref() is not None, "compiled reference unavailable"
    ds = os.path.join(HERE, "data_small")
    os.makedirs(ds, exist_ok=True)
    shutil.copyfile(os.path.join(REF, "data/data_small/genome.chr22.5K.fa"), os.path.join(ds, "genome.chr22.5K.fa"))
    shutil.copyfile(os.path.join(REF, "data/data_small_ground_truth.csv"), os.path.join(ds, "data_small_ground_truth.csv"))
    ref = read_fasta(os.path.join(ds, "genome.chr22.5K.fa"))
    truth = read_truth(os.path.join(ds, "data_small_ground_truth.csv"))
    print(f"data_small: ref {len(ref)} bp, {len(truth)} reads")

    for name, kw in [("data_small_sw_skewed", dict(smt=0)), ("data_small_sw_float", dict(smt=1)),
                     ("data_small_p4", dict(smt=0, npiece=4, ratio=2.0)), ("data_small_p17", dict(smt=0, npiece=17, ratio=2.0))]:
        t0 = time.time()
        rows = []
        for idx, _, seq, _ in truth:
            r = o.ref_align(seq, ref, **kw)
            rows.append((idx, r["score"], r["pos"], r["cx"], r["cy"]))
        dump(os.path.join(HERE, name + ".csv"), rows)
        print(f"{name}: {time.time() - t0:.1f}s, pos==POS {sum(1 for a, b in zip(rows, truth) if a[2] == b[3])}/{len(rows)}")

    # ---- seeded random pairs (non-square: SURVEY F9; always share a character: F10) ------------------
    rng = np.random.default_rng(1234)
    cases = []
    scorings = [dict(kind=0, match=3, mismatch=-3, gap=2), dict(kind=1, match=2, mismatch=-1, gap=1),
                dict(kind=1, match=5, mismatch=-4, gap=3), dict(kind=1, match=1, mismatch=-1, gap=1),
                dict(kind=1, match=7, mismatch=-2, gap=5)]
    shapes = [(1, 2), (2, 1), (3, 7), (7, 3), (9, 8), (31, 64), (32, 33), (33, 31), (64, 65), (65, 63), (40, 200), (200, 40), (125, 400), (150, 333), (97, 1000)]
    for rep in range(12):
        for (m, n) in shapes:
            alpha = "ACGT" if rep % 3 else "ACGTN"
            y = "".join(rng.choice(list(alpha), size=n))
            if rep % 2 == 0 and m <= n:  # a mutated window of y (a realistic read)
                s = int(rng.integers(0, n - m + 1))
                x = list(y[s:s + m])
                for k in range(m):
                    if rng.random() < 0.08:
                        x[k] = str(rng.choice(list("ACGT")))
                x = "".join(x)
            else:
                x = "".join(rng.choice(list(alpha), size=m))
            sc = scorings[(rep + m) % len(scorings)]
            if not (set(x) & set(y)):
                continue  # all-zero matrix: reference UB — it reads H(-1,-1)
            for smt in (0, 1):
                r = o.ref_align(x, y, smt=smt, scoring_kind=sc["kind"], match=sc["match"], mismatch=sc["mismatch"], gap=sc["gap"])
                if r["score"] == 0:
                    continue  # reference UB (F10)
                cases.append(dict(x=x, y=y, smt=smt, scoring=sc, npiece=0, ratio=0.0, score=r["score"], pos=r["pos"], cx=r["cx"], cy=r["cy"]))
    # chunked (OMPParallelLocalAligner, serial) — default and custom scoring (exercises SURVEY F8)
    for rep in range(40):
        m = int(rng.integers(20, 130)); n = int(rng.integers(900, 3000)); npiece = int(rng.choice([2, 3, 4, 5, 8])); ratio = float(rng.choice([1.0, 1.5, 2.0]))
        y = "".join(rng.choice(list("ACGT"), size=n))
        s = int(rng.integers(0, n - m + 1)); x = list(y[s:s + m])
        for k in range(m):
            if rng.random() < 0.05:
                x[k] = str(rng.choice(list("ACGT")))
        x = "".join(x)
        sc = scorings[rep % len(scorings)]
        for smt in (0, 1):
            r = o.ref_align(x, y, smt=smt, scoring_kind=sc["kind"], match=sc["match"], mismatch=sc["mismatch"], gap=sc["gap"], npiece=npiece, ratio=ratio)
            cases.append(dict(x=x, y=y, smt=smt, scoring=sc, npiece=npiece, ratio=ratio, score=r["score"], pos=r["pos"], cx=r["cx"], cy=r["cy"]))
    with open(os.path.join(HERE, "random_pairs.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print(f"random_pairs: {len(cases)} cases")

    # ---- config-3-shaped sample: 150 bp reads vs the seeded 1 Mbp synthetic reference ----------------
    ref1m = synth.c3_reference(1_000_000)
    reads = synth.c3_reads(ref1m, 24, read_len=150)
    out = dict(ref_len=len(ref1m), ref_sha256=hashlib.sha256(ref1m.encode()).hexdigest(), reads=[])
    t0 = time.time()
    for rd in reads:
        r = o.ref_align(rd, ref1m, smt=0)
        out["reads"].append(dict(x=rd, score=r["score"], pos=r["pos"], cx=r["cx"], cy=r["cy"]))
    print(f"c3_sample: {len(reads)} reads in {time.time() - t0:.1f}s")
    with open(os.path.join(HERE, "c3_sample.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))

    # ---- config-4-shaped sample: DB proteins (x) vs 300-aa query (y), BLOSUM62 through the callback ----
    query = synth.c4_queries(1, 300)[0]
    db = synth.c4_database(160)
    table = synth.blosum62_table()
    out = dict(query=query, gap=10, entries=[])
    t0 = time.time()
    for p in db:
        if len(p) == len(query) or not (set(p) & set(query)):
            continue  # square (F9 is Skewed-only, but keep the sample uniform)
        r = o.ref_align(p, query, smt=1, scoring_kind=2, gap=10.0, table=table.astype(np.float32))
        if r["score"] == 0:
            continue
        out["entries"].append(dict(x=p, score=r["score"], pos=r["pos"], cx=r["cx"], cy=r["cy"]))
    print(f"c4_sample: {len(out['entries'])} proteins in {time.time() - t0:.1f}s")
    with open(os.path.join(HERE, "c4_sample.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))

    sums = []
    for root, _, files in os.walk(HERE):
        for fn in sorted(files):
            if fn in ("SHA256SUMS", "make_golden.py") or fn.endswith(".pyc"):
                continue
            p = os.path.join(root, fn)
            with open(p, "rb") as f:
                sums.append(f"{hashlib.sha256(f.read()).hexdigest()}  {os.path.relpath(p, HERE)}")
    with open(os.path.join(HERE, "SHA256SUMS"), "w") as f:
        f.write("\n".join(sorted(sums, key=lambda s: s.split()[1])) + "\n")
    print("\n".join(sums))


if __name__ == "__main__":
    main()
