#!/usr/bin/env python3
"""Throughput of all five BASELINE.json configs on one GPU (device-timed GCUPS, reference cell convention).

C1 data_small (shipped reads) x1 and replicated x64; C2 chunked (4 and 17 pieces); C3 is bench.py; C4 protein
DB search (BLOSUM62, gap 10, EXACT, x = DB protein, y = 300-aa query); C5 long pair (10 kbp reads, EXACT and
SAT_U8) against a reference of --c5-ref bases.  Writes one JSON document to stdout."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(eng, reps=3):
    eng.run()
    us = [eng.run() for _ in range(reps)]
    return float(np.median(us)), eng.stats()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c4-proteins", type=int, default=100_000)
    ap.add_argument("--c5-ref", type=int, default=2_000_000)
    ap.add_argument("--c5-reads", type=int, default=16)
    args = ap.parse_args()
    from conftest import read_fasta, read_truth, GOLDEN
    pkg = importlib.import_module("parallel-genomeseq_b200")
    eng = pkg.Engine(0)
    out = {}
    ref = read_fasta(os.path.join(GOLDEN, "data_small", "genome.chr22.5K.fa"))
    reads = [t[2] for t in read_truth(os.path.join(GOLDEN, "data_small", "data_small_ground_truth.csv"))]
    eng.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
    eng.set_reference(ref)
    for name, rs, npiece in (("C1 data_small x1", reads, 0), ("C1 data_small x64", reads * 64, 0), ("C2 data_small 4 pieces x1", reads, 4),
                             ("C2 data_small 4 pieces x64", reads * 64, 4), ("C2 data_small 17 pieces x64", reads * 64, 17)):
        eng.stage(rs, npiece=npiece, ratio=2.0, consensus=True)
        us, st = timed(eng)
        out[name] = dict(gcups=st["cells_reference"] / us / 1e3, device_us=us, reads=len(rs), geometry=[st["lanes_per_pair"], st["rows_per_lane"], st["block_steps"]],
                         pass2_share=st["pass2_us"] / us, executed_over_reference_cells=st["cells_executed"] / st["cells_reference"])
    # C4
    query = pkg.synth.c4_queries(1, 300)[0]
    db = pkg.synth.c4_database(args.c4_proteins)
    eng.set_scoring_table(pkg.MODE_EXACT, pkg.synth.blosum62_table(), 10)
    eng.set_reference(query)
    t0 = time.perf_counter()
    eng.stage(db, consensus=False)
    stage_s = time.perf_counter() - t0
    us, st = timed(eng)
    out["C4 protein DB (BLOSUM62, gap 10)"] = dict(gcups=st["cells_reference"] / us / 1e3, device_us=us, proteins=len(db), residues=int(sum(len(p) for p in db)),
                                                   stage_seconds=stage_s, pass2_share=st["pass2_us"] / us,
                                                   executed_over_reference_cells=st["cells_executed"] / st["cells_reference"])
    eng.set_scoring_match(pkg.MODE_EXACT, 3, -3, 2)
    eng.stage(db, consensus=False)
    us, st = timed(eng)
    out["C4 protein DB (+3/-3, gap 2: the reference default)"] = dict(gcups=st["cells_reference"] / us / 1e3, device_us=us, pass2_share=st["pass2_us"] / us)
    # C5
    lref = pkg.synth.c5_reference(args.c5_ref)
    lreads = pkg.synth.c5_reads(lref, args.c5_reads, 10_000)
    for mode, nm in ((pkg.MODE_EXACT, "EXACT"), (pkg.MODE_SAT_U8, "SAT_U8")):
        eng.set_scoring_match(mode, 3, -3, 2)
        eng.set_reference(lref)
        eng.stage(lreads, consensus=True, cons_stride=25_000)
        us, st = timed(eng, reps=2)
        out[f"C5 long pair {nm} ({args.c5_reads} x 10 kbp vs {args.c5_ref} bp)"] = dict(gcups=st["cells_reference"] / us / 1e3, device_us=us, pass2_share=st["pass2_us"] / us,
                                                                                  geometry=[st["lanes_per_pair"], st["rows_per_lane"], st["block_steps"]])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
