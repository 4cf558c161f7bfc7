#!/usr/bin/env python3
"""Stage one BASELINE config and run it a few times (profiling / SWB_DEBUG driver; prints device times).

    python tools/run_config.py c1 | c1x64 | c2x64 | c4 [proteins] | c5 [ref_len] [mode] | c3 [reads]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("parallel-genomeseq_b200")


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "c1"
    args = sys.argv[2:]
    reps = int(os.environ.get("RUN_REPS", "3"))
    e = pkg.Engine(0)
    if what in ("c1", "c1x64", "c2x64"):
        from conftest import read_fasta, read_truth, GOLDEN
        ref = read_fasta(os.path.join(GOLDEN, "data_small", "genome.chr22.5K.fa"))
        reads = [t[2] for t in read_truth(os.path.join(GOLDEN, "data_small", "data_small_ground_truth.csv"))]
        e.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
        e.set_reference(ref)
        e.stage(reads * (64 if what.endswith("x64") else 1), npiece=4 if what.startswith("c2") else 0, ratio=2.0, consensus=True)
    elif what == "c4":
        n = int(args[0]) if args else 100_000
        e.set_scoring_table(pkg.MODE_EXACT, pkg.synth.blosum62_table(), 10)
        e.set_reference(pkg.synth.c4_queries(1, 300)[0])
        e.stage(pkg.synth.c4_database(n), consensus=False)
    elif what == "c5":
        nref = int(args[0]) if args else 4_000_000
        mode = int(args[1]) if len(args) > 1 else 1
        lref = pkg.synth.c5_reference(nref)
        e.set_scoring_match(mode, 3, -3, 2)
        e.set_reference(lref)
        e.stage(pkg.synth.c5_reads(lref, 16, 10_000), consensus=True, cons_stride=25_000)
    elif what == "c3":
        import numpy as np
        n = int(args[0]) if args else 37_888
        ref = pkg.synth.c3_reference()
        reads = pkg.synth.mutated_reads_fast(np.frombuffer(ref.encode(), dtype=np.uint8), n, 150, seed=2300)
        e.set_scoring_match(pkg.MODE_SAT_U8, 3, -3, 2)
        e.set_reference(ref)
        e.stage(reads, consensus=True, cons_stride=364)
    else:
        raise SystemExit(__doc__)
    for _ in range(reps):
        us = e.run()
        st = e.stats()
        print(f"{what}: {us:.1f} us  pass1 {st['pass1_us']:.1f}  pass2 {st['pass2_us']:.1f}  geometry {st['lanes_per_pair']}x{st['rows_per_lane']} B={st['block_steps']}  "
              f"GCUPS {st['cells_reference'] / us / 1e3:.1f}", flush=True)
    e.close()


if __name__ == "__main__":
    main()
