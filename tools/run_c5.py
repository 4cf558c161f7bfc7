import sys, os, importlib
sys.path.insert(0, '.')
pkg = importlib.import_module('parallel-genomeseq_b200')
nref = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lref = pkg.synth.c5_reference(nref)
lreads = pkg.synth.c5_reads(lref, 16, 10_000)
e = pkg.Engine(0)
e.set_scoring_match(mode, 3, -3, 2); e.set_reference(lref)
e.stage(lreads, consensus=True, cons_stride=25000)
for _ in range(2):
    us = e.run(); st = e.stats()
    print(us, st['pass1_us'], st['pass2_us'], st['rows_per_lane'], st['block_steps'])
