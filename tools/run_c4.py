import sys, os, importlib, time
sys.path.insert(0, '.')
pkg = importlib.import_module('parallel-genomeseq_b200')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
query = pkg.synth.c4_queries(1, 300)[0]
db = pkg.synth.c4_database(n)
e = pkg.Engine(0)
e.set_scoring_table(pkg.MODE_EXACT, pkg.synth.blosum62_table(), 10)
e.set_reference(query)
e.stage(db, consensus=False)
for _ in range(3):
    us = e.run(); st = e.stats()
    print(us, st, file=sys.stderr)
