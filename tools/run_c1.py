import sys, os, importlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import read_fasta, read_truth, GOLDEN
pkg = importlib.import_module('parallel-genomeseq_b200')
ref = read_fasta(os.path.join(GOLDEN, 'data_small', 'genome.chr22.5K.fa'))
reads = [t[2] for t in read_truth(os.path.join(GOLDEN, 'data_small', 'data_small_ground_truth.csv'))]
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 64
e = pkg.Engine(0)
e.set_scoring_match(0, 3, -3, 2); e.set_reference(ref)
e.stage(reads * rep, consensus=True)
for _ in range(3):
    us = e.run(); st = e.stats()
    print(us, st['pass1_us'], st['pass2_us'])
