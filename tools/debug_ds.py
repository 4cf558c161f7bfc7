import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import importlib, numpy as np
from conftest import read_fasta, read_truth, read_golden_csv, GOLDEN
pkg = importlib.import_module('parallel-genomeseq_b200')
ref = read_fasta(os.path.join(GOLDEN, 'data_small', 'genome.chr22.5K.fa'))
truth = read_truth(os.path.join(GOLDEN, 'data_small', 'data_small_ground_truth.csv'))
gold = read_golden_csv('data_small_sw_skewed.csv')
e = pkg.Engine(0)
e.set_scoring_match(0, 3, -3, 2); e.set_reference(ref)
r = e.align([t[2] for t in truth])
bad = [i for i, g in enumerate(gold) if (int(r['score'][i]), int(r['pos'][i]), r['cx'][i], r['cy'][i]) != (g['score'], g['pos'], g['cx'], g['cy'])]
print('bad', len(bad), bad[:20], e.stats())
for i in bad[:5]:
    print(i, int(r['score'][i]), int(r['pos'][i]), tuple(r['end'][i]), gold[i]['score'], gold[i]['pos'], len(r['cx'][i]), len(gold[i]['cx']), r['flags'][i])
