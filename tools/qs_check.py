import importlib, json, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
os.environ["SWB_QSTAT"] = "1"
pkg = importlib.import_module('parallel-genomeseq_b200')
import synth
c4 = json.load(open('tests/golden/c4_sample.json'))
e = pkg.Engine(0)
e.set_scoring_table(pkg.MODE_EXACT, synth.blosum62_table(), c4["gap"])
e.set_reference(c4["query"])
ents = c4["entries"]
r = e.align([x["x"] for x in ents], cons_stride=6000)
print("stats", e.stats())
bad = 0
for i, w in enumerate(ents):
    got = (int(r["score"][i]), int(r["pos"][i]), r["cx"][i], r["cy"][i])
    exp = (w["score"], w["pos"], w["cx"], w["cy"])
    if got != exp:
        bad += 1
        if bad < 6: print("MISMATCH", i, len(w["x"]), got[:2], exp[:2], tuple(r["end"][i]), len(got[2]), len(exp[2]))
print("c4 golden under QS:", len(ents) - bad, "of", len(ents), "ok")
